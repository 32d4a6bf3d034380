bash scripts/gpu_final.sh
bash scripts/gpu_r02e_launches.sh
bash scripts/gpu_r02e_evidence.sh
