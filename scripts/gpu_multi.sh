#!/bin/bash
# usage: gpu_multi.sh N  - sharded verification + bench.py at N GPUs (cfg3 weak, cfg5 strong), as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/verify_sharded.py > gpurun_out/verify_sharded_${N}gpu.log 2>&1; echo "verify exit $?"; tail -4 gpurun_out/verify_sharded_${N}gpu.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"; head -c 400 gpurun_out/bench_${N}gpu.json; echo
timeout 300 $TR scripts/pcie_bw_multi.py > gpurun_out/pcie_bw_${N}gpu.txt 2>&1; cat gpurun_out/pcie_bw_${N}gpu.txt | tail -2
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload cfg5 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg5_${N}gpu.json 2> gpurun_out/bench_cfg5_${N}gpu.err; echo "cfg5 exit $?"; head -c 400 gpurun_out/bench_cfg5_${N}gpu.json; echo
