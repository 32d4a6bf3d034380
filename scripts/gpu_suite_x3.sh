for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/suite_rep_$i.log 2>&1
  echo "full GPU suite run $i exit $? $(tail -1 gpurun_out/suite_rep_$i.log)"
  grep -h "^E \|^FAILED" gpurun_out/suite_rep_$i.log | head -6
done
