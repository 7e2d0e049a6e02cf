#!/bin/bash
# One gpurun call: cluster-SOR tests first, then the whole GPU suite, the large configurations and a
# short bench.  Everything is logged under gpurun_out/<tag>_*.  usage: tools/gpu_round.sh <tag> [steps...]
tag=${1:-r2}; shift
steps=${@:-"newtests suite big bench"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
for s in $steps; do
  case $s in
    newtests)
      timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster or taller or tall_level" > gpurun_out/${tag}_newtests.log 2>&1
      echo "newtests rc=$?" >> gpurun_out/${tag}_status.txt ;;
    suite)
      timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_suite.log 2>&1
      echo "suite rc=$?" >> gpurun_out/${tag}_status.txt ;;
    big)
      timeout 600 python tools/big_configs.py 1 8 > gpurun_out/${tag}_big.jsonl 2> gpurun_out/${tag}_big.err
      echo "big rc=$?" >> gpurun_out/${tag}_status.txt ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "bench rc=$?" >> gpurun_out/${tag}_status.txt ;;
    sanitizer)
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/${tag}_memcheck.log 2>&1
      echo "memcheck rc=$?" >> gpurun_out/${tag}_status.txt
      timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/${tag}_racecheck.log 2>&1
      echo "racecheck rc=$?" >> gpurun_out/${tag}_status.txt ;;
    *) bash -c "$s" >> gpurun_out/${tag}_extra.log 2>&1 ;;
  esac
done
cat gpurun_out/${tag}_status.txt
tail -n 15 gpurun_out/${tag}_newtests.log gpurun_out/${tag}_suite.log 2>/dev/null | tail -n 40
