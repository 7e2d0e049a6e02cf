mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2d_suite.log 2>&1; echo "suite rc=$?" > gpurun_out/r2d_status.txt
python tools/big_configs.py 1 8 > gpurun_out/r2d_big.jsonl 2> gpurun_out/r2d_big.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> gpurun_out/r2d_status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; echo "benchref rc=$?" >> gpurun_out/r2d_status.txt
cat gpurun_out/r2d_status.txt; grep -E "passed|failed" gpurun_out/r2d_suite.log | tail -3; grep FAILED gpurun_out/r2d_suite.log | head
