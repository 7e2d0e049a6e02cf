mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2g_suite.log 2>&1; echo "suite rc=$?" > gpurun_out/r2g_status.txt
for rt in 1 2; do python tools/big_configs.py 1 --opt sor_rows_per_thread=$rt > gpurun_out/r2g_big_rt$rt.jsonl 2>> gpurun_out/r2g_big.err; done
python tools/big_configs.py 8 > gpurun_out/r2g_big_b8.jsonl 2>> gpurun_out/r2g_big.err
python tools/big_configs.py 1 --opt sor_max_cluster=16 > gpurun_out/r2g_big_c16.jsonl 2>> gpurun_out/r2g_big.err
for rt in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --opt sor_rows_per_thread=$rt > gpurun_out/r2g_bench_rt$rt.json 2> gpurun_out/r2g_bench_rt$rt.err; echo "bench rt$rt rc=$?" >> gpurun_out/r2g_status.txt; done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/r2g_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2g_status.txt
cat gpurun_out/r2g_status.txt; grep -E "passed|failed" gpurun_out/r2g_suite.log | tail -3; grep FAILED gpurun_out/r2g_*.log | head; tail -5 gpurun_out/r2g_memcheck.log
