mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2i_suite.log 2>&1; echo "suite rc=$?" > gpurun_out/r2i_status.txt
python tools/big_configs.py 1 8 > gpurun_out/r2i_big.jsonl 2>> gpurun_out/r2i_big.err
python tools/big_configs.py 1 8 --opt patch_window_tma=1 > gpurun_out/r2i_big_tma.jsonl 2>> gpurun_out/r2i_big.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --opt sor_single_max=32 > gpurun_out/r2i_bench_s32.json 2> gpurun_out/r2i_bench_s32.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
cat gpurun_out/r2i_status.txt; grep -E "passed|failed" gpurun_out/r2i_suite.log | tail -3; grep FAILED gpurun_out/r2i_*.log | head
