mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_suite.log 2>&1; echo "suite rc=$?" > gpurun_out/r2c_status.txt
python tools/big_configs.py 1 8 > gpurun_out/r2c_big.jsonl 2> gpurun_out/r2c_big.err
python tools/big_configs.py 1 --opt sor_max_cluster=16 > gpurun_out/r2c_big_c16.jsonl 2>> gpurun_out/r2c_big.err
OFDIS_SOR_TIMING=1 python -m of_dis_b200.build --force > /dev/null 2>&1 && python tools/sor_timing.py > gpurun_out/r2c_sor_timing.txt 2>&1
python -m of_dis_b200.build --force > /dev/null 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_status.txt
cat gpurun_out/r2c_status.txt; tail -5 gpurun_out/r2c_suite.log
