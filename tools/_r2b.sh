tools/gpu_round.sh r2b newtests
python tools/big_configs.py 1 8 > gpurun_out/r2b_big.jsonl 2> gpurun_out/r2b_big.err
python tools/big_configs.py 1 --opt sor_max_cluster=16 > gpurun_out/r2b_big_c16.jsonl 2>> gpurun_out/r2b_big.err
python tools/big_configs.py 1 --opt sor_single_max=64 > gpurun_out/r2b_big_s64.jsonl 2>> gpurun_out/r2b_big.err
python tools/big_configs.py 1 --opt sor_single_max=32 > gpurun_out/r2b_big_s32.jsonl 2>> gpurun_out/r2b_big.err
