tag=r2y
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_suite.log 2>&1; echo "suite rc=$?" > gpurun_out/${tag}_status.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err; echo "reference arm rc=$?" >> gpurun_out/${tag}_status.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_b64.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_status.txt
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
ncu --metrics $M --clock-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b64.csv python tools/one_step.py 64 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b64_warm.csv python tools/one_step.py 64 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b8_warm.csv python tools/one_step.py 8 2 > /dev/null 2>&1
F="--set full --import-source on --clock-control none --cache-control none"
ncu $F -k regex:sor_lane --launch-skip 12 -c 1 -f -o gpurun_out/${tag}_sor_lane python tools/one_step.py 8 2 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_sor_lane.ncu-rep > gpurun_out/${tag}_sor_lane_ncu.txt 2>&1
timeout 600 python tools/big_configs.py 1 8 > gpurun_out/${tag}_big_configs.jsonl 2> gpurun_out/${tag}_big.err; echo "big rc=$?" >> gpurun_out/${tag}_status.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_status.txt
cat gpurun_out/${tag}_status.txt; tail -n 3 gpurun_out/${tag}_suite.log; cat gpurun_out/${tag}_smoke.log | tail -2
