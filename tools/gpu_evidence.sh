#!/bin/bash
# Evidence pass of a round (one gpurun call): bench lines, ncu launch lists of one step (64 pairs: sor_wave_kernel;
# 8 pairs: sor_lane_kernel), --set full captures of the dominant kernels, the A/B of the two exact SOR kernels, the
# large configurations, racecheck with the barrier-synchronised kernels only.  Output: gpurun_out/<tag>_*.
tag=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?" > gpurun_out/${tag}_status.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_b64.json 2> gpurun_out/${tag}_bench_b64.err; echo "bench rc=$?" >> gpurun_out/${tag}_status.txt
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
ncu --metrics $M --clock-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b64.csv python tools/one_step.py 64 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b64_warm.csv python tools/one_step.py 64 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b8_warm.csv python tools/one_step.py 8 2 > /dev/null 2>&1
F="--set full --import-source on --clock-control none --cache-control none"
ncu $F -k regex:sor_lane --launch-skip 12 -c 1 -f -o gpurun_out/${tag}_sor_lane python tools/one_step.py 8 2 > /dev/null 2>&1
ncu $F -k regex:sor_wave --launch-skip 12 -c 1 -f -o gpurun_out/${tag}_sor_wave python tools/one_step.py 64 2 > /dev/null 2>&1
ncu $F -k regex:assemble --launch-skip 12 -c 1 -f -o gpurun_out/${tag}_asm python tools/one_step.py 64 2 > /dev/null 2>&1
ncu $F -k regex:patch_p8c1 --launch-skip 2 -c 1 -f -o gpurun_out/${tag}_patch8 python tools/one_step.py 64 2 > /dev/null 2>&1
for r in sor_lane sor_wave asm patch8; do python tools/ncu_summary.py gpurun_out/${tag}_$r.ncu-rep > gpurun_out/${tag}_${r}_ncu.txt 2>&1; done
rm -f gpurun_out/${tag}_sor_wave.ncu-rep gpurun_out/${tag}_asm.ncu-rep gpurun_out/${tag}_patch8.ncu-rep
timeout 300 python tools/lane_ab.py 1 8 64 > gpurun_out/${tag}_lane_ab.jsonl 2> gpurun_out/${tag}_lane_ab.err; echo "lane_ab rc=$?" >> gpurun_out/${tag}_status.txt
timeout 600 python tools/big_configs.py 1 8 > gpurun_out/${tag}_big_configs.jsonl 2> gpurun_out/${tag}_big.err; echo "big rc=$?" >> gpurun_out/${tag}_status.txt
SANITIZER_LANE=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/${tag}_racecheck_wave.log 2>&1; echo "racecheck (sor_lane 0) rc=$?" >> gpurun_out/${tag}_status.txt
cat gpurun_out/${tag}_status.txt; ls -la gpurun_out | grep ${tag}_ | head -40
