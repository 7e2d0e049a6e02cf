#!/bin/bash
# Evidence pass of a round (one gpurun call): ncu launch lists of one 64-pair step (cold and warm caches),
# --set full captures of the dominant kernels, SOR phase stamps, sanitizers.  Output: gpurun_out/<tag>_*.
tag=${1:-r2}
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
ncu --metrics $M --clock-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b64.csv python tools/one_step.py 64 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none --launch-skip 48 --csv --log-file gpurun_out/${tag}_launches_step_b64_warm.csv python tools/one_step.py 64 2 > /dev/null 2>&1
F="--set full --import-source on --clock-control none --cache-control none"
ncu $F -k regex:sor_wave --launch-skip 26 -c 1 -f -o gpurun_out/${tag}_sor python tools/one_step.py 64 2 > /dev/null 2>&1
ncu $F -k regex:patch_p8c1 --launch-skip 5 -c 1 -f -o gpurun_out/${tag}_patch8 python tools/one_step.py 64 2 > /dev/null 2>&1
ncu $F -k regex:assemble --launch-skip 26 -c 1 -f -o gpurun_out/${tag}_asm python tools/one_step.py 64 2 > /dev/null 2>&1
ncu $F -k regex:patch_p12 --launch-skip 5 -c 1 -f -o gpurun_out/${tag}_patch12 python tools/one_step_big.py cfg5 8 > /dev/null 2>&1
ncu $F -k regex:sor_wave --launch-skip 25 -c 1 -f -o gpurun_out/${tag}_sor_cluster python tools/one_step_big.py cfg5 8 > /dev/null 2>&1
for r in sor patch8 asm patch12 sor_cluster; do python tools/ncu_summary.py gpurun_out/${tag}_$r.ncu-rep > gpurun_out/${tag}_${r}_ncu.txt 2>&1; done
OFDIS_SOR_TIMING=1 python -m of_dis_b200.build --force > /dev/null 2>&1 && python tools/sor_timing.py > gpurun_out/${tag}_sor_timing.txt 2>&1
python -m of_dis_b200.build --force > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?" > gpurun_out/${tag}_evidence_status.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${tag}_evidence_status.txt
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/${tag}_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/${tag}_evidence_status.txt
cat gpurun_out/${tag}_evidence_status.txt; ls -la gpurun_out | grep ${tag}_ | head -40
