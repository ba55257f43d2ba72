#!/bin/bash
# Round-3-session evidence run (under gpurun, one B200): bench line, ncu launch list of the SAME command, DRAM traffic of the
# dominant kernel family, one `ncu --set full` capture each of the new kernels.  Outputs land in gpurun_out/r3p_*.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r3p_bench_b8.json 2> $O/r3p_bench_b8.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r3p_bench_reference.json 2> $O/r3p_bench_reference.err; echo "reference rc=$?"
L=$(python -c "import json; print(json.load(open('$O/r3p_bench_b8.json'))['launches_per_step'])")
echo "launches per step: $L"
# every dp:: launch of one timed step with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:dp:: -s $((3 * L)) -c $L --csv \
    --log-file $O/r3p_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-train > $O/r3p_ncu_launches.log 2>&1
echo "launch list rc=$?"
# DRAM traffic of every launch of the dominant family (conv3d_stack) in one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:conv3d_stack -s 0 -c 40 --csv --log-file $O/r3p_stack_traffic.csv python scripts/gpu_layers.py 8 128 r3p > $O/r3p_stack_traffic.log 2>&1
echo "traffic rc=$?"
# full capture: the split-half folded 3^3 stacked conv (first three launches: 16->16 and 32->16 at 128^3)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3d_stack3h_kernel" \
    -c 3 -o $O/r3p_new_kernels python scripts/gpu_layers.py 8 128 r3pn > $O/r3p_new_kernels.log 2>&1
echo "full capture rc=$?"
ls -la $O/r3p_*
