set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "conv3d" > $O/r3q_ktests.log 2>&1; echo "ktests rc=$?"; tail -2 $O/r3q_ktests.log
python scripts/gpu_layers.py 8 128 r3q > $O/r3q_layers.log 2>&1; head -1 $O/r3q_layers.log
python scripts/gpu_layers.py 8 128 r3q2 > $O/r3q2_layers.log 2>&1; head -1 $O/r3q2_layers.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:conv3d_stack -s 0 -c 40 --csv --log-file $O/r3q_stack_traffic.csv python scripts/gpu_layers.py 8 128 r3qt > $O/r3q_stack_traffic.log 2>&1
echo "traffic rc=$?"
