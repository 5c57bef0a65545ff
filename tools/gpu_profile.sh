# One GPU trip: full GPU test suite, bench line, ncu launch list, ncu --set full on the top kernels.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $B > /dev/null 2>&1
for k in fast_cells octree_kernel blur_kernel pyramid_resize brief_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_$k $B > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_scan -s 1 -c 1 -f -o gpurun_out/prof_bf_scan python bench.py --steps 1 --warmup 3 --frames 64 --no-cpu > /dev/null 2>&1
ls -la gpurun_out
