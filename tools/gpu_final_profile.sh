# Round-end evidence: GPU tests, bench line, ncu launch list, one ncu --set full capture per kernel.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming --no-latency"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:pyramid|fast_cells|octree|blur_kernel|brief" -s 36 -c 24 --csv --log-file gpurun_out/launches_final.csv $B > /dev/null 2>&1
for k in fast_cells octree_kernel blur_kernel brief_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/final_$k $B > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid_resize -s 14 -c 1 -f -o gpurun_out/final_pyramid_resize $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_scan -s 1 -c 1 -f -o gpurun_out/final_bf_scan python bench.py --steps 1 --warmup 3 --frames 64 --no-cpu --no-latency > /dev/null 2>&1
ls -la gpurun_out | tail -12
