# Round-2 end evidence on one B200: GPU tests, smoke, the full bench line, the ncu launch list of two steady-state steps,
# one `ncu --set full` capture per extraction kernel (256-frame launches) and of the brute-force matcher.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 300 gpurun_out/r02_bench_reference.err
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming --no-latency --no-allpairs --no-kitti"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:pyramid|fast_warp|octree|blur_|brief" -s 57 -c 38 --csv --log-file gpurun_out/r02_launches_final.csv $B > /dev/null 2>&1
for k in fast_warp octree_kernel brief_staged; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_final_$k $B > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blur_staged -s 16 -c 1 -f -o gpurun_out/r02_final_blur_staged $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid_resize3 -s 14 -c 1 -f -o gpurun_out/r02_final_pyramid_resize3 $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid_level0_bulk -s 2 -c 1 -f -o gpurun_out/r02_final_pyramid_level0 $B > /dev/null 2>&1
# whole-stage DRAM traffic: every launch of two steps, dram bytes per kernel
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:pyramid|fast_warp|octree|blur_|brief" -s 57 -c 19 --csv --log-file gpurun_out/r02_traffic_final.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_scan -s 1 -c 1 -f -o gpurun_out/r02_final_bf_scan python bench.py --steps 1 --warmup 3 --frames 64 --no-cpu --no-latency --no-allpairs --no-kitti > /dev/null 2>&1
ls -la gpurun_out | tail -14
