# ncu launch list + per-kernel throughput metrics of the section-8f kernels (one GPU, one pass each)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 240 ncu --clock-control none --csv --log-file gpurun_out/frame_side_launches.csv \
  -k regex:'stereo_|frame_import|undistort_kernel|distinctive|grid_|scan_kernel|bow_descend' \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  python tools/frame_side_once.py > gpurun_out/frame_side_prof.log 2>&1
tail -3 gpurun_out/frame_side_prof.log
wc -l gpurun_out/frame_side_launches.csv
