set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:pyramid|fast_cells|octree|blur_kernel|brief" -s 36 -c 12 --csv --log-file gpurun_out/launches2.csv $B > /dev/null 2>&1
