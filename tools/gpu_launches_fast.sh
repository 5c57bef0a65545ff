set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming --no-latency"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k "regex:fast_" -s 6 -c 2 --csv --log-file gpurun_out/launches_fast.csv $B > /dev/null 2>&1
