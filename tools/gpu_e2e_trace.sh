# per-chunk timeline of the host pipeline (ORBB_PIPE_TRACE=1): h2d done / compute done / d2h done per chunk, in ms
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
env "$@" ORBB_PIPE_TRACE=1 python bench.py --steps 2 --warmup 1 --no-cpu --no-hamming --no-latency --no-allpairs --no-kitti > /dev/null 2> gpurun_out/e2e_trace.txt
grep -c chunk gpurun_out/e2e_trace.txt
