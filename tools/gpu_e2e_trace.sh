cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
ORBB_PIPE_TAIL=100000 ORBB_PIPE_TRACE=1 python bench.py --steps 2 --warmup 1 --no-cpu --no-hamming --no-latency > /dev/null 2> gpurun_out/e2e_trace_old.txt
ORBB_PIPE_TRACE=1 python bench.py --steps 2 --warmup 1 --no-cpu --no-hamming --no-latency > /dev/null 2> gpurun_out/e2e_trace_new.txt
grep -c chunk gpurun_out/e2e_trace_old.txt gpurun_out/e2e_trace_new.txt
