# round-2 checkpoint on one B200: every GPU test, the smoke, a short bench (extraction + Hamming + latency legs)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -5 gpurun_out/r02_bench.err
} > gpurun_out/r02_check.log 2>&1
tail -30 gpurun_out/r02_check.log
tail -c 6000 gpurun_out/r02_bench.json
