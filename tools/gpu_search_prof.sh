cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:proj_replay -s 2 -c 1 -f -o gpurun_out/r01c_proj_replay python tools/search_latency.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:init_replay -s 2 -c 1 -f -o gpurun_out/r01c_init_replay python tools/search_latency.py > /dev/null 2>&1
ls -la gpurun_out/r01c_*replay*
