cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:proj_replay -s 2 -c 1 -f -o gpurun_out/r01c_proj_replay python tools/search_latency.py > gpurun_out/ncu_search_prof.log 2>&1
ls -la gpurun_out/r01c_*replay*
