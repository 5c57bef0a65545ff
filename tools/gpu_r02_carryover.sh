# round-2 carry-overs on one B200: launch list + full captures of the final stereo / replay kernels, sanitizer over the
# section-8f kernels, compute-sanitizer over one search of every kind
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bash tools/gpu_frame_side_prof.sh
bash tools/gpu_sanitize_frame_side.sh
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stereo_match -c 1 -f -o gpurun_out/r02_stereo_match python tools/frame_side_once.py > gpurun_out/ncu_r02_stereo.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:proj_replay -s 2 -c 1 -f -o gpurun_out/r02_proj_replay python tools/search_latency.py > gpurun_out/ncu_r02_proj_replay.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:init_replay -s 2 -c 1 -f -o gpurun_out/r02_init_replay python tools/search_latency.py > gpurun_out/ncu_r02_init_replay.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:distinctive -c 1 -f -o gpurun_out/r02_distinctive python tools/frame_side_once.py > gpurun_out/ncu_r02_distinctive.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep
