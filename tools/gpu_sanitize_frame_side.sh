# compute-sanitizer over the section-8f kernels (stereo, device-resident frame, distinctive descriptors)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 150 compute-sanitizer --tool $tool --print-limit 5 python tools/frame_side_once.py --small > gpurun_out/sanitize_8f_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^stereo kept|Error|hazard" gpurun_out/sanitize_8f_$tool.log | head -8
done
