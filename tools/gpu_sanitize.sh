set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
