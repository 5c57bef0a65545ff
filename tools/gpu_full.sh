cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
} > gpurun_out/full.log 2>&1
tail -40 gpurun_out/full.log
