cd "${GRAFT_REPO_ROOT:-.}"
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys, json
sys.path.insert(0, "vi-orb-slam-icra2018_b200")
import bench
print(json.dumps(bench.single_frame_latency(0), indent=0))
PY
ORBB_NO_GRAPH=1 python - <<'PY'
import sys, json
sys.path.insert(0, "vi-orb-slam-icra2018_b200")
import bench
d = bench.single_frame_latency(0)
print("no graph:", {k: v for k, v in d.items() if k.startswith("extract_") and not k.endswith("stage_ms")})
PY
