cd "${GRAFT_REPO_ROOT:-.}"
python -m pytest tests/test_search_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/search_latency.py 2>&1 | tail -6
