cd "${GRAFT_REPO_ROOT:-.}"
python -m pytest tests/test_search_gpu.py tests/test_hamming_gpu.py tests/test_adapter_gpu.py -m gpu -x -q 2>&1 | tail -15
