set -x
python -m pytest tests/test_search_gpu.py -m gpu -x -q 2>&1 | tail -40
