set -x
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2>&1 | tail -5
