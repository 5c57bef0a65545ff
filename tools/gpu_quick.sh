set -x
cd "${GRAFT_REPO_ROOT:-.}"
python -m pytest tests/test_extractor_gpu.py tests/test_golden_gpu.py tests/test_adapter_gpu.py -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu --no-hamming 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f f/s  e2e %.0f f/s  ms/step %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
print(d['kernels_ms_per_step'])
print(d['roofline'])
"
