set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('value %.0f e2e %.0f ms/step %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
print(d['kernels_ms_per_step']); print(d['clocks']); print(d['roofline']); print(d.get('latency_ms')); print(d.get('cpu_baseline')); print(d['hamming']['value'], d['hamming']['roofline']['frac'])
PY
