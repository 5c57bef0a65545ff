cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -40
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-hamming --no-latency > gpurun_out/q.json 2> gpurun_out/q.err
tail -5 gpurun_out/q.err
tail -1 gpurun_out/q.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f f/s  e2e %.0f f/s  ms/step %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
print(d['kernels_ms_per_step'])
print(d['roofline'])
"
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/extract_once.py 2>&1 | grep -E "RACECHECK SUMMARY|^ok|Race reported" | head -5
} > gpurun_out/quick.log 2>&1
tail -60 gpurun_out/quick.log
