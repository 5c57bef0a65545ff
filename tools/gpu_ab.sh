# A/B of environment-variable tuning knobs: usage gpu_ab.sh "VAR=a" "VAR=b" ...
cd "${GRAFT_REPO_ROOT:-.}"
for cfg in "$@"; do
env $cfg python bench.py --steps 10 --warmup 3 --no-cpu --no-hamming --no-latency --no-allpairs --no-kitti 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg: value %.0f  ms/step %.2f '%(d['value'],d['ms_per_step']), {k: round(v,3) for k,v in d['kernels_ms_per_step'].items()})"
done
