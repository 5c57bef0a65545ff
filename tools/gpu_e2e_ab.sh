# end-to-end A/B of the host pipeline's knobs: usage gpu_e2e_ab.sh "VAR=a VAR2=b" ...
cd "${GRAFT_REPO_ROOT:-.}"
for cfg in "$@"; do
env $cfg python bench.py --steps 10 --warmup 3 --no-cpu --no-hamming --no-latency --no-allpairs --no-kitti 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg: value %.0f  e2e %.0f  (%.3f of the link ceiling)'%(d['value'],d['e2e']['value'],d['e2e']['frac_of_link_ceiling']))"
done
