# A/B of library variants built by tools/build_variant.py: usage gpu_variants.sh <name> [<name> ...] ("base" = the in-tree library)
cd "${GRAFT_REPO_ROOT:-.}"
P=vi-orb-slam-icra2018_b200
cp $P/liborbb200.so /tmp/liborbb200_base.so
for v in "$@"; do
if [ "$v" = base ]; then cp /tmp/liborbb200_base.so $P/liborbb200.so; else cp $P/variants/liborbb200_$v.so $P/liborbb200.so; fi
python bench.py --steps 10 --warmup 3 --no-cpu --no-hamming --no-latency --no-allpairs --no-kitti 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v: value %.0f  ms/step %.2f '%(d['value'],d['ms_per_step']), {k: round(v,3) for k,v in d['kernels_ms_per_step'].items()})"
done
cp /tmp/liborbb200_base.so $P/liborbb200.so
