cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-2}
mkdir -p gpurun_out
python -m pytest tests/test_hamming_gpu.py -m gpu -x -q 2>&1 | tail -2
for cfg in "ORBB_SHARD_OVERLAP=0" "ORBB_SHARD_OVERLAP=1"; do
env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --frames 256 --no-hamming > gpurun_out/ap.json 2> gpurun_out/ap.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ap.json").read().strip().splitlines()[-1])
a=d["allpairs"]
print("$cfg: %.0f Gcmp/s step %.1f ms popc %.3f gather %.1f ms (%.1f GB/s) launches %d check %s" % (a["value"]/1e9, a["ms_per_step"], a["roofline"]["frac"], a["collective"]["gather_ms"], a["collective"]["nvlink_gbs_per_rank"] or 0, a["gpu_launches"], a.get("check")))
PY
done
python bench.py --steps 5 --warmup 3 --frames 256 --no-hamming --no-cpu --no-latency 2>/dev/null | tail -1 | python -c "
import json,sys
a=json.loads(sys.stdin.read())['allpairs']
print('N=1: %.0f Gcmp/s step %.1f ms popc %.3f' % (a['value']/1e9, a['ms_per_step'], a['roofline']['frac']))"
