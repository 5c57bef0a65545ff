# round-2 multi-GPU checkpoint (N GPUs of one box): bench.py with the all-pairs leg, distinctive_sharded
cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench exit $?"; tail -3 gpurun_out/r02_bench_n$N.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/distinctive_sharded_run.py 200000 2>&1 | tail -1 | tee gpurun_out/r02_distinctive_sharded_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]))
print({k: v for k, v in d["e2e"].items() if k.startswith("h2d") or k.startswith("frac") or k.startswith("frames")})
print(d.get("strong_scaling"))
a=d.get("allpairs") or {}
print({k: a.get(k) for k in ("value","ms_per_step","check")}, a.get("roofline",{}).get("frac"), {k:v for k,v in a.get("collective",{}).items() if k!="nccl_log"})
print(d.get("hamming", {}).get("value"))
PY
