# weak-scaling bench at N GPUs of one box: N=${N:-2}
cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read())
print("N=$N value %.0f e2e %.0f ms/step %.2f hamming %.3g" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["hamming"]["value"]))
PY
