cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; grep -E "NCCL INFO.*(Channel|NVLS|Connected|comm 0x.*rank)" gpurun_out/bench_n$N.err | head -12; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]))
print({k: v for k, v in d["e2e"].items() if k.startswith("h2d") or k.startswith("frac") or k == "numa"})
print(d["strong_scaling"])
print(json.dumps(d.get("allpairs"), indent=1))
print(d.get("hamming", {}).get("value"))
PY
