set -x
cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/allpairs_sharded.py --kf 1024 --desc 1000 --check 2>&1 | tail -1 > gpurun_out/allpairs_n${N}_check.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/allpairs_sharded.py --kf 8192 --desc 1000 2>&1 | tail -1 > gpurun_out/allpairs_n${N}_8192.json
cat gpurun_out/bench_n$N.json | cut -c1-400; cat gpurun_out/allpairs_n${N}_check.json gpurun_out/allpairs_n${N}_8192.json
