set -x
cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/allpairs_sharded.py --kf 512 --desc 1000 --check 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1
