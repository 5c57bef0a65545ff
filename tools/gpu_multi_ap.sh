set -x
cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/allpairs_sharded.py --kf ${KF:-512} --desc 1000 --check 2>&1 | tail -2
