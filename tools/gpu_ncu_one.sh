# usage: gpu_ncu_one.sh <kernel-regex> <out-name> [skip]   -- one `ncu --set full` capture of a 256-frame launch
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming --no-latency"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${3:-2} -c 1 -f -o gpurun_out/$2 $B > gpurun_out/ncu_$2.log 2>&1
ls -la gpurun_out/$2.ncu-rep
