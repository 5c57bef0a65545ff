set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid_resize -s 14 -c 1 -f -o gpurun_out/prof_pyramid_resize $B > gpurun_out/ncu_pr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:octree_kernel -s 2 -c 1 -f -o gpurun_out/prof_octree_kernel $B > gpurun_out/ncu_oc.log 2>&1
ls -la gpurun_out/*.ncu-rep
