# ncu --set full captures of the extractor kernels (one launch each, after warm-up)
set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming"
for k in ${KERNELS:-fast_cells octree_kernel blur_kernel brief_kernel pyramid_level0}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out
