#!/bin/bash
# Turns the scratch output of tools/gpu_final_profile.sh (gpurun_out/) into the committed summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
python tools/ncu_summary.py r01_final final_fast_cells.ncu-rep final_octree_kernel.ncu-rep final_blur_kernel.ncu-rep \
    final_brief_kernel.ncu-rep final_pyramid_resize.ncu-rep final_bf_scan.ncu-rep > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_final.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 24 : python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming --no-latency (two steady-state steps of 256 frames)" > profiles/r01_final_launches_summary.txt 2>/dev/null
cp gpurun_out/launches_final.csv profiles/r01_final_launches.csv
python - <<'PY'
import csv, io, json, subprocess
out = {}
for key, rep in (("fast", "final_fast_cells"), ("blur", "final_blur_kernel"), ("quadtree", "final_octree_kernel"),
                 ("brief", "final_brief_kernel"), ("pyramid_resize_level1", "final_pyramid_resize")):
    o = subprocess.run(["ncu", "-i", "gpurun_out/%s.ncu-rep" % rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(o)))
    d = dict(zip(r[0], zip(r[1], r[2])))
    def b(k):
        u, v = d[k]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    out[key] = {"dram_bytes_per_frame": round((b("dram__bytes_read.sum") + b("dram__bytes_write.sum")) / 256), "frames_in_capture": 256,
                "source": "gpurun_out/%s.ncu-rep (ncu --set full --clock-control none, one launch of bench.py --frames 256)" % rep}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
d = json.loads(open("gpurun_out/bench_final.json").read().strip().split("\n")[-1])
json.dump(d, open("profiles/r01_bench_final.json", "w"), indent=1)
print("value %.0f  e2e %.0f  ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
print(d["kernels_ms_per_step"]); print(d["clocks"]); print(d["latency_ms"]); print(d["cpu_baseline"].get("matcher"))
print(d["e2e"]); print(d["extract_roofline"]); print(d["hamming"]["value"], d["hamming"]["roofline"]["frac"]); print(d["cpu_baseline"]["value"])
PY
