#!/bin/bash
# Turns the scratch output of tools/gpu_r02_final_profile.sh (gpurun_out/) into the committed summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
python tools/ncu_summary.py r02_final r02_final_fast_warp.ncu-rep r02_final_octree_kernel.ncu-rep r02_final_brief_staged.ncu-rep \
    r02_final_blur_staged.ncu-rep r02_final_pyramid_resize3.ncu-rep r02_final_pyramid_level0.ncu-rep r02_final_bf_scan.ncu-rep > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_final.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 57 -c 38 : python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-hamming --no-latency --no-allpairs --no-kitti (the two timed steps of 256 frames, 19 launches each)" > profiles/r02_final_launches_summary.txt
cp gpurun_out/r02_launches_final.csv profiles/r02_final_launches.csv
python - <<'PY'
import csv, json, collections
# whole-stage DRAM traffic of one 256-frame step: dram__bytes_read.sum + dram__bytes_write.sum summed over the stage's launches
rows = [r for r in csv.reader(open("gpurun_out/r02_traffic_final.csv")) if len(r) > 10 and r[0].isdigit()]
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage_of = lambda name: ("pyramid" if "pyramid" in name else "fast" if "fast_warp" in name else "quadtree" if "octree" in name
                         else "blur" if "blur" in name else "brief" if "brief" in name else None)
acc = collections.defaultdict(float)
launches = collections.Counter()
for r in rows:
    st = stage_of(r[4])
    if st is None: continue
    if r[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        acc[st] += float(r[-1]) * unit[r[-2]]
    if r[-3] == "gpu__time_duration.sum":
        launches[st] += 1
out = {st: {"dram_bytes_per_frame": round(acc[st] / 256), "launches_per_step": launches[st], "frames_in_capture": 256,
            "source": "gpurun_out/r02_traffic_final.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                      "the 19 launches of one 256-frame step of bench.py --frames 256)"} for st in acc}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
d = json.loads(open("gpurun_out/r02_bench_final.json").read().strip().split("\n")[-1])
json.dump(d, open("profiles/r02_bench_final.json", "w"), indent=1)
r = json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().split("\n")[-1])
json.dump(r, open("profiles/r02_bench_reference.json", "w"), indent=1)
print("value %.0f  e2e %.0f  ms/step %.2f   reference arm %.1f %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["value"], r["unit"]))
print(d["kernels_ms_per_step"]); print(d["clocks"]); print(d["latency_ms"]); print(d["cpu_baseline"].get("matcher"))
print(d["e2e"]); print(d["extract_roofline"]); print(d["hamming"]["value"], d["hamming"]["roofline"]["frac"]); print(d["cpu_baseline"]["value"])
PY
