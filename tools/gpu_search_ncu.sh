cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/search_launches.csv python tools/search_latency.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/search_launches.csv')) if len(r)>10 and r[0].isdigit()]
# print the last iteration's kernels (kitti, it=2): last ~14 launches
for r in rows[-16:]: print(r[4][:60], r[-1])
PY
