#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
cat > /tmp/run5.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_secondary as b
L = b._bind()
which = sys.argv[1]
if which == "5": b.run("config5", 524288, 3, 4, 12, 0x5EED0005, L.rafting_wl_mixed_step, elect=True, pool=True, quiet=False)
else: b.run("config3", 262144, 5, 1, 8, 0x5EED0003, L.rafting_wl_vote_step, local_slot=2, quiet=False)
PY
for c in 5 3; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2_launches_cfg$c.csv python /tmp/run5.py $c > $O/r2_launches_cfg$c.log 2>&1
done
python - <<'PY'
import csv, collections
for c in (5,3):
    rows=list(csv.reader(l for l in open("gpurun_out/r2_launches_cfg%d.csv"%c) if l.startswith('"')))
    h=rows[0]; ix={k:i for i,k in enumerate(h)}
    per=collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[1:]:
        try: per[r[ix["Kernel Name"]][:50]][r[ix["Metric Name"]]].append(float(r[ix["Metric Value"]].replace(",","")))
        except Exception: pass
    print("config", c)
    for k,m in per.items():
        d=m.get("gpu__time_duration.sum",[])
        tail=d[len(d)//2:]
        print("  %-50s n=%3d  dur(us, 2nd half mean) %.1f  inst %.3g  dramR %.3g dramW %.3g warps%% %.1f issue%% %.1f thr/inst %.1f"%(k,len(d), sum(tail)/max(len(tail),1)/1e3 if tail and tail[0]>1000 else sum(tail)/max(len(tail),1),
              sum(m.get("smsp__inst_executed.sum",[0])[len(d)//2:])/max(len(tail),1), sum(m.get("dram__bytes_read.sum",[0])[len(d)//2:])/max(len(tail),1), sum(m.get("dram__bytes_write.sum",[0])[len(d)//2:])/max(len(tail),1),
              sum(m.get("sm__warps_active.avg.pct_of_peak_sustained_active",[0])[len(d)//2:])/max(len(tail),1), sum(m.get("smsp__issue_active.avg.pct_of_peak_sustained_active",[0])[len(d)//2:])/max(len(tail),1), sum(m.get("smsp__thread_inst_executed_per_inst_executed.ratio",[0])[len(d)//2:])/max(len(tail),1)))
PY
