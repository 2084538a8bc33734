#!/bin/bash
# After tools/launch_list.sh came back through gpurun_out/: rebuild the committed evidence under profiles/.
cd "$(dirname "$0")/.."
python tools/summarize_launches.py gpurun_out/launches.csv profiles/${ROUND:-r02}_launches_summary.csv profiles/${ROUND:-r02}_traffic.json 4 3 | tail -24
python - <<'PY'
import csv
lines=[ln for ln in open('gpurun_out/launches.csv') if not ln.startswith('==')]
rows=list(csv.DictReader(lines))
ids=sorted({int(r['ID']) for r in rows}); per=len(ids)//4; keep=set(ids[3*per:4*per])
by={}
for r in rows:
    i=int(r['ID'])
    if i not in keep: continue
    e=by.setdefault(i,{'launch':i,'kernel':r['Kernel Name'],'grid':r['Grid Size'],'block':r['Block Size']})
    e[r['Metric Name']]=f"{r['Metric Value']} {r['Metric Unit']}"
import os
with open('profiles/' + os.environ.get('ROUND', 'r02') + '_launches_timed_step.csv','w',newline='') as f:
    w=csv.writer(f); w.writerow(['launch','kernel','grid','block','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum'])
    for i in sorted(by):
        e=by[i]; w.writerow([e['launch'],e['kernel'],e['grid'],e['block'],e.get('gpu__time_duration.sum'),e.get('dram__bytes_read.sum'),e.get('dram__bytes_write.sum')])
print(len(by),'launches in the timed step')
PY
