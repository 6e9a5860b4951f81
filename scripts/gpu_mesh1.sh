#!/bin/bash
# mesh latency breakdown, full bench line, launch list of one 512^3 octree reconstruction
mkdir -p gpurun_out
timeout 600 python scripts/mesh_latency.py 512 3 > gpurun_out/mesh_latency.json 2> gpurun_out/mesh_latency.err
tail -3 gpurun_out/mesh_latency.err; cat gpurun_out/mesh_latency.json
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_mesh512.csv python scripts/profile_mesh.py 512 octree > gpurun_out/prof_mesh.log 2>&1
tail -2 gpurun_out/prof_mesh.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_mesh512.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
H=rows[hdr]; data=rows[hdr+1:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value'); mi=H.index('Metric Name')
tot={}
for r in data:
    if r[mi]!='gpu__time_duration.sum': continue
    k=r[ki].split('(')[0].split('::')[-1][:40]; t=tot.setdefault(k,[0,0.0]); t[0]+=1; t[1]+=float(r[vi].replace(',',''))/1000
for k,v in sorted(tot.items(), key=lambda x:-x[1][1])[:16]: print('%-42s %5d %10.1f us' % (k,v[0],v[1]))
PY
