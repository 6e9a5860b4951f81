#!/bin/bash
# run-list chain: full parity suite, smoke, bench line, launch list of the 512^3 octree reconstruction, ncu full capture of chain_kernel<true>
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cut -c1-300 gpurun_out/bench_full.json
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
    k=r[ki].split('(')[0].split('::')[-1][:40]; t=tot.setdefault(k,[0,0.0,0.0,0.0])
    v=float(r[vi].replace(',',''))
    if r[mi]=='gpu__time_duration.sum': t[0]+=1; t[1]+=v/1000
    elif r[mi]=='dram__bytes_read.sum': t[2]+=v
    elif r[mi]=='dram__bytes_write.sum': t[3]+=v
for k,v in sorted(tot.items(), key=lambda x:-x[1][1])[:24]: print('%-42s %5d %10.1f us  rd %12.0f wr %12.0f' % (k,v[0],v[1],v[2],v[3]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chain_kernel" --launch-skip 16 -c 1 -o gpurun_out/chain_rows_r01 -f python scripts/profile_mesh.py 512 octree > gpurun_out/prof_rows_full.log 2>&1
tail -2 gpurun_out/prof_rows_full.log
