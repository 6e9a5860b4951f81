#!/bin/bash
# quick iteration: parity of the query path, bench, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_query_gpu.py -q --no-header -x 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_iter.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g q/s  ms/step %.2f  e2e %.4g  gemm TF %.1f frac %.3f share %.3f launches %d clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'], d['gpu_launches'], d['clocks']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_iter.csv python scripts/profile_step.py 75776 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_iter.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
H=rows[hdr]; data=rows[hdr+1:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value')
sel=[(r[ki][:50], float(r[vi])/1000) for r in data if 'gemm_tc' in r[ki] or 'gather' in r[ki]]
print(' | '.join('%s %.1f' % (k.split('::')[-1][:14], v) for k, v in sel[-7:]))
PY
