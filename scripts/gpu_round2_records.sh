#!/bin/bash
# final single-GPU records of round 2: bench line, its ncu launch list, full captures of the dominant kernels
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 600 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mesh > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_bench_launches.csv 12
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -o gpurun_out/r02_chain \
    python scripts/profile_step.py 2424832 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_mesh512_final_launches.csv python scripts/profile_mesh.py 512 octree > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_mesh512_final_launches.csv 40
ncu --set full --clock-control none --import-source on -k regex:classify_warp_kernel -c 1 -o gpurun_out/r02_mc_classify_final python scripts/profile_mc.py 512 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_final.json"))
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "parity", "precision", "group_norm"):
    print(k, json.dumps(d[k])[:900])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "algorithmic_frac", "share_of_step", "avg_launch_us")})
m = d["mesh_512"]
for k in m:
    print(k, json.dumps(m[k])[:1500])
print("reference", open("gpurun_out/r02_bench_reference.json").read()[:600])
PY
