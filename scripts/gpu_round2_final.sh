#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_suite.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 400 gpurun_out/r02_bench_final.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_mesh512_final_launches.csv python scripts/profile_mesh.py 512 octree > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_mesh512_final_launches.csv 40 | grep -E "total|classify|emit|scan_|active_rows"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_final.json"))
print("value %.4g e2e %.4g ms %.2f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]), d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "algorithmic_frac", "share_of_step")})
print("parity", d["parity"])
for k in ("fast", "hybrid", "split"):
    print(k, d["precision"][k]["queries_per_s"], d["precision"][k]["parity"]["max_abs_err"], d["precision"][k]["parity"]["sign_agreement"])
m = d["mesh_512"]
for k in ("octree", "dense"):
    print(k, {q: m[k][q] for q in ("latency_ms", "field_ms", "mc_ms", "mesh_d2h_ms", "verts", "faces")}, m[k]["mc_roofline"]["frac"])
print("hybrid", m["octree_hybrid"]["latency_ms"], m["octree_hybrid"]["vs_cpu"])
print("narrow", m["octree_hybrid_narrow_band"]["latency_ms"], m["octree_hybrid_narrow_band"]["vs_cpu"]["fraction_over_1e-3"])
print("tail", m["octree"]["gen_mesh_tail"], m["octree"]["postprocess"])
print("cpu", m["cpu_baseline_ms"], "enc", d["encoders"]["frames_256_octree"]["frames_per_s"])
PY
# marching cubes, dominant kernel: DRAM traffic and issue utilisation of one launch on the bench's octree field
ncu --set full --clock-control none -k regex:classify_warp_kernel -c 1 -o gpurun_out/r02_mc_classify -f python scripts/profile_mesh.py 512 octree > /dev/null 2>&1
ncu -i gpurun_out/r02_mc_classify.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h,u,v=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
for w in want:
    for i,n in enumerate(h):
        if n==w: print(w, v[i], u[i])
" | tee gpurun_out/r02_mc_classify_ncu.txt
rm -f gpurun_out/r02_mc_classify.ncu-rep
