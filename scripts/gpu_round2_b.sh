#!/bin/bash
# launch list + full capture of the marching-cubes kernels at 512^3 (octree field)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_mesh512_v1_launches.csv python scripts/profile_mesh.py 512 octree > gpurun_out/r02_mesh512_v1.log 2>&1
tail -3 gpurun_out/r02_mesh512_v1.log
python scripts/launch_summary.py gpurun_out/r02_mesh512_v1_launches.csv 30
ncu --set full --clock-control none --import-source on -k regex:classify_kernel -c 1 -o gpurun_out/r02_mc_classify python scripts/profile_mesh.py 512 octree > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
