#!/bin/bash
# A/B: weight ring depth of the lattice chain kernel (5 vs 3 slots)
mkdir -p gpurun_out
timeout 300 python bench.py --no-mesh --no-cpu-baseline --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NW=5 value %.4g ms %.2f clocks %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))"
PIFU_NVCC_FLAGS="-DCHAIN_NW=3 -DCHAIN_NWR=1" timeout 600 python -c "
from pifu_b200 import build
build.build(force=True)"
timeout 300 python bench.py --no-mesh --no-cpu-baseline --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NW=3 value %.4g ms %.2f clocks %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))"
PIFU_NVCC_FLAGS="-DCHAIN_NW=4 -DCHAIN_NWR=2" timeout 600 python -c "
from pifu_b200 import build
build.build(force=True)"
timeout 300 python bench.py --no-mesh --no-cpu-baseline --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NW=4 value %.4g ms %.2f clocks %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))"
