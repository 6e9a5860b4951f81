#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_chain_gpu.py tests/test_query_gpu.py -m gpu -q --no-header -x 2>&1 | tail -4
timeout 300 python -m pytest tests/test_octree_mc_gpu.py tests/test_gen_mesh_flow_gpu.py -m gpu -q --no-header -x 2>&1 | tail -4
PIFU_CHAIN_TRACE=2 PIFU_CHAIN_TRACE_MIN_TILES=2000 timeout 200 python scripts/recon_phases.py 512 2>&1 | grep -E "chain trace tile [3-4]|rep 3|step 8" | tail -5
timeout 300 python bench.py --no-mesh --no-cpu-baseline --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dense value %.4g ms %.2f clocks %s launches %d' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches']))"
