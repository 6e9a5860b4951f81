#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
