#!/bin/bash
# compute-sanitizer memcheck over the GPU parity suite (everything but the 512^3 and encoder tests); summary to gpurun_out/
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_suite.log \
    python -m pytest tests -m gpu -q --no-header -x --ignore=tests/test_fullsize_gpu.py --ignore=tests/test_encoders_gpu.py 2>&1 | tail -3
echo "rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|Misaligned" gpurun_out/memcheck_suite.log | sort | uniq -c | head
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/synccheck_rows.log \
    python -m pytest tests/test_chain_gpu.py -m gpu -q --no-header -x -k "runlist and all" 2>&1 | tail -2
echo "rc=$?"; grep -E "ERROR SUMMARY|Barrier|divergent" gpurun_out/synccheck_rows.log | sort | uniq -c | head
