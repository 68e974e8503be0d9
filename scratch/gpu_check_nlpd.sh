#!/bin/bash
# round check with the NLPD 2x kernels: full parity suite (no -x), smoke, bench, NLPD timing on/off
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-420
timeout 100 python scratch/prof_nlpd.py 2>&1 | grep NLPD
