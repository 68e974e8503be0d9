#!/bin/bash
# round check with the NLPD 2x kernels: full parity suite (no -x), generic-path loss tests, smoke, bench, NLPD timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
SRK_NLPD_2X=0 timeout 300 python -m pytest tests -m gpu -q -k "losses or nlpd or metrics or sharded_evaluate or srcnn_bf16" > gpurun_out/pytest_nlpd_generic.log 2>&1; echo "pytest(generic NLPD) rc=$?"; tail -2 gpurun_out/pytest_nlpd_generic.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python scratch/prof_nlpd.py 2>&1 | grep NLPD
echo -n "bench 2X=0: "; SRK_NLPD_2X=0 timeout 300 python bench.py --no-cpu-baseline --steps 40 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-420
