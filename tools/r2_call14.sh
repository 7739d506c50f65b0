#!/bin/bash
# round 2, call 14 (1 GPU): final sanity of the committed tree — GPU gate (incl. the new unstructured-mesh tests) and a short N=1 bench
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/c14_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/c14_pytest.log | cut -c1-200
echo "== bench N=1 short"
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-two-level > gpurun_out/c14_bench1.out 2> gpurun_out/c14_bench1.err; echo "rc=$?"; python -c "
import json; b=json.loads([l for l in open('gpurun_out/c14_bench1.out') if l.startswith('{')][-1]); print(b['value'], b['ms_per_step'], b['e2e']['ms_per_step'], b['stages']['pcg_iterations_per_step'], b['stages']['stale_cuda_errors'], b['roofline']['frac'])"
