#!/bin/bash
# round 2, call 4 (1 GPU): does a single GPU show transient faults too (history comparison + operator soak)?  GPU gate, ROWS v2, N=1 bench
mkdir -p gpurun_out
echo "== 1: single-GPU soak (graph path, standard CG)"
DIAG_SOAK=30000 timeout 300 python tools/dist_diag.py 260,110,58 5 2 > gpurun_out/c4_soak1.log 2>&1
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c4_soak1.log | sort | uniq -c | tr '\n' ';')  hist: $(grep -o 'hist [0-9a-f]*' gpurun_out/c4_soak1.log | sort | uniq -c | tr '\n' ';')"
grep -E "operator soak|rror|leaves_ref_at [0-9]" gpurun_out/c4_soak1.log | cut -c1-250 | head
echo "== 2: pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c4_pytest.log | cut -c1-250
echo "== 3: assembly variants"
timeout 200 python tools/variants_probe.py 10M --only asm 2>&1 | tail -3 | cut -c1-1500
echo "== 4: bench N=1"
TOE_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/c4_bench1.out 2> gpurun_out/c4_bench1.err; echo "bench rc=$?"
grep "^\[rank 0" gpurun_out/c4_bench1.err | cut -c1-160 | tail -12
tail -c 6000 gpurun_out/c4_bench1.out
