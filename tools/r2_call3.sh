#!/bin/bash
# round 2, call 3 (gpurun --gpus 2): soak of the default (all-gather) transport, NCCL inside graph capture, the N=2 bench as the driver runs it
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 1: soak, default transport"
timeout 200 $T --master-port 29621 tools/dist_diag.py 260,110,58 10 2 > gpurun_out/c3_soak.log 2>&1
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c3_soak.log | sort | uniq -c | tr '\n' ';')  hist: $(grep -o 'hist [0-9a-f]*' gpurun_out/c3_soak.log | sort | uniq -c | tr '\n' ';')"
grep -o "solve_s [0-9.]*" gpurun_out/c3_soak.log | sort | uniq -c | sort -rn | head -3
grep -iE "error|Traceback" gpurun_out/c3_soak.log | head -5
echo "== 2: graph capture of the partitioned loop (TOE_DIST_GRAPH=1)"
TOE_DIST_GRAPH=1 timeout 90 $T --master-port 29622 tools/dist_diag.py 260,110,58 2 2 > gpurun_out/c3_graph.log 2>&1; echo "rc=$?"
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c3_graph.log | sort | uniq -c | tr '\n' ';')"; grep -o "solve_s [0-9.]*" gpurun_out/c3_graph.log | sort | uniq -c | sort -rn | head -3
grep -iE "error|Traceback" gpurun_out/c3_graph.log | head -5
echo "== 3: bench N=2"
TOE_BENCH_VERBOSE=1 timeout 300 $T --master-port 29623 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/c3_bench2.out 2> gpurun_out/c3_bench2.err; echo "bench rc=$?"
grep "^\[rank 0" gpurun_out/c3_bench2.err | cut -c1-160 | head -60
tail -c 3500 gpurun_out/c3_bench2.out
