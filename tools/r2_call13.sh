#!/bin/bash
# round 2, call 13 (gpurun --gpus 4): fresh set-ups at N=4 (default transport), and the peer-memory transport at N=4 (timed out once in round 1)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
echo "== soak N=4, default transport: 12 set-ups x 2 solves"
timeout 200 $T --master-port 29731 tools/dist_diag.py 260,110,58 12 1 > gpurun_out/c13_soak4.log 2>&1
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c13_soak4.log | sort | uniq -c | tr '\n' ';')  hist: $(grep -o 'hist [0-9a-f]*' gpurun_out/c13_soak4.log | sort | uniq -c | tr '\n' ';')"
grep -o "solve_s [0-9.]*" gpurun_out/c13_soak4.log | sort | uniq -c | sort -rn | head -2; grep -E "rror" gpurun_out/c13_soak4.log | head -3
echo "== peer-memory transport N=4: 8 set-ups x 2 solves"
TOE_DIST_XCHG=p2p timeout 200 $T --master-port 29732 tools/dist_diag.py 260,110,58 8 1 > gpurun_out/c13_soak4_p2p.log 2>&1
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c13_soak4_p2p.log | sort | uniq -c | tr '\n' ';')  hist: $(grep -o 'hist [0-9a-f]*' gpurun_out/c13_soak4_p2p.log | sort | uniq -c | tr '\n' ';')"
grep -o "solve_s [0-9.]*" gpurun_out/c13_soak4_p2p.log | sort | uniq -c | sort -rn | head -2; grep -E "rror|timed out" gpurun_out/c13_soak4_p2p.log | head -3
echo "== graph capture of the partitioned loop N=4"
TOE_DIST_GRAPH=1 timeout 120 $T --master-port 29733 tools/dist_diag.py 260,110,58 3 1 > gpurun_out/c13_graph4.log 2>&1; echo "rc=$?"
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c13_graph4.log | sort | uniq -c | tr '\n' ';')"; grep -o "solve_s [0-9.]*" gpurun_out/c13_graph4.log | sort | uniq -c | sort -rn | head -2
