#!/bin/bash
# round 2, call 8 (gpurun --gpus 2): soak of the fixed SpMV pipeline protocol — both NCCL transports, then the N=2 bench
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for tr in sendrecv allgather; do
echo "== soak, transport $tr"
TOE_DIST_XCHG=$tr timeout 330 $T --master-port 29651 tools/dist_diag.py 260,110,58 16 2 > gpurun_out/c8_soak_$tr.log 2>&1
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c8_soak_$tr.log | sort | uniq -c | tr '\n' ';')  hist: $(grep -o 'hist [0-9a-f]*' gpurun_out/c8_soak_$tr.log | sort | uniq -c | tr '\n' ';')"
grep -o "solve_s [0-9.]*" gpurun_out/c8_soak_$tr.log | sort | uniq -c | sort -rn | head -2
grep -E "rror" gpurun_out/c8_soak_$tr.log | head -3
done
echo "== no-align run (the illegal address of call 1)"
TOE_DIST_NO_ALIGN=1 TOE_DIST_XCHG=sendrecv timeout 120 $T --master-port 29652 tools/dist_diag.py 260,110,58 4 1 > gpurun_out/c8_noalign.log 2>&1
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c8_noalign.log | sort | uniq -c | tr '\n' ';')"; grep -E "rror" gpurun_out/c8_noalign.log | head -3
