#!/bin/bash
# round 2, call 6 (gpurun --gpus 2): per-iteration scalar trace of the partitioned CG — which quantity deviates first in a faulty solve?
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for tr in sendrecv allgather; do
echo "== trace, transport $tr"
TOE_CG_TRACE=1 TOE_DIST_XCHG=$tr timeout 260 $T --master-port 29641 tools/dist_diag.py 260,110,58 12 2 > gpurun_out/c6_trace_$tr.log 2>&1
echo "solves: $(grep -o 'niter [0-9]* conv [01] brk [01]' gpurun_out/c6_trace_$tr.log | sort | uniq -c | tr '\n' ';')"
grep -E "TRACE|  it |LOCAL|leaves_ref_at [0-9]|rror" gpurun_out/c6_trace_$tr.log | cut -c1-260 | head -40
done
