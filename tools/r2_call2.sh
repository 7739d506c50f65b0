#!/bin/bash
# round 2, call 2 (gpurun --gpus 2): attribute the illegal address seen without the stream rendezvous; does the mid-solve transient need NCCL?
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 1: no align, CUDA_LAUNCH_BLOCKING=1"
CUDA_LAUNCH_BLOCKING=1 TOE_DIST_NO_ALIGN=1 DIAG_ITMAX=3000 timeout 120 $T --master-port 29611 tools/dist_diag.py 260,110,58 2 1 > gpurun_out/c2_blocking.log 2>&1
grep -E "^\[r|Error|error" gpurun_out/c2_blocking.log | head -20
echo "== 2: no align, compute-sanitizer memcheck"
TOE_DIST_NO_ALIGN=1 DIAG_ITMAX=100 timeout 240 $T --master-port 29612 --no-python /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 --log-file gpurun_out/c2_san_%p.log python tools/dist_diag.py 260,110,58 1 1 > gpurun_out/c2_san.out 2>&1
grep -E "^\[r|rror" gpurun_out/c2_san.out | head; for f in gpurun_out/c2_san_*.log; do echo "-- $f"; head -60 $f; done
echo "== 3: peer-memory transport (no NCCL in the loop)"
TOE_DIST_P2P=1 timeout 120 $T --master-port 29613 tools/dist_diag.py 260,110,58 6 2 2>&1 | grep -E "^\[r0|rror" | cut -c1-200 | tee gpurun_out/c2_p2p.log
echo "== 4: all-gather transport"
TOE_DIST_XCHG=allgather timeout 120 $T --master-port 29614 tools/dist_diag.py 260,110,58 6 2 2>&1 | grep -E "^\[r0|rror" | cut -c1-200 | tee gpurun_out/c2_ag.log
