#!/bin/bash
# round 2, call 5 (gpurun --gpus 2): where does the N=2 transient come from?  bit-compare soaks: local product / exchange / both
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 1: soaks, default transport (all-gather)"
DIAG_SOAK=150000 timeout 200 $T --master-port 29631 tools/dist_diag.py 260,110,58 1 1 > gpurun_out/c5_soak_ag.log 2>&1
grep -E "operator soak|rror" gpurun_out/c5_soak_ag.log | cut -c1-220
echo "== 2: soaks, send/recv transport"
TOE_DIST_XCHG=sendrecv DIAG_SOAK=150000 timeout 200 $T --master-port 29632 tools/dist_diag.py 260,110,58 1 1 > gpurun_out/c5_soak_sr.log 2>&1
grep -E "operator soak|rror" gpurun_out/c5_soak_sr.log | cut -c1-220
echo "== 3: soaks, peer-memory transport"
TOE_DIST_XCHG=p2p DIAG_SOAK=150000 timeout 200 $T --master-port 29633 tools/dist_diag.py 260,110,58 1 1 > gpurun_out/c5_soak_p2p.log 2>&1
grep -E "operator soak|rror" gpurun_out/c5_soak_p2p.log | cut -c1-220
