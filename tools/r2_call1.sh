#!/bin/bash
# round 2, call 1 (gpurun --gpus 2): partitioned-solve diagnostic, current defaults vs round-1 conditions, then the N=2 bench as the driver runs it
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== A: defaults (align on, no warm-up burst)"
timeout 150 $T --master-port 29601 tools/dist_diag.py 260,110,58 4 2 2>&1 | grep "^\[r" | tee gpurun_out/c1_diag_default.log
echo "== B: no align (round-1 conditions)"
TOE_DIST_NO_ALIGN=1 timeout 150 $T --master-port 29602 tools/dist_diag.py 260,110,58 5 2 2>&1 | grep "^\[r" | tee gpurun_out/c1_diag_noalign.log
echo "== C: bench N=2"
TOE_BENCH_VERBOSE=1 timeout 240 $T --master-port 29603 bench.py --gpus 2 --steps 2 --warmup 1 --no-transport-probes > gpurun_out/c1_bench2.out 2> gpurun_out/c1_bench2.err; echo "bench rc=$?"
grep -v "^W1018\|^\*\*\*\|OMP_NUM" gpurun_out/c1_bench2.err | tail -40
tail -c 3000 gpurun_out/c1_bench2.out
