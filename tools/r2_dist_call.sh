#!/bin/bash
# Second GPU call of round 2 (gpurun --gpus 2):  bash tools/r2_dist_call.sh
# The open item of DESIGN.md §6: does the first partitioned solve after a set-up still break down at 10M tets?  A/B: with / without the
# stream rendezvous (dist_align), NCCL / peer-memory transport; `probe` fingerprints diag, f and K·x before each solve.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29601 tools/dist_resetup_check.py 260,110,58 4 2>&1 | grep "^rep" | tee gpurun_out/r2_resetup_align.log
TOE_DIST_NO_ALIGN=1 timeout 200 $T --master-port 29602 tools/dist_resetup_check.py 260,110,58 4 2>&1 | grep "^rep" | tee gpurun_out/r2_resetup_noalign.log
TOE_DIST_NO_ALIGN=1 timeout 200 $T --master-port 29603 tools/dist_resetup_check.py 260,110,58 4 probe 2>&1 | grep "^rep" | tee gpurun_out/r2_resetup_noalign_probe.log
TOE_DIST_XCHG=allgather timeout 200 $T --master-port 29606 tools/dist_resetup_check.py 260,110,58 4 2>&1 | grep "^rep" | tee gpurun_out/r2_resetup_allgather.log
TOE_DIST_P2P=1 timeout 200 $T --master-port 29604 tools/dist_resetup_check.py 260,110,58 4 2>&1 | grep "^rep" | tee gpurun_out/r2_resetup_p2p.log
timeout 300 $T --master-port 29605 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 1500 gpurun_out/r2_bench_n2.json
TOE_DIST_XCHG=allgather timeout 300 $T --master-port 29607 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2_allgather.json 2> gpurun_out/r2_bench_n2_allgather.err; tail -c 1500 gpurun_out/r2_bench_n2_allgather.json
