#!/bin/bash
# round 2, call 10 (gpurun --gpus 8): the driver's SCALE commands at N=8 and N=4, config 5 (60M tets, matrix-free) at N=8, send/recv transport at N=8
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    b=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    if b.get('error'): print('ERROR', b['error'], b.get('stage')); raise SystemExit
    print('value %.4g el/s  ms/step %.1f  e2e ms %.1f  pcg_s %.3f  its %s  exchange %s  spmv %.0f GB/s  asm %.3g el/s  l2 %s' % (b['value'], b['ms_per_step'], b['e2e']['ms_per_step'], b['metric_parts']['pcg_seconds_to_1e-8'],
          sorted(set(b['stages']['pcg_iterations_per_step'])), b['stages']['exchange'], b['stages']['spmv_gbs'], b['metric_parts']['elements_assembled_per_s'] or 0, b['stages'].get('l2_criterion')))
except Exception as ex: print('no line', ex)
PY
}
echo "== a: bench N=8 (driver command)"
TOE_BENCH_VERBOSE=1 timeout 400 $T --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c10_bench8.out 2> gpurun_out/c10_bench8.err; echo "rc=$?"; summ gpurun_out/c10_bench8.out
grep "^\[rank 0" gpurun_out/c10_bench8.err | cut -c1-120 | tail -14
echo "== c: bench N=4 (driver command)"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $T --nproc-per-node 4 --master-port 29702 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c10_bench4.out 2> gpurun_out/c10_bench4.err; echo "rc=$?"; summ gpurun_out/c10_bench4.out
echo "== d: config 5 (60M tets, matrix-free) N=8"
TOE_BENCH_VERBOSE=1 timeout 400 $T --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --workload C5_60M --matrix-free --steps 1 --warmup 1 --e2e-steps 1 --no-e2e-warmup --no-l2 > gpurun_out/c10_bench8_c5.out 2> gpurun_out/c10_bench8_c5.err; echo "rc=$?"; summ gpurun_out/c10_bench8_c5.out
grep "^\[rank 0" gpurun_out/c10_bench8_c5.err | cut -c1-140 | tail -6
echo "== b: send/recv transport N=8"
TOE_DIST_XCHG=sendrecv timeout 200 $T --nproc-per-node 8 --master-port 29704 bench.py --gpus 8 --steps 2 --warmup 1 --e2e-steps 1 --no-l2 > gpurun_out/c10_bench8_sr.out 2> gpurun_out/c10_bench8_sr.err; echo "rc=$?"; summ gpurun_out/c10_bench8_sr.out
echo "== e: peer-memory transport N=8"
TOE_DIST_XCHG=p2p timeout 200 $T --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 2 --warmup 1 --e2e-steps 1 --no-l2 > gpurun_out/c10_bench8_p2p.out 2> gpurun_out/c10_bench8_p2p.err; echo "rc=$?"; summ gpurun_out/c10_bench8_p2p.out
