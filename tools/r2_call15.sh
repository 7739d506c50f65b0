#!/bin/bash
# round 2, call 15 (gpurun --gpus 8): the default transport at N=8 is now the peer-memory kernel — the bench as the driver launches it (fewer steps: budget)
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/c15_bench8.out 2> gpurun_out/c15_bench8.err; echo "rc=$?"
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/c15_bench8.out') if l.startswith('{')][-1])
print({k: b.get(k) for k in ('value','ms_per_step','error','stage')}, 'e2e', (b.get('e2e') or {}).get('ms_per_step'))
if not b.get('error'): print(b['stages']['exchange'], sorted(set(b['stages']['pcg_iterations_per_step'])), b['stages']['l2_criterion'], b['stages']['stale_cuda_errors'])
PY
tail -3 gpurun_out/c15_bench8.err | cut -c1-300
