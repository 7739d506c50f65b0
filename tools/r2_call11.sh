#!/bin/bash
# round 2, call 11 (gpurun --gpus 2): multi-GPU parity tests (2-GPU subset) + the driver's N=2 bench command
mkdir -p gpurun_out
echo "== multi-GPU pytest"
timeout 900 python -m pytest tests/test_gpu_y_dist.py tests/test_gpu_zzz_dist_sendrecv.py -m gpu -q > gpurun_out/c11_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/c11_pytest.log | cut -c1-200
echo "== bench N=2 (driver command)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c11_bench2.out 2> gpurun_out/c11_bench2.err; echo "rc=$?"
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/c11_bench2.out') if l.startswith('{')][-1])
print({k: b.get(k) for k in ('value','ms_per_step','error')}, b['e2e']['ms_per_step'], b['e2e'].get('host_wall_ms_per_call'), sorted(set(b['stages']['pcg_iterations_per_step'])), b['stages']['l2_criterion'])
PY
