#!/bin/bash
# round 2, call 12 (1 GPU): rehearsal of the driver's single-GPU commands
mkdir -p gpurun_out
echo "== build() + smoke()"
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/c12_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/c12_pytest.log | cut -c1-200
echo "== reference arm"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/c12_ref.out 2> gpurun_out/c12_ref.err; echo "rc=$?"; python -c "
import json; b=json.loads([l for l in open('gpurun_out/c12_ref.out') if l.startswith('{')][-1]); print(b['value'], b['ms_per_step'], b['metric_parts']); print(b['cpu_baseline']['sample'])"
echo "== bench N=1 (driver command)"
TOE_BENCH_VERBOSE=1 timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c12_bench1.out 2> gpurun_out/c12_bench1.err; echo "rc=$?"
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/c12_bench1.out') if l.startswith('{')][-1])
print({k: b.get(k) for k in ('value','ms_per_step','error')}, 'e2e', b['e2e']['ms_per_step'], b['e2e'].get('host_wall_ms_per_call'))
print(sorted(set(b['stages']['pcg_iterations_per_step'])), b['metric_parts'], b['roofline']['frac'], b['stages']['l2_criterion'], b['clocks'])
print(b['cpu_baseline']['value'], b['stages']['two_level_preconditioner'])
PY
grep "^\[rank 0" gpurun_out/c12_bench1.err | tail -6 | cut -c1-120
