#!/bin/bash
# round 2, call 9 (1 GPU): ROWS assembly after the ncu-guided changes, ncu of it and of the two-level kernels, short N=1 bench (l2 extra)
mkdir -p gpurun_out
echo "== 1: assembly variants at 10M"
timeout 200 python tools/variants_probe.py 10M --only asm > gpurun_out/c9_asm.json 2>gpurun_out/c9_asm.err; python -c "
import json; d=json.load(open('gpurun_out/c9_asm.json'))['assembly']; print({k:(v['ms_min'],v['max_rel_diff_Kx_vs_gather']) for k,v in d.items()})"
echo "== 2: ncu rows + gather"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_asm_rows_tet|k_asm_offdiag|k_asm_diag" -c 6 -o gpurun_out/r2_asm_10M python tools/variants_probe.py 10M --only asm > gpurun_out/c9_ncu1.log 2>&1; tail -1 gpurun_out/c9_ncu1.log
echo "== 3: ncu two-level kernels"
timeout 300 ncu --set full --clock-control none -k regex:"k_tl_|k_gjb_|k_gj_" -c 24 -o gpurun_out/r2_twolevel_10M python tools/variants_probe.py 10M --only pc > gpurun_out/c9_ncu2.log 2>&1; tail -1 gpurun_out/c9_ncu2.log
echo "== 4: bench N=1 short"
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c9_bench1.out 2> gpurun_out/c9_bench1.err; echo "rc=$?"; python -c "
import json; b=json.loads([l for l in open('gpurun_out/c9_bench1.out') if l.startswith('{')][-1]); print(b['value'], b['ms_per_step'], b['e2e']['ms_per_step'], b['metric_parts'], b['stages']['l2_criterion'], b['stages']['two_level_preconditioner'])"
