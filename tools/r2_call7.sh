#!/bin/bash
# round 2, call 7 (1 GPU): new kernels timed next to their defaults, ncu captures, bench lines of configs 3 and 5 at N=1
mkdir -p gpurun_out
echo "== 1: variants (assembly, matrix-free operator, preconditioner) at 10M"
timeout 400 python tools/variants_probe.py 10M > gpurun_out/c7_variants_10M.json 2> gpurun_out/c7_variants_10M.err; tail -c 3500 gpurun_out/c7_variants_10M.json
echo "== 2: ncu --set full of the kernels that changed (one launch each)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm_rows_tet|k_asm_offdiag|k_asm_diag|k_ebe_tile|k_ebe_nodes|k_tl_restrict|k_tl_gemv|k_tl_z|k_tl_xr|k_tl_p|k_gjb_update|k_tl_coarse_direct" -c 16 \
    -o gpurun_out/r2_kernels_10M python tools/variants_probe.py 10M > gpurun_out/c7_ncu1.log 2>&1; tail -2 gpurun_out/c7_ncu1.log
echo "== 3: ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_10M.csv \
    python bench.py --steps 1 --warmup 1 --no-two-level --no-cpu-baseline --e2e-steps 1 --no-e2e-warmup > gpurun_out/c7_ncu2.log 2>&1; tail -2 gpurun_out/c7_ncu2.log | cut -c1-300
echo "== 4: bench, config 3 (1M tets)"
timeout 300 python bench.py --workload C3_1M --steps 5 --warmup 3 > gpurun_out/c7_bench_c3.out 2> gpurun_out/c7_bench_c3.err; echo "rc=$?"; tail -c 2500 gpurun_out/c7_bench_c3.out
echo "== 5: bench, config 5 (60M tets, matrix-free) at N=1"
TOE_BENCH_VERBOSE=1 timeout 900 python bench.py --workload C5_60M --matrix-free --steps 1 --warmup 1 --e2e-steps 1 --no-e2e-warmup --no-two-level --no-cpu-baseline --deadline 880 > gpurun_out/c7_bench_c5.out 2> gpurun_out/c7_bench_c5.err; echo "rc=$?"
grep "^\[rank 0" gpurun_out/c7_bench_c5.err | cut -c1-160 | tail -8; tail -c 3000 gpurun_out/c7_bench_c5.out
