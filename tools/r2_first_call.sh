#!/bin/bash
# First GPU call of round 2 (run under gpurun, 1 GPU):  bash tools/r2_first_call.sh
# 1. the full GPU test suite (defaults first, opt-in variants last), 2. every opt-in variant measured next to its default,
# 3. the default bench, 4. ncu launch list + full captures of the new kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2_pytest.log
timeout 300 python tools/variants_probe.py 1M  > gpurun_out/r2_variants_1M.json  2> gpurun_out/r2_variants_1M.err;  tail -c 1500 gpurun_out/r2_variants_1M.json
timeout 600 python tools/variants_probe.py 10M > gpurun_out/r2_variants_10M.json 2> gpurun_out/r2_variants_10M.err; tail -c 3000 gpurun_out/r2_variants_10M.json
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 2500 gpurun_out/r2_bench_n1.json
if [ -s gpurun_out/r2_variants_10M.json ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_asm_rows_tet|k_ebe_pipe|k_tl_restrict|k_tl_gemv|k_tl_z|k_gj_update" -c 12 \
      -o gpurun_out/r2_new_kernels_10M python tools/variants_probe.py 10M > gpurun_out/r2_ncu.log 2>&1
  tail -3 gpurun_out/r2_ncu.log
fi
