#!/bin/bash
# First GPU call of the next round: measure the prepared-but-unmeasured experiments (ROUND_NOTES.md) against
# the default path in ONE call.  Everything is wrapped in `timeout`; results land in gpurun_out/r2_*.
# usage (from the repo root):  gpurun --timeout 1200 -- 'bash scripts/r2_experiments.sh'   (about 10 minutes)
mkdir -p gpurun_out
SUB='ptap or unfitted or cube or golden'
run_bench() {  # $1 = tag, rest = env assignments
  tag=$1; shift
  env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2_bench_{tag}.json"))
    r = d["roofline"]
    print(f"[{tag}] step {d['ms_per_step']:.2f} ms  spmv {r['launch_ms']*1e3:.0f} us  cg/it {r['cg_iteration']['ms']*1e3:.0f} us  ptap numeric {r['ptap_numeric']['ms']:.2f} ms")
except Exception as exc:
    print(f"[{tag}] no bench line: {exc}")
PY
}
# 0. A/B of every experimental kernel against the default one in one process (says which variant is wrong and where)
timeout 240 python scripts/compare_variants.py > gpurun_out/r2_compare.log 2>&1
cat gpurun_out/r2_compare.log
run_bench default IIFE_NOP=1
# 1. slot-plan kernel v2
IIFE_PTAP_V2=1 timeout 120 python -m pytest tests -q -m gpu --tb=line -k "$SUB" 2>&1 | tail -5 > gpurun_out/r2_tests_v2.log
tail -1 gpurun_out/r2_tests_v2.log
run_bench v2 IIFE_PTAP_V2=1
# 1a. stage 2 as a gather program
IIFE_PTAP_V3=1 timeout 120 python -m pytest tests -q -m gpu --tb=line -k "$SUB" 2>&1 | tail -5 > gpurun_out/r2_tests_v3.log
tail -1 gpurun_out/r2_tests_v3.log
run_bench v3 IIFE_PTAP_V3=1
# 1b. register-budget variants (scripts/build_variants.sh builds lib/tuned/libiife.so with 64 registers for the SELL and
#     slot-plan kernels; the .so travels with the snapshot)
TUNED=$PWD/interpolation-based-immersed-fea_b200/lib/tuned/libiife.so
if [ -f "$TUNED" ]; then
  IIFE_LIB=$TUNED timeout 150 python -m pytest tests -q -m gpu --tb=line -k "ptap or spmv or ksp or cube or golden or unfitted" 2>&1 | tail -5 > gpurun_out/r2_tests_tuned.log
  tail -1 gpurun_out/r2_tests_tuned.log
  run_bench tuned IIFE_LIB=$TUNED
  run_bench tuned_fuseddot IIFE_LIB=$TUNED IIFE_SELL_FUSED_DOT=1
  run_bench tuned_v2 IIFE_LIB=$TUNED IIFE_PTAP_V2=1
  run_bench tuned_v3 IIFE_LIB=$TUNED IIFE_PTAP_V3=1
  run_bench tuned_u8 IIFE_LIB=$TUNED IIFE_SELL_UNROLL=8
fi
# 2. SELL path for transposed products
IIFE_SPMV_SELL_T=1 timeout 120 python -m pytest tests -q -m gpu --tb=line -k "spmv or cube or golden or end_to_end" 2>&1 | tail -5 > gpurun_out/r2_tests_sellt.log
tail -1 gpurun_out/r2_tests_sellt.log
IIFE_SPMV_SELL_T=1 SKIP_FGMRES=1 timeout 120 python scripts/phase_bench.py 184 2>&1 | grep -E "spmvT|spmv\(M\)" > gpurun_out/r2_phase_sellt.log
cat gpurun_out/r2_phase_sellt.log
IIFE_SPMV_ILP=1 SKIP_FGMRES=1 timeout 120 python scripts/phase_bench.py 184 2>&1 | grep -E "spmv\(M\)" > gpurun_out/r2_phase_ilp.log
cat gpurun_out/r2_phase_ilp.log
# 3. persistent cooperative CG, single GPU (never leave a hung kernel behind: short timeouts)
IIFE_KSP_PERSIST=1 timeout 120 python -m pytest tests -q -m gpu --tb=line -k "ksp or golden or cube or unfitted or end_to_end" 2>&1 | tail -5 > gpurun_out/r2_tests_persist.log
tail -1 gpurun_out/r2_tests_persist.log
run_bench persist IIFE_KSP_PERSIST=1
IIFE_KSP_PERSIST=1 timeout 120 python scripts/small_configs.py > gpurun_out/r2_small_persist.log 2>&1
timeout 120 python scripts/small_configs.py > gpurun_out/r2_small_default.log 2>&1
tail -6 gpurun_out/r2_small_default.log gpurun_out/r2_small_persist.log
# 3b. row N4 (never run on a GPU so far)
IIFE_TEST_UNVERIFIED=1 timeout 90 python -m pytest tests -q -m gpu --tb=short -k condition_estimate 2>&1 | tail -15 > gpurun_out/r2_tests_n4.log
tail -1 gpurun_out/r2_tests_n4.log
# 3c. randomised parity sweep (hypothesis), also never run on a GPU so far
IIFE_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests/test_gpu_fuzz.py -q -x --tb=short 2>&1 | tail -25 > gpurun_out/r2_tests_fuzz.log
tail -1 gpurun_out/r2_tests_fuzz.log
# 4. S2 stress case at a size where the hashing kernels matter
timeout 200 python scripts/phase_bench_s2.py 96:1 64:2 > gpurun_out/r2_phase_s2.log 2>&1
cat gpurun_out/r2_phase_s2.log
