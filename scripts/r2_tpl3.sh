#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short -x -k "ptap or unfitted or cube or golden or fuzz or fullsize" 2>&1 | tail -15 > gpurun_out/tpl3_tests.log
tail -3 gpurun_out/tpl3_tests.log
run_bench() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/tpl3_bench_$tag.json 2> gpurun_out/tpl3_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/tpl3_bench_{tag}.json")); r = d["roofline"]
    print(f"[{tag}] step {d['ms_per_step']:.2f} ms  cg/it {r['cg_iteration']['ms']*1e3:.0f} us  ptap numeric {r['ptap_numeric']['ms']:.2f} ms  cold {d['config'].get('cold_ptap_symbolic_plus_numeric_ms'):.0f}")
except Exception as exc:
    print(f"[{tag}] no bench line: {exc}"); print(open(f"gpurun_out/tpl3_bench_{tag}.err").read()[-800:])
PY
}
L=$PWD/interpolation-based-immersed-fea_b200/lib
run_bench mb4 IIFE_NOP=1
run_bench mb3 IIFE_LIB=$L/tpl3/libiife.so
run_bench mb5 IIFE_LIB=$L/tpl5/libiife.so
run_bench mb4_chunk4 IIFE_TPL_CHUNK=4
run_bench mb4_chunk32 IIFE_TPL_CHUNK=32
run_bench mb3_chunk32 IIFE_LIB=$L/tpl3/libiife.so IIFE_TPL_CHUNK=32
TAG=3 bash scripts/r2_prof_tpl.sh
timeout 900 python scripts/robustness.py > gpurun_out/robustness.md 2> gpurun_out/robustness.err; echo "robustness rc=$?"; tail -12 gpurun_out/robustness.md; tail -3 gpurun_out/robustness.err
