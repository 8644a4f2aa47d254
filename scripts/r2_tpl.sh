#!/bin/bash
# Template numeric PtAP: parity subset, then A/B of the bench line against the per-row kernels and a chunk-size sweep.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu --tb=short -x -k "ptap or unfitted or cube or golden or fuzz" 2>&1 | tail -25 > gpurun_out/tpl_tests.log
tail -3 gpurun_out/tpl_tests.log
run_bench() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/tpl_bench_$tag.json 2> gpurun_out/tpl_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/tpl_bench_{tag}.json"))
    r = d["roofline"]
    print(f"[{tag}] step {d['ms_per_step']:.2f} ms  spmv {r['launch_ms']*1e3:.0f} us  cg/it {r['cg_iteration']['ms']*1e3:.0f} us  ptap numeric {r['ptap_numeric']['ms']:.2f} ms  cold {d['config'].get('cold_ptap_symbolic_plus_numeric_ms')}")
except Exception as exc:
    print(f"[{tag}] no bench line: {exc}")
    print(open(f"gpurun_out/tpl_bench_{tag}.err").read()[-1500:])
PY
}
run_bench tpl IIFE_NOP=1
run_bench notpl IIFE_PTAP_TPL=0 IIFE_PTAP_V3=1
run_bench chunk4 IIFE_TPL_CHUNK=4
run_bench chunk64 IIFE_TPL_CHUNK=64
run_bench wpc4 IIFE_TPL_WPC=4
run_bench wpc16 IIFE_TPL_WPC=16
