#!/bin/bash
# final single-GPU numbers: default bench line (e2e + CPU sample) and the CPU reference arm at the headline size (host cores
# of the GPU box); the configs 1-4 table, the phase table and the robustness table come from scripts/capture_profiles.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6 > gpurun_out/final_tests.log; tail -2 gpurun_out/final_tests.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(f"step {d['ms_per_step']:.2f} ms value {d['value']:.1f}  e2e {d['e2e']['ms_per_step']:.1f} ms ({d['e2e']['value']:.1f})  fixed {d['e2e'].get('fixed_pattern')}  cpu {d['cpu_baseline']}")
print(f"spmv {r['launch_ms']*1e3:.0f} us frac {r['frac']:.3f}  cg/it {r['cg_iteration']['ms']*1e3:.0f} us  ptap {r['ptap_numeric']['ms']:.2f} ms frac {r['ptap_numeric']['frac']:.3f}  cold {d['config']['cold_ptap_symbolic_plus_numeric_ms']:.0f}/{d['config']['cold_repeat_ms']:.0f} ms launches {d['gpu_launches']}")
try:
    q = json.loads(open("gpurun_out/final_reference.json").read().strip().splitlines()[-1])
    print("reference:", q["value"], q["ms_per_step"], q["cpu_baseline"])
except Exception as e:
    print("reference failed", e, open("gpurun_out/final_reference.err").read()[-500:])
PY
