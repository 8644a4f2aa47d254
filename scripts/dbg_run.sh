mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 300 python scripts/configs_1_4.py > gpurun_out/configs_1_4.md 2> gpurun_out/configs_1_4.err; echo "configs rc=$?"
grep "cold calls" gpurun_out/configs_1_4.err
tail -7 gpurun_out/configs_1_4.md | cut -c1-220
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(f"step {d['ms_per_step']:.2f} ms value {d['value']:.1f}  e2e {d['e2e']['ms_per_step']:.1f} ms ({d['e2e']['value']:.1f})  fixed {d['e2e'].get('fixed_pattern')}")
print(f"spmv {r['launch_ms']*1e3:.0f} us frac {r['frac']:.3f} traffic {r['traffic']}  cg/it {r['cg_iteration']['ms']*1e3:.0f} us  ptap {r['ptap_numeric']['ms']:.2f} ms frac {r['ptap_numeric']['frac']:.3f}  cold {d['config']['cold_ptap_symbolic_plus_numeric_ms']:.0f}/{d['config']['cold_repeat_ms']:.0f} ms launches {d['gpu_launches']}")
PY
