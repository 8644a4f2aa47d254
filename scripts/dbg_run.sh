mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -8 > gpurun_out/verify_tests.log; tail -3 gpurun_out/verify_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/dbg_bench.json 2> gpurun_out/dbg_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/dbg_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("step", d["ms_per_step"], "spmv ms", r.get("launch_ms"), "frac", r.get("frac"), "cg", r.get("cg_iteration"), "ptap", r.get("ptap_numeric"))
print("e2e", d["e2e"]["ms_per_step"], d["e2e"].get("fixed_pattern"))
PY
timeout 200 python scripts/cg_slope.py 184 2>&1 | tail -1
