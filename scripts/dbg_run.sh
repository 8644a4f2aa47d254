mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -3
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/dist_check.py 40 2>&1 | grep -E "dist_check ok|Error|error|assert" | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/dbg_bench.json 2> gpurun_out/dbg_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/dbg_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("step", d["ms_per_step"], "spmv ms", r.get("launch_ms"), "cg", r["cg_iteration"]["ms"], "samples", r["cg_iteration"]["solve_ms_samples"], "ptap", r["ptap_numeric"]["ms"], "its", d["config"].get("cg_iterations"))
PY
