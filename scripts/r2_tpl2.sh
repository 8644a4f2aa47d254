#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short -x -k "ptap or unfitted or cube or golden or fuzz or spmv or fullsize" 2>&1 | tail -15 > gpurun_out/tpl2_tests.log
tail -3 gpurun_out/tpl2_tests.log
SKIP_FGMRES=1 timeout 200 python scripts/phase_bench.py 184 2>&1 | grep -E "spmv"
timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/tpl2_bench.json 2> gpurun_out/tpl2_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/tpl2_bench.json")); r = d["roofline"]
print(f"step {d['ms_per_step']:.2f} ms  spmv {r['launch_ms']*1e3:.0f} us  cg/it {r['cg_iteration']['ms']*1e3:.0f} us  ptap numeric {r['ptap_numeric']['ms']:.2f} ms  cold {d['config'].get('cold_ptap_symbolic_plus_numeric_ms')}")
PY
TAG=2 bash scripts/r2_prof_tpl.sh
