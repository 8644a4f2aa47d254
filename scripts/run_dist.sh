#!/bin/bash
# Multi-GPU validation on N GPUs of one box ($1, default 2): parity of the row-partitioned path against the single-GPU
# path (scripts/dist_check.py, two sizes), the bench line with its in-run parity check and per-phase times, and
# ($2 = trace) the A/B of the CG iteration's variants with the globaltimer split of an iteration (scripts/dist_cg_ab.py).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511"
for nb in 12 40; do
  timeout 300 $TR scripts/dist_check.py $nb > gpurun_out/dist_check_w${N}_n${nb}.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_w${N}_n${nb}.log
  grep -E "dist_check ok|rc=|Error|error|assert" gpurun_out/dist_check_w${N}_n${nb}.log | tail -4
done
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_w${N}.json 2> gpurun_out/bench_w${N}.err; echo "bench rc=$?"
if [ "${2:-}" = "trace" ]; then
  IIFE_CG_TRACE=1 AB_REPS=1 timeout 600 $TR scripts/dist_cg_ab.py 184 > gpurun_out/dist_cg_trace_w${N}.log 2>&1; echo "trace rc=$?"
  grep -E "^\[w" gpurun_out/dist_cg_trace_w${N}.log | tail -6
fi
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_w{n}.json").read().strip().splitlines()[-1])
    c = d["config"]
    print(f"[w{n}] step {d['ms_per_step']:.2f} ms  phases {c.get('phases_ms')}  parity {c.get('parity_check')} rel {c.get('rel_diff')}  e2e {d.get('e2e') and d['e2e'].get('ms_per_step')}")
except Exception as exc:
    print(f"[w{n}] no line: {exc}")
    print(open(f"gpurun_out/bench_w{n}.err").read()[-2000:])
PY
