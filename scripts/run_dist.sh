#!/bin/bash
# Multi-GPU validation on N GPUs of one box ($1, default 2; $2 = quick skips the A/B against the 5-kernel iteration): parity of the row-partitioned path against the single-GPU
# path (scripts/dist_check.py, two sizes; with the three-kernel CG iteration and without), then the bench line with its
# in-run parity check and per-phase times, A/B of the three-kernel iteration.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511"
for nb in 12 40; do
  timeout 300 $TR scripts/dist_check.py $nb > gpurun_out/dist_check_w${N}_n${nb}.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_w${N}_n${nb}.log
  grep -E "dist_check ok|rc=|Error|error|assert" gpurun_out/dist_check_w${N}_n${nb}.log | tail -4
done
if [ "${2:-}" != "quick" ]; then
IIFE_CG_FUSED3=0 timeout 300 $TR scripts/dist_check.py 40 > gpurun_out/dist_check_w${N}_n40_nofused.log 2>&1
grep -E "dist_check ok|Error|error|assert" gpurun_out/dist_check_w${N}_n40_nofused.log | tail -3
fi
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_w${N}.json 2> gpurun_out/bench_w${N}.err; echo "bench rc=$?"
if [ "${2:-}" != "quick" ]; then
IIFE_CG_FUSED3=0 timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_w${N}_nofused.json 2> gpurun_out/bench_w${N}_nofused.err; echo "bench (5-kernel iteration) rc=$?"
fi
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for tag in ("", "_nofused"):
    try:
        d = json.loads(open(f"gpurun_out/bench_w{n}{tag}.json").read().strip().splitlines()[-1])
        c = d["config"]
        print(f"[w{n}{tag}] step {d['ms_per_step']:.2f} ms  phases {c.get('phases_ms')}  parity {c.get('parity_check')} rel {c.get('rel_diff')}  e2e {d.get('e2e') and d['e2e'].get('ms_per_step')}")
    except Exception as exc:
        print(f"[w{n}{tag}] no line: {exc}")
        print(open(f"gpurun_out/bench_w{n}{tag}.err").read()[-2000:])
PY
