N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -3
for nb in 12 40; do
  timeout 300 $TR scripts/dist_check.py $nb > gpurun_out/dist_check_w${N}_n${nb}.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_w${N}_n${nb}.log
  grep -E "dist_check ok|rc=|Error|error|assert" gpurun_out/dist_check_w${N}_n${nb}.log | tail -4
done
IIFE_KSP_DEBUG=1 AB_REPS=1 timeout 600 $TR scripts/dist_cg_ab.py 184 > gpurun_out/dist_cg_ab_w${N}.log 2>&1; echo "ab rc=$?"
grep -E "^\[w" gpurun_out/dist_cg_ab_w${N}.log | tail -8
grep -c "graph: cached" gpurun_out/dist_cg_ab_w${N}.log; grep -c "graph: capture" gpurun_out/dist_cg_ab_w${N}.log
