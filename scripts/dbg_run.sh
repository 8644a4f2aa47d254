N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511"
for nb in 12 40; do
  timeout 300 $TR scripts/dist_check.py $nb > gpurun_out/dist_check_w${N}_n${nb}.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_w${N}_n${nb}.log
  grep -E "dist_check ok|rc=|Error|error|assert" gpurun_out/dist_check_w${N}_n${nb}.log | tail -4
done
IIFE_CG_TRACE=1 AB_REPS=1 timeout 600 $TR scripts/dist_cg_ab.py 184 > gpurun_out/dist_cg_trace_w${N}.log 2>&1; echo "trace rc=$?"
grep -E "^\[w" gpurun_out/dist_cg_trace_w${N}.log | tail -40
grep -E "cg trace rank 0" gpurun_out/dist_cg_trace_w${N}.log | awk 'NR%4==0' | tail -4
