N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511"
echo "--- single GPU, PDL off/on"
timeout 200 python scripts/cg_slope.py 184 2>&1 | tail -2
IIFE_CG_PDL=1 timeout 200 python scripts/cg_slope.py 184 2>&1 | tail -2
IIFE_CG_PDL=1 timeout 300 python -m pytest tests -q -m gpu -x -k "ksp or cg or solve" 2>&1 | tail -3
echo "--- $N GPUs"
IIFE_CG_PDL=1 timeout 300 $TR scripts/dist_check.py 40 > gpurun_out/dist_check_w${N}_n40_pdl.log 2>&1; echo "rc=$?"
grep -E "dist_check ok|Error|error|assert" gpurun_out/dist_check_w${N}_n40_pdl.log | tail -4
AB_EXTRA="IIFE_CG_PDL=1;IIFE_KSP_CHUNK=64" timeout 600 $TR scripts/dist_cg_ab.py 184 > gpurun_out/dist_cg_ab_w${N}.log 2>&1; echo "ab rc=$?"
grep "^\[w" gpurun_out/dist_cg_ab_w${N}.log | tail -6 || tail -30 gpurun_out/dist_cg_ab_w${N}.log
