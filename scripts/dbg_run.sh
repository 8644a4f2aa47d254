mkdir -p gpurun_out
IIFE_PLAN_DEBUG=1 timeout 300 python scripts/configs_1_4.py > gpurun_out/configs_dbg.md 2> gpurun_out/configs_dbg.err
grep "cold calls" gpurun_out/configs_dbg.err
tail -7 gpurun_out/configs_dbg.md | cut -c1-200
