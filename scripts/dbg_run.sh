mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -8 > gpurun_out/verify_tests.log; tail -3 gpurun_out/verify_tests.log
ROBUST_ONLY_S2=1 timeout 600 python scripts/robustness.py > gpurun_out/robust_s2.md 2> gpurun_out/robust_s2.err; tail -3 gpurun_out/robust_s2.md | cut -c1-400; tail -3 gpurun_out/robust_s2.err
IIFE_PTAP_SLOTS_WIDE=0 ROBUST_ONLY_S2=1 timeout 600 python scripts/robustness.py 184 171 128 2>&1 | tail -1 | cut -c1-300
