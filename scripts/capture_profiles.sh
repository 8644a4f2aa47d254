#!/bin/bash
# Runs on the GPU box (under gpurun): everything profiles/ is built from, in one call.
#   1. the bench command without a profiler (must exit 0 first), 2. its ncu launch list,
#   3. ncu --set full captures of the top kernels (SELL SpMV with fused dot, template numeric PtAP, CG update),
#   4. the configs 1-4 table and the robustness table.
# Then, in the build container:  python scripts/summarize_profiles.py r02
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/prof_plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/prof_plain_bench.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/prof_ncu_list.log 2>&1
tail -1 gpurun_out/prof_ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 60 -c 2 -f -o gpurun_out/prof_spmv_sell $CMD > gpurun_out/prof_ncu_sell.log 2>&1
tail -1 gpurun_out/prof_ncu_sell.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ptap_numeric_tpl" -s 2 -c 1 -f -o gpurun_out/prof_ptap_numeric $CMD > gpurun_out/prof_ncu_ptap.log 2>&1
tail -1 gpurun_out/prof_ncu_ptap.log
timeout 600 ncu --set full --clock-control none -k regex:"k_cg_update|k_cg_p" -s 60 -c 2 -f -o gpurun_out/prof_cg_vec $CMD > gpurun_out/prof_ncu_cgvec.log 2>&1
tail -1 gpurun_out/prof_ncu_cgvec.log
timeout 600 ncu --set full --clock-control none -k regex:"k_ptap_symbolic" -c 2 -f -o gpurun_out/prof_ptap_symbolic $CMD > gpurun_out/prof_ncu_sym.log 2>&1
tail -1 gpurun_out/prof_ncu_sym.log
timeout 300 python scripts/configs_1_4.py > gpurun_out/configs_1_4.md 2> gpurun_out/configs_1_4.err; echo "configs rc=$?"
timeout 300 python scripts/phase_bench.py 184 > gpurun_out/phase184.log 2>&1; echo "phase rc=$?"
timeout 900 python scripts/robustness.py > gpurun_out/robustness.md 2> gpurun_out/robustness.err; echo "robustness rc=$?"
