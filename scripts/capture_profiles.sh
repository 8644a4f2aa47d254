#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + full captures of the two top kernels.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/prof_plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/prof_plain_bench.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/prof_ncu_list.log 2>&1
tail -1 gpurun_out/prof_ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 60 -c 2 -o gpurun_out/prof_spmv_sell $CMD > gpurun_out/prof_ncu_sell.log 2>&1
tail -1 gpurun_out/prof_ncu_sell.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ptap_numeric" -s 2 -c 1 -o gpurun_out/prof_ptap_numeric $CMD > gpurun_out/prof_ncu_ptap.log 2>&1
tail -1 gpurun_out/prof_ncu_ptap.log
timeout 600 ncu --set full --clock-control none -k regex:"k_cg_update|k_cg_p" -s 60 -c 2 -o gpurun_out/prof_cg_vec $CMD > gpurun_out/prof_ncu_cgvec.log 2>&1
tail -1 gpurun_out/prof_ncu_cgvec.log
