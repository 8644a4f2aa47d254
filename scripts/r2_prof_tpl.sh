#!/bin/bash
# ncu --set full capture of the template numeric PtAP kernel at the headline size (one launch)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu ${BENCH_EXTRA:-}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ptap_numeric_tpl" -s 2 -c 1 -f -o gpurun_out/prof_ptap_tpl${TAG:-} $CMD > gpurun_out/prof_ncu_tpl.log 2>&1
tail -2 gpurun_out/prof_ncu_tpl.log
