#!/bin/bash
# stream SpMV + shim tests, phase table (M x / M^T x), cold-path stage timings at N_b=184 and on config 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short -k "spmv or shim or nonlinear or golden or fullsize or end_to_end or cube" 2>&1 | tail -15 > gpurun_out/cold_tests.log
tail -4 gpurun_out/cold_tests.log
SKIP_FGMRES=1 timeout 200 python scripts/phase_bench.py 184 > gpurun_out/phase184_stream.log 2>&1; grep -E "spmv" gpurun_out/phase184_stream.log
IIFE_SPMV_STREAM=0 SKIP_FGMRES=1 timeout 200 python scripts/phase_bench.py 184 2>&1 | grep -E "spmv\(M\)"
IIFE_PLAN_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/cold184.json 2> gpurun_out/cold184.err; grep "\[plan\]" gpurun_out/cold184.err | head -40
IIFE_PLAN_DEBUG=1 timeout 200 python scripts/configs_1_4.py > gpurun_out/configs_dbg.md 2> gpurun_out/configs_dbg.err; grep "\[plan\]" gpurun_out/configs_dbg.err | head -80
