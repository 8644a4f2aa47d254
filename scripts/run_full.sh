#!/bin/bash
# Full GPU validation on one GPU: the whole `-m gpu` suite, then everything profiles/ is built from.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40 > gpurun_out/full_tests.log
tail -5 gpurun_out/full_tests.log
bash scripts/capture_profiles.sh
