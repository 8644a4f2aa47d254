mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 > gpurun_out/verify_tests.log; tail -3 gpurun_out/verify_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/configs_1_4.py > gpurun_out/configs_1_4.md 2> gpurun_out/configs_1_4.err; tail -7 gpurun_out/configs_1_4.md
timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['ptap_numeric']['ms'], d['config']['cold_repeat_ms'])"
