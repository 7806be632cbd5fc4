#!/bin/bash
# A/B of planner knobs on the bench workload: prints ms per step and the per-level search times
run() { env "$@" FE_BENCH_STEADY=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-batch 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],3), [(l['T'], l['search_ms'], l['passes'], '%.2e' % l['evaluated']) for l in d['levels']])"; }
for spec in "$@"; do run $spec; done
