run() { env "$@" FE_BENCH_STEADY=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-batch 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],3), [(l['T'], l['search_ms']) for l in d['levels']])"; }
run A=1
run FE_F16_ITEMS=12
run FE_F16_ITEMS=16
run FE_F16_ITEMS=16 FE_F16_MIN_RUN=8
run FE_F16_ITEMS=6
run FE_F16_MIN_RUN=32
run FE_I8_ITEMS=4
run FE_I8_ITEMS=9
run FE_I8_ITEMS=12 FE_I8_MIN_RUN=16
run FE_I8_MIN_RUN=48
