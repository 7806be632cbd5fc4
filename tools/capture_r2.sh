#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): both bench arms, the launch list and the ncu --set full captures of the
# search launches of one timed encode, exported to CSV on the box (the .ncu-rep files exceed the copy-back limit).
set -x
O=gpurun_out
python bench.py > $O/bench_r2_b200.json 2> $O/bench_r2_b200.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference.json 2> $O/bench_r2_reference.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-batch > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-batch > $O/launches_ncu.log 2>&1
FE_BENCH_STEADY=0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batch > /dev/null 2>&1 || exit 1
FE_BENCH_STEADY=0 ncu --set full --clock-control none -k regex:k_search_i8 -s 8 -c 2 -o /tmp/i8 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batch > $O/ncu_i8.log 2>&1
FE_BENCH_STEADY=0 ncu --set full --clock-control none -k regex:k_search_f16 -s 25 -c 13 -o /tmp/f16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batch > $O/ncu_f16.log 2>&1
ncu -i /tmp/i8.ncu-rep --page raw --csv > $O/raw_i8.csv
ncu -i /tmp/f16.ncu-rep --page raw --csv > $O/raw_f16.csv
ls -la /tmp/*.ncu-rep
