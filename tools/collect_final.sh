#!/bin/bash
# Round-end evidence collection on one B200 (writes gpurun_out/f_*): bench lines of the BASELINE configurations, the
# context arms on the same box, the per-step timeline, the ncu launch list of one eager iteration.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/f_bench_b256.json 2> $O/f_bench_b256.err
python bench.py --steps 10 --warmup 3 --batch-per-gpu 512 --no-cpu-baseline > $O/f_bench_b512.json 2> /dev/null
python bench.py --steps 10 --warmup 3 --batch-per-gpu 512 --precision bf16x1 --no-cpu-baseline > $O/f_bench_b512_bf16x1.json 2> /dev/null
python bench.py --steps 10 --warmup 3 --batch-per-gpu 384 --no-cpu-baseline > $O/f_bench_b384.json 2> /dev/null
python bench.py --steps 20 --warmup 5 --batch-per-gpu 128 --no-cpu-baseline > $O/f_bench_b128.json 2> /dev/null
python bench.py --impl reference-gpu --steps 5 --warmup 2 > $O/f_reference_gpu.json 2> /dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/f_reference_cpu.json 2> /dev/null
python tools/bench_longform.py > $O/f_longform.json 2> /dev/null
S2AG_TRACE=$O/f_trace.json python tools/step_timeline.py 256 > $O/f_timeline.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 900 --csv --log-file $O/f_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/f_ncu_bench.log 2>&1
ls -la $O/f_*
