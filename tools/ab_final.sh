run() { env $1 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1:', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))
" || tail -5 /tmp/err.txt; }
run S2AG_PASS_ORDER=32
run S2AG_PASS_ORDER=23
run S2AG_PASS_ORDER=32
run S2AG_PASS_ORDER=23
run S2AG_TCN_FUSED=1
