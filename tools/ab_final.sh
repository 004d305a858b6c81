run() { env $1 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1:', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))
" 2>/dev/null || echo "$1: FAILED $(grep -m1 'CUDA error\|Error' /tmp/err.txt)"; }
run S2AG_STREAM_PRIO=-1,0,0,0
run S2AG_STREAM_PRIO=-1,-1,0,0
run S2AG_STREAM_PRIO=-1,-1,-1,0
run S2AG_STREAM_PRIO=-2,-1,0,0
run S2AG_STREAM_PRIO=-1,0,0,0
