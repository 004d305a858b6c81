run() { env $1 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1:', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))
" 2>/dev/null || echo "$1: FAILED $(grep -m1 'CUDA error' /tmp/err.txt)"; }
run "S2AG_TCN_FUSED=1 S2AG_EARLY_FORK=0 S2AG_MFCC_STREAM=0"
run "S2AG_TCN_FUSED=1 S2AG_TCN_WGRAD_STREAM=c"
run "S2AG_TCN_FUSED=1 S2AG_CONV_WGRAD_STREAM=0 S2AG_TCN_WGRAD_STREAM=c"
run "S2AG_TCN_FUSED=1 S2AG_STREAM_PRIO=0,0,0"
