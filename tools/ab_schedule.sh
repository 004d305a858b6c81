run() { S2AG_PASS_ORDER=$1 S2AG_TRI_LATE=$2 S2AG_STREAM_PRIO=$3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('order $1 tri $2 prio $3:', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))
"; }
run 32 mid 0,0,0
run 32 mid -2,0,-1
run 32 mid -1,0,0
run 23 1 -2,0,-1
run 23 mid -2,0,-1
run 32 1 -2,0,-1
run 32 mid -2,-1,0
