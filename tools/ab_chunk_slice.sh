CMD="python bench.py --steps 16 --warmup 3 --no-extra --no-cpu-baseline --min-timed-ms 1500"
run() { env "$@" $CMD 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', round(d['value']/1e6,4), round(d['e2e']['value']/1e6,4), round(d['e2e_f32']['value']/1e6,4), d['clocks']['sm_mhz'])"; }
run A=0
run NSC_CHUNK=65536
run NSC_HOST_SLICE=32768
run A=0
run NSC_CHUNK=65536
run NSC_HOST_SLICE=32768
run NSC_HOST_SLICE=8192
