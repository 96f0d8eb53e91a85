# Where the conv family's time goes (default arithmetic): the same bench step with parts of conv_umma_kernel switched off
# (NSC_UMMA_SKIP bits: 2 = no fp16 MMAs, 4 = no epilogue body, 8 = no A-operand TMA loads, 16 = no weight loads).  Results are
# garbage by construction; only the times and the SM clock (power cap!) of each run are read.  Output: one JSON line per run.
CMD="python bench.py --steps 16 --warmup 3 --no-extra --no-cpu-baseline --no-e2e --min-timed-ms 600"
for skip in 0 16 8 4 24 28 2 0; do
  NSC_UMMA_SKIP=$skip $CMD 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(json.dumps({'skip': $skip, 'ms_per_step': round(d['ms_per_step'],3), 'sm_mhz': d['clocks']['sm_mhz'], 'power_w': d['clocks'].get('power_w_median'), 'per_kernel_ms': r['per_kernel_ms']}))"
done
