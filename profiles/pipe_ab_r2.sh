for v in 0 1 0 1; do
  CDQ_PIPE=$v python bench.py --no-cpu-baseline --no-gpu-baseline --no-selfplay --no-dqn 2>/dev/null > gpurun_out/ab_$v.json
  python -c "
import sys,json
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print('pipe', $v, round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['clocks']['sm_mhz'], round(d['clocks']['power_w_median']))"
done
