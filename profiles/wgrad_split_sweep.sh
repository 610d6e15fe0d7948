# K splits of fc1's weight gradient (CDQ_WGRAD_SPLITS): CTAs = 32 x splits, the kernel runs on the low-priority side stream
for k in 1 2 3 4 6; do
  CDQ_WGRAD_SPLITS=$k timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('splits $k', round(d['dqn']['ms_per_step']*1e3,1), 'us', round(d['dqn']['value']/1e6,2), 'M samples/s')"
done
