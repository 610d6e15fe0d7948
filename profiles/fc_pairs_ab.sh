# fc head with two tiles per weight stage (default) against one (CDQ_FC_PAIRS=0), alternating on one box
run() { timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-dqn 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['value']/1e6,1), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_median'))"; }
CDQ_FC_PAIRS=0 run single
run pairs
CDQ_FC_PAIRS=0 run single
run pairs
