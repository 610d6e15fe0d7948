# end-to-end (host buffers) A/B of the chunk ramp in cdq_leaf_eval_host, alternating on one box
run() { timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-dqn 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), d['clocks']['sm_mhz'])"; }
CDQ_HOST_RAMP=0 run flat
run ramp
CDQ_HOST_RAMP=0 run flat
run ramp
