# sweep of the pipe kernel's conv share (% of SMs) and ring size (MiB); bench value in M positions/s
for ring in 60 64 68; do for pct in 60 63 66 69 72; do
  CDQ_PIPE=1 CDQ_PIPE_CONV_PCT=$pct CDQ_PIPE_RING=$ring timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-dqn 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ring $ring pct $pct', round(d['value']/1e6,1), round(d['roofline']['kernel_ms_per_step'].get('leaf_pipe_kernel'),3))"
done; done
