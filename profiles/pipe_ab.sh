# A/B on one box: pipe kernel (with / without the L2 persistence window) against the two-kernel path
run() { timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-dqn 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['value']/1e6,1), d['roofline']['kernel_ms_per_step'], d['clocks']['sm_mhz'])"; }
CDQ_PIPE=0 run two-kernel
CDQ_PIPE=1 run pipe+persist
CDQ_PIPE=1 CDQ_PIPE_PERSIST=0 run pipe
CDQ_PIPE=0 run two-kernel
CDQ_PIPE=1 run pipe+persist
CDQ_PIPE=1 CDQ_PIPE_RING=72 run pipe72+persist
CDQ_PIPE=1 CDQ_PIPE_RING=56 run pipe56+persist
[ -f chess-deep-q_b200/libcdq_gf4.so ] && CDQ_PIPE=1 CDQ_LIBCDQ=chess-deep-q_b200/libcdq_gf4.so run pipe-gf4+persist
