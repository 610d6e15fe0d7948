# share of the SMs given to the online network's conv launch in the training forward (CDQ_TRAIN_CONV_SPLIT, percent)
for pct in 35 40 45 50 55 61; do
  CDQ_TRAIN_CONV_SPLIT=$pct timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('online $pct %', round(d['dqn']['ms_per_step']*1e3,1), 'us', round(d['dqn']['value']/1e6,2), 'M samples/s')"
done
