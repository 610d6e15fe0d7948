# training-step check on one box: train tests, bench DQN line, ncu launch list of the eager step
python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('leaf', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'dqn', d['dqn'])"
STEPS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/train_launches_$1.csv python profiles/train_step_probe.py > gpurun_out/ncu_$1.log 2>&1; tail -1 gpurun_out/ncu_$1.log
