timeout 400 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do timeout 120 python profiles/train_timeline_dp.py 2>&1 | grep -E "ms per step|conv1_wgrad|adam_single|step span"; done
