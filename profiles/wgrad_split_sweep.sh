#!/bin/bash
# fc1 weight gradient: K splits (CTAs = 32 x splits) against the step time; the kernel is off the critical path, what it
# costs is SM-time
for rep in 1 2; do
for k in 1 2 3 4; do
  echo "CDQ_WGRAD_SPLITS=$k: $(CDQ_WGRAD_SPLITS=$k timeout 120 python profiles/train_timeline_dp.py 2>&1 | grep -E 'ms per step|fc1_wgrad' | sed 's/# world 1, global batch 4096 (4096 per rank): //' | tr '\n' ' ')"
done; done
