#!/bin/bash
# A/B of conv2_bwd2_kernel build variants (-DCB2_*): stand-alone duration of the kernel under ncu (cold L2, serialised)
for v in "" $VARIANTS; do
  so=/root/repo/chess-deep-q_b200/libcdq$v.so
  [ -f $so ] || continue
  CDQ_LIBCDQ=$so timeout 200 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:"conv2_bwd2" -c 3 --csv --log-file gpurun_out/cb2$v.csv python profiles/train_timeline_dp.py > /dev/null 2>&1
  echo "variant '$v': $(grep -o 'gpu__time_duration.sum","ns","[0-9]*' gpurun_out/cb2$v.csv | grep -o '[0-9]*$' | tr '\n' ' ') ns; cycles $(grep -o 'sm__cycles_elapsed.max","cycle","[0-9]*' gpurun_out/cb2$v.csv | grep -o '[0-9]*$' | tr '\n' ' '); tensor $(grep -o 'sustained_active","%","[0-9.]*' gpurun_out/cb2$v.csv | grep -o '[0-9.]*$' | tr '\n' ' ')"
done
