MODEL=vit_b_16_384 BATCH=64 STEPS=1 BLOCKS=1 WARM=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_384.csv python tools/time_forward.py > gpurun_out/ncu_384.log 2>&1
tail -1 gpurun_out/ncu_384.log | cut -c1-200
