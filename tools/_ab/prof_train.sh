mkdir -p gpurun_out/r2w
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w/train_launches.csv --launch-skip 20000 --launch-count 40000 python tools/bench_train.py --batch 2 --latent 64 --steps 1 --warmup 0 > gpurun_out/r2w/prof.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r2w/train_launches.csv 2>/dev/null | head -40
