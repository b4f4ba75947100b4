python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2z_pytest_gpu.log
cat gpurun_out/r2z_pytest_gpu.log
python bench.py > gpurun_out/r2z_bench_joint.json 2> gpurun_out/r2z_bench.err
tail -c 3000 gpurun_out/r2z_bench_joint.json
