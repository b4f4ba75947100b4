python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2z2_pytest_gpu.log
cat gpurun_out/r2z2_pytest_gpu.log
python bench.py > gpurun_out/r2z2_bench_joint.json 2> gpurun_out/r2z2_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2z2_bench_joint.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['denoise_step_ms'], d['step_frac_of_tensor_peak'], d['roofline']['frac'])
print({k:(round(v['value'],3),round(v['denoise_step_ms'],3)) for k,v in d['modes'].items()})
P
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2z_step_metrics_joint_b4.csv python tools/profile_step.py --mode joint > gpurun_out/r2z_profile_step.log 2>&1
tail -2 gpurun_out/r2z_profile_step.log
python tools/ncu_step_metrics.py gpurun_out/r2z_step_metrics_joint_b4.csv gpurun_out/r2z_gemm_traffic.json > gpurun_out/r2z_step_metrics_joint_b4_summary.txt 2>&1
head -20 gpurun_out/r2z_step_metrics_joint_b4_summary.txt
