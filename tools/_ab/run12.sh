python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-train --no-vae > gpurun_out/r2z_bench_joint_n2.json 2> gpurun_out/r2z_bench_n2.err
tail -c 600 gpurun_out/r2z_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2z_bench_joint_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['denoise_step_ms'], d.get('clocks'))
print({k:(round(v['value'],3),round(v['denoise_step_ms'],3)) for k,v in d.get('modes',{}).items()})
P
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
