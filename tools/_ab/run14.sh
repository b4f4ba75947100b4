out=gpurun_out/ab_attn_setmaxnreg.txt; : > $out
for v in 71 67 43091; do
echo "=== setmaxnreg build, variant $v" >> $out
UNIB200_ATTN_VARIANT=$v timeout 120 python tests/gpu_probe.py attention_d40 2>&1 | tail -1 | cut -c1-400 >> $out
UNIB200_ATTN_VARIANT=$v timeout 120 python tools/bench_attn.py 1 >> $out 2>&1
done
timeout 120 python tests/gpu_probe.py attention_d80 2>&1 | tail -1 | cut -c1-300 >> $out
timeout 120 python tests/gpu_probe.py attention_d160 2>&1 | tail -1 | cut -c1-300 >> $out
timeout 120 python tools/bench_attn.py 5 >> $out 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-vae --no-modes --no-train --no-torch-eager"
$B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step', round(d['denoise_step_ms'],3))" >> $out
cat $out
