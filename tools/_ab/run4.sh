VARIANTS="7 71 199 87 215 43091 43219 219" NSHAPES=1 bash tools/ab_attn.sh gpurun_out/ab_attn_r2z4.txt
grep -E "=== variant|attn B" gpurun_out/ab_attn_r2z4.txt | cut -c1-230
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-vae --no-modes --no-train --no-torch-eager"
for v in 7 71 87 43091; do
UNIB200_ATTN_VARIANT=$v $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step variant $v', round(d['denoise_step_ms'],3))" | tee -a gpurun_out/ab_attn_r2z4.txt
done
