out=gpurun_out/ab_attn_r2z7.txt; : > $out
for cfg in "0 71" "1 71" "1 67"; do
  set -- $cfg
  echo "=== split $1 variant $2" >> $out
  UNIB200_ATTN_SPLIT=$1 UNIB200_ATTN_VARIANT=$2 timeout 120 python tests/gpu_probe.py attention_d40 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('parity d40:', {k: (round(v['rel_l2'],6), round(v['rel_to_max'],6)) for k, v in d.items()})" >> $out 2>&1
  UNIB200_ATTN_SPLIT=$1 UNIB200_ATTN_VARIANT=$2 timeout 120 python tools/bench_attn.py 1 >> $out 2>&1
done
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-vae --no-modes --no-train --no-torch-eager"
for sp in 0 1; do
UNIB200_ATTN_SPLIT=$sp $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step split $sp', round(d['denoise_step_ms'],3))" >> $out
done
cat $out | cut -c1-400
