out=gpurun_out/ab_attn_r2z8.txt; : > $out
for v2 in 0 64 65; do
  echo "=== variant2 $v2" >> $out
  UNIB200_ATTN_VARIANT2=$v2 timeout 120 python tests/gpu_probe.py attention_d80 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('parity d80:', {k: (round(v['rel_l2'],6), round(v['rel_to_max'],6)) for k, v in d.items()})" >> $out 2>&1
  UNIB200_ATTN_VARIANT2=$v2 timeout 120 python tools/bench_attn.py 5 >> $out 2>&1
done
cat $out | cut -c1-400
