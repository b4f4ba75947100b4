#!/bin/bash
# A/B of the attention softmax variants (UNIB200_ATTN_VARIANT bits: 1 packed, 2 stagger, 4 poly 1/4, 8 poly 1/2):
# per-op parity + micro-benchmark per variant.  Optimisation instrument.
out=${1:-gpurun_out/ab_attn.txt}
: > $out
for v in ${VARIANTS:-0 1 2 3 5 7 9 11}; do
  echo "=== variant $v" >> $out
  UNIB200_ATTN_VARIANT=$v timeout 120 python tests/gpu_probe.py attention_d40 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('parity d40:', {k: (v['rel_l2'], v['rel_to_max']) for k, v in d.items()})" >> $out 2>&1
  UNIB200_ATTN_VARIANT=$v timeout 120 python tests/gpu_probe.py attention_d80 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('parity d80:', {k: (v['rel_l2'], v['rel_to_max']) for k, v in d.items()})" >> $out 2>&1
  UNIB200_ATTN_VARIANT=$v timeout 120 python tools/bench_attn.py ${NSHAPES:-3} >> $out 2>&1
done
