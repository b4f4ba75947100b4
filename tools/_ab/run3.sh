VARIANTS="7 71 43075 75 87 43091 91" NSHAPES=1 bash tools/ab_attn.sh gpurun_out/ab_attn_r2z3.txt
grep -E "=== variant|attn B|parity d40" gpurun_out/ab_attn_r2z3.txt | cut -c1-230
