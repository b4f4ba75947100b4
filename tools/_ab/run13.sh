out=gpurun_out/ab_attn_nreg.txt; : > $out
for l in "" _nreg184 _nreg192; do
  echo "=== lib$l" >> $out
  UNIB200_LIB=$PWD/uni_renderer_b200/libunib200$l.so timeout 120 python tests/gpu_probe.py attention_d40 2>&1 | tail -1 | cut -c1-300 >> $out
  UNIB200_LIB=$PWD/uni_renderer_b200/libunib200$l.so timeout 120 python tools/bench_attn.py 2 >> $out 2>&1
  UNIB200_ATTN_VARIANT=67 UNIB200_LIB=$PWD/uni_renderer_b200/libunib200$l.so timeout 120 python tools/bench_attn.py 1 >> $out 2>&1
done
cat $out
