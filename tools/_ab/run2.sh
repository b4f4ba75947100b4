VARIANTS="7 67 71 75 43075" NSHAPES=1 bash tools/ab_attn.sh gpurun_out/ab_attn_r2z2.txt
grep -E "=== variant|attn B|parity d40" gpurun_out/ab_attn_r2z2.txt | cut -c1-230
for l in "" _gelu1 _gelu2; do
  echo "== lib$l" >> gpurun_out/gelu_ab.txt
  UNIB200_LIB=$PWD/uni_renderer_b200/libunib200$l.so python tools/bench_gemm.py 2>&1 | grep -E "geglu|shape" >> gpurun_out/gelu_ab.txt
done
UNIB200_LIB=$PWD/uni_renderer_b200/libunib200_gelu2.so python tests/gpu_probe.py conv_variants 2>&1 | tail -1 | cut -c1-1500 >> gpurun_out/gelu_ab.txt
cat gpurun_out/gelu_ab.txt
