out=gpurun_out/wi2; mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-vae --no-modes --no-train --no-torch-eager"
run() { name=$1; shift; env "$@" $B 2>$out/$name.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['denoise_step_ms'],3), d['gpu_launches'])" | tee -a $out/whatif.txt; }
run base A=0
run no_attn_small UNIB200_SKIP_ATTN=1
run no_attn_long UNIB200_SKIP_ATTN=2
run no_groupnorm UNIB200_SKIP_KINDS=8
run no_gemm_shortk UNIB200_SKIP_GEMM=1
run base2 A=0
