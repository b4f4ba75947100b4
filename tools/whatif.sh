#!/bin/bash
# what-if marginal costs inside the real two-lane step graph: each line drops one op class (results are garbage)
set -u
out=${1:-gpurun_out/wi}
mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-vae --no-modes --no-train --no-torch-eager"
run() { name=$1; shift; env "$@" $B 2>$out/$name.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['denoise_step_ms'],3), d['gpu_launches'])" | tee -a $out/whatif.txt; }
run base A=0
run no_attention UNIB200_SKIP_KINDS=4
run no_groupnorm UNIB200_SKIP_KINDS=8
run no_gemm_shortk UNIB200_SKIP_GEMM=1
run no_gemm_geglu UNIB200_SKIP_GEMM=2
run no_gemm_splitk UNIB200_SKIP_GEMM=4
run no_finalize UNIB200_SKIP_FINALIZE=1
run no_gn_cluster UNIB200_SKIP_GN_CLUSTER=1
run no_finalize_no_gn_cluster UNIB200_SKIP_FINALIZE=1 UNIB200_SKIP_GN_CLUSTER=1
run no_gemm_pair UNIB200_SKIP_GEMM=8
run no_gemm_long_single UNIB200_SKIP_GEMM=16
run base2 A=0
