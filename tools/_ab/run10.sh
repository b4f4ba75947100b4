out=gpurun_out/pdl_modes.txt; : > $out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-vae --no-modes --no-train --no-torch-eager"
for mode in forward inverse joint; do for pdl in 0 1 2; do
UNIB200_PDL=$pdl $B --mode $mode 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$mode pdl=$pdl', round(d['denoise_step_ms'],3), round(d['value'],3))" >> $out
done; done
cat $out
