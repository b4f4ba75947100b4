timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 --launch-skip 3 -c 1 -f -o gpurun_out/r2z_attention_p_tmem_full python tools/bench_attn.py 1 > gpurun_out/r2z_ncu_attn.log 2>&1
tail -3 gpurun_out/r2z_ncu_attn.log
ls -la gpurun_out/*.ncu-rep
