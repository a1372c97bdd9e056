# ablations of the saved-output backward + one full ncu capture of it
B="python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline"
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f fwd=%.1f bwd=%.1f' % (d['ms_per_step']*1e3, d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3))"; }
for suf in "" _xNO_RED _xNO_MATH _xNO_FLUSH _xNO_OCOPY _xNO_ZERO _xALL; do
MOT_LIB_SUFFIX=$suf $B | pp "48k lib=$suf"
done
for suf in "" _xNO_RED _xNO_MATH _xNO_OCOPY _xALL; do
MOT_LIB_SUFFIX=$suf $B --workload mot-sum-1m --steps 10 | pp "1m lib=$suf"
done
B2="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:mot_bwd_sum_kernel -s 4 -c 1 -o gpurun_out/prof_r1i_bwdsum -f $B2 > gpurun_out/ncu_r1i.log 2>&1
tail -2 gpurun_out/ncu_r1i.log | cut -c1-300
