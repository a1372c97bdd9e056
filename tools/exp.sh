B="python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline"
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f fwd=%.1f bwd=%.1f' % (d['ms_per_step']*1e3, d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3))"; }
for suf in "" _xNO_RED _xNO_TCOPY _xNO_MATH _xNO_FLUSH _xALL; do
MOT_LIB_SUFFIX=$suf $B | pp "48k lib=$suf"
done
for suf in "" _xNO_MATH _xALL; do
MOT_LIB_SUFFIX=$suf $B --workload mot-sum-1m --steps 10 | pp "1m lib=$suf"
done
