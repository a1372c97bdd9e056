B="python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline"
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f fwd=%.1f bwd=%.1f' % (d['ms_per_step']*1e3, d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3))"; }
for suf in "" _xREP4 _xREP8 _xREP32 _xREP64; do
MOT_LIB_SUFFIX=$suf $B | pp "48k lib=$suf"
MOT_LIB_SUFFIX=$suf $B --dist zipf | pp "48k zipf lib=$suf"
MOT_LIB_SUFFIX=$suf $B --tokens 1048576 --steps 10 | pp "1m-768 lib=$suf"
done
for n in 131072 262144 524288; do
$B --tokens $n --steps 20 | pp "saved N=$n"
MOT_NO_SAVED_BWD=1 $B --tokens $n --steps 20 | pp "recompute N=$n"
done
