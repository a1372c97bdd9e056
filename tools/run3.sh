B="python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline"
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f fwd=%.1f bwd=%.1f' % (d['ms_per_step']*1e3, d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3))"; }
for suf in "" _xT384P2 _xT512P2 _xT640P2 _xT768P2; do
MOT_LIB_SUFFIX=$suf $B | pp "48k lib=$suf"
MOT_LIB_SUFFIX=$suf $B --workload mot-sum-medium-64k | pp "64k-1024 lib=$suf"
MOT_LIB_SUFFIX=$suf $B --tokens 1048576 --steps 10 | pp "1m-768 lib=$suf"
done
MOT_NO_SAVED_BWD=1 $B --tokens 1048576 --steps 10 | pp "1m-768 recompute"
MOT_NO_SAVED_BWD=1 $B --workload mot-sum-medium-64k | pp "64k-1024 recompute"
