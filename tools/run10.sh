timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -2
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f' % (d['ms_per_step']*1e3), d.get('kernel_ms'), round(d['roofline']['frac'],3))"; }
for suf in "" _xW384; do
MOT_LIB_SUFFIX=$suf python bench.py --workload mot-norm-lambdas-71041 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e | pp "v3d lib=$suf"
done
python bench.py --workload mot-concat-711 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e | pp v4
python bench.py --workload mot-sum-medium-64k --steps 50 --warmup 10 --no-cpu-baseline --no-e2e | pp v3-1024
MOT_NO_SAVED_BWD=1 python bench.py --workload mot-sum-medium-64k --steps 50 --warmup 10 --no-cpu-baseline --no-e2e | pp v3-1024-recompute
