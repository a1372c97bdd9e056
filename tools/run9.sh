timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f' % (d['ms_per_step']*1e3), d.get('kernel_ms'), round(d['roofline']['frac'],3))"; }
for w in mot-norm-lambdas-71041 value-embeds-64k mot-sum-124M-48k; do
python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline --no-e2e | pp $w
done
