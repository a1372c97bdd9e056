B="python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline"
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f fwd=%.1f bwd=%.1f' % (d['ms_per_step']*1e3, d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3))"; }
for wl in mot-sum-124M-48k mot-sum-medium-64k mot-sum-1m; do
for thr in 1024 768 512; do for st in 4 3 2; do
MOT_FWD_THREADS=$thr MOT_STAGES=$st $B --workload $wl | pp "$wl fwd_threads=$thr stages=$st"
done; done; done
