B="python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline"
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f fwd=%.1f bwd=%.1f' % (d['ms_per_step']*1e3, d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3))"; }
for st in 2 3 4; do for th in 512 768 1024; do
MOT_STAGES=$st MOT_FWD_THREADS=$th $B | pp "48k stages=$st fwd_threads=$th"
done; done
MOT_STAGES=4 MOT_FWD_THREADS=768 $B --workload mot-sum-medium-64k | pp "64k-1024 stages=4 768"
MOT_STAGES=2 MOT_FWD_THREADS=768 $B --workload mot-sum-medium-64k | pp "64k-1024 stages=2 768"
MOT_STAGES=4 MOT_FWD_THREADS=768 $B --tokens 1048576 --steps 10 | pp "1m stages=4 768"
