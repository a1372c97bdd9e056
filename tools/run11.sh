pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f' % (d['ms_per_step']*1e3), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step']*1e3,1), 'us', d['e2e']['api'][-60:])"; }
python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2> gpurun_out/e2e_graph.err | pp graphed; tail -3 gpurun_out/e2e_graph.err
MOT_E2E_FAIL=1 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2> gpurun_out/e2e_fail.err | pp injected-failure; tail -2 gpurun_out/e2e_fail.err | cut -c1-300
MOT_E2E_EAGER=1 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | pp eager
python bench.py --workload mot-sum-medium-64k --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | pp medium
python bench.py --workload mot-norm-lambdas-71041 --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | pp v3d
