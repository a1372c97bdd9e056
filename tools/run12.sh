pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f' % (d['ms_per_step']*1e3), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step']*1e3,1), 'us', d['e2e']['api'][-60:])"; }
python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2> gpurun_out/e2e_graph.err | pp graphed; grep -v Warning gpurun_out/e2e_graph.err | tail -3 | cut -c1-300
python bench.py --workload mot-norm-lambdas-71041 --steps 50 --warmup 10 --no-cpu-baseline 2> gpurun_out/e2e_v3d.err | pp v3d; grep "bench:" gpurun_out/e2e_v3d.err | cut -c1-400
python bench.py --workload mot-concat-711 --steps 50 --warmup 10 --no-cpu-baseline 2> gpurun_out/e2e_v4.err | pp v4; grep "bench:" gpurun_out/e2e_v4.err | cut -c1-400
