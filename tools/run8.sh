B="python bench.py --workload mot-norm-lambdas-71041 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:mot_bwd_kernel -s 3 -c 1 -o gpurun_out/prof_r1j_bwd_v3d -f $B > gpurun_out/ncu_v3d.log 2>&1
tail -1 gpurun_out/ncu_v3d.log | cut -c1-200
python bench.py --workload mot-norm-lambdas-71041 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --tokens 16384 | tail -1 | cut -c1-100
