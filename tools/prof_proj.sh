# ncu evidence for the projection path: launch list + one full capture of each GEMM flavour
set -e
R=${1:-r1}
B="python bench.py --workload mot-proj-runs7-64k --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_proj_$R.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_proj_$R.csv $B > gpurun_out/ncu_launches_proj_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_gemm_kernel -s 9 -c 3 -o gpurun_out/prof_${R}_gemm -f $B > gpurun_out/ncu_proj_$R.log 2>&1
tail -2 gpurun_out/ncu_proj_$R.log
