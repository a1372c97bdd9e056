# ncu evidence for profiles/: launch list of a short bench run, then one full capture each of the two main kernels
set -e
R=${1:-r1}
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-gpu"
$B > gpurun_out/plain_$R.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $B > gpurun_out/ncu_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_bwd_sum_kernel -s 4 -c 1 -o gpurun_out/prof_${R}_bwdsum -f $B > gpurun_out/ncu_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_fwd_kernel -s 4 -c 1 -o gpurun_out/prof_${R}_fwd -f $B >> gpurun_out/ncu_$R.log 2>&1
tail -2 gpurun_out/ncu_$R.log
