set -e
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_r1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_bwd_kernel -s 4 -c 1 -o gpurun_out/prof_r1e_bwd -f $B > gpurun_out/ncu_r1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_fwd_kernel -s 4 -c 1 -o gpurun_out/prof_r1e_fwd -f $B >> gpurun_out/ncu_r1e.log 2>&1
tail -3 gpurun_out/ncu_r1e.log
