O=gpurun_out; R=r1n
python bench.py > $O/bench_$R.log 2>$O/bench_$R.err; echo "bench rc=$?"
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv $B > $O/ncu_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_fwd_kernel -s 4 -c 1 -o $O/prof_${R}_fwd -f $B > $O/ncu_$R.log 2>&1
tail -1 $O/bench_$R.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step']*1e3,1), d['kernel_ms'], round(d['roofline']['frac'],3), d['roofline']['traffic'], 'e2e', round(d['e2e']['value']/1e6,1), d['e2e']['api'][-45:], 'cpu', round(d['cpu_baseline']['value']))"
