# Round-2 evidence in one 1-GPU call: GPU tests, smoke, default bench (ours / reference / torch-gpu arms), the other
# workloads, the 1-GPU size sweep (BASELINE config 5), ncu launch list and one full capture each of the forward and
# the saved-output backward kernel.      bash tools/final_r2.sh [tag]
R=${1:-r2f}
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_$R.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_$R.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$R.log 2>&1; echo "smoke rc=$? $(tail -2 $O/smoke_$R.log | head -1 | cut -c1-200)"
timeout 400 python bench.py > $O/bench_$R.log 2> $O/bench_$R.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_$R.log 2> $O/bench_ref_$R.err; echo "bench ref rc=$?"
for w in mot-sum-medium-64k mot-sum-1m mot-norm-lambdas-71041 mot-concat-711 mot-proj-runs7-64k mot-proj-spt-64k mot-proj-spt-bpt32-64k mot-proj-spt-addpp-64k value-embeds-64k mathblations-concat; do
  timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu > $O/bench_${w}_$R.log 2> $O/bench_${w}_$R.err; echo "$w rc=$?"
done
timeout 120 python bench.py --dist zipf --steps 50 --warmup 10 --no-cpu-baseline --no-torch-gpu > $O/bench_zipf_$R.log 2> $O/bench_zipf_$R.err
SWEEP_SIZES=1024,4096,16384,65536,262144,1048576 timeout 200 python tools/scale_sweep.py > $O/sweep1_$R.log 2>&1
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-gpu"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv $B > $O/ncu_launches_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mot_bwd_sum_kernel -s 4 -c 1 -o $O/prof_${R}_bwdsum -f $B > $O/ncu_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mot_fwd_kernel -s 4 -c 1 -o $O/prof_${R}_fwd -f $B >> $O/ncu_$R.log 2>&1
for f in $O/bench_$R.log $O/bench_*_$R.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,1),'us', {k: round(v*1e3,1) for k,v in (d.get('kernel_ms') or {}).items()}, round(d['roofline']['frac'],3) if 'roofline' in d else '', 'e2e', round((d.get('e2e') or {}).get('value', 0)/1e6, 1), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'torch-gpu', d.get('torch_gpu'))
except Exception as e: print('parse error', e)
"; done
grep "^|" $O/sweep1_$R.log
