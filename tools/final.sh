# Round-end evidence in one call: GPU tests, smoke, default bench (both arms), the other workloads, ncu launch list and
# one full capture each of the forward and the saved-output backward kernel.
R=${1:-r1j}
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_$R.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_$R.log)"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$R.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke_$R.log | cut -c1-200)"
python bench.py > $O/bench_$R.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_$R.log 2>&1; echo "bench ref rc=$?"
for w in mot-sum-medium-64k mot-norm-lambdas-71041 mot-concat-711 mot-proj-runs7-64k mot-proj-spt-64k mot-proj-spt-addpp-64k value-embeds-64k mathblations-concat; do
  python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline > $O/bench_${w}_$R.log 2>&1; echo "$w rc=$?"
done
python bench.py --dist zipf --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > $O/bench_zipf_$R.log 2>&1
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv $B > $O/ncu_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_bwd_sum_kernel -s 4 -c 1 -o $O/prof_${R}_bwdsum -f $B > $O/ncu_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mot_fwd_kernel -s 4 -c 1 -o $O/prof_${R}_fwd -f $B >> $O/ncu_$R.log 2>&1
for f in $O/bench_*_$R.log $O/bench_$R.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,1),'us', d.get('kernel_ms'), round(d['roofline']['frac'],3) if 'roofline' in d else '', 'e2e', round(d.get('e2e', {}).get('value', 0)/1e6, 1), 'cpu', d.get('cpu_baseline', {}).get('value'))
except Exception as e: print('parse error', e)
"; done
