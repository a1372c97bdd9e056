# GPU check of the saved-output backward: parity tests, then bench A/B (saved vs recompute), 48K and 1M tokens
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1i.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r1i.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1i_saved.log 2>&1; echo "bench rc=$?"
MOT_NO_SAVED_BWD=1 python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1i_recompute.log 2>&1
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --tokens 1048576 > gpurun_out/bench_r1i_saved_1m.log 2>&1
python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline --workload mot-sum-medium-64k > gpurun_out/bench_r1i_saved_1024.log 2>&1
for st in 2 3; do MOT_SUM_STAGES=$st python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1i_saved_st$st.log 2>&1; done
for f in gpurun_out/bench_r1i_*.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print(d['value']/1e6, 'Mtok/s', d['ms_per_step'], d['kernel_ms'], d['roofline']['frac'], d.get('e2e', {}))
except Exception as e: print('parse error', e)
"; done
