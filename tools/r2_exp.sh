# Round-2 experiment pass: the saved-output backward with its byte gradients through TMA bulk reductions
# (MOT_SUM_BULK_RED builds, never shipped unless they win) against the shipped library, then parity of the variant.
#   bash tools/r2_exp.sh     (on the GPU box; results in gpurun_out/r2_exp.log)
O=gpurun_out
mkdir -p $O
L=$O/r2_exp.log
: > $L
B="python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-torch-gpu"
pp() { tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'step_us=%.1f' % (d['ms_per_step']*1e3), 'fwd=%.1f bwd=%.1f' % (d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3), 'frac=%.3f' % d['roofline']['frac'])
except Exception as e: print('$1', 'parse error', e)
"; }
run() {  # suffix stages workload
  MOT_LIB_SUFFIX=$1 MOT_SUM_STAGES=$2 timeout 120 $B --workload $3 2>>$O/r2_exp.err | pp "lib=$1 stages=$2 $3" >> $L
}
run "" 4 mot-sum-124M-48k
for st in 4 5 6; do run _xBULK $st mot-sum-124M-48k; done
for st in 3 4; do run _xBULK16 $st mot-sum-124M-48k; done
run _xBULKR8 4 mot-sum-124M-48k
run "" 4 mot-sum-medium-64k
run _xBULK 4 mot-sum-medium-64k
run _xBULK16 3 mot-sum-medium-64k
MOT_LIB_SUFFIX= timeout 120 $B --dist zipf 2>>$O/r2_exp.err | pp "lib= zipf" >> $L
MOT_LIB_SUFFIX=_xBULK timeout 120 $B --dist zipf 2>>$O/r2_exp.err | pp "lib=_xBULK zipf" >> $L
for suf in _xBULK _xBULK16; do
  MOT_LIB_SUFFIX=$suf timeout 300 python -m pytest -q -x tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu \
      -k "headline or elementwise or slabs or saved_output or full_size_properties or V3_256k or V3_zipf" > $O/r2_exp_pytest$suf.log 2>&1
  echo "pytest $suf rc=$? $(tail -1 $O/r2_exp_pytest$suf.log)" >> $L
done
cat $L
