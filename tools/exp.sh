# Ablation builds of the saved-output backward: one piece compiled out per library (never shipped), then the bench.
#   bash tools/exp.sh          (on the GPU box; results go to stdout)
cd mixture-of-tokenizers_b200
for x in "NO_RED:-DMOT_EXPERIMENT_NO_RED" "NO_MATH:-DMOT_X_SUM_NO_MATH" "NO_FLUSH:-DMOT_X_SUM_NO_FLUSH" "NO_OCOPY:-DMOT_X_SUM_NO_OCOPY" \
         "NO_ZERO:-DMOT_X_SUM_NO_ZERO" "ALL:-DMOT_EXPERIMENT_NO_RED -DMOT_X_SUM_NO_MATH -DMOT_X_SUM_NO_FLUSH -DMOT_X_SUM_NO_ZERO"; do
  MOT_LIB_SUFFIX=_x${x%%:*} MOT_EXTRA_NVCC="${x#*:}" python build.py > /dev/null
done
cd ..
B="python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline"
pp() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', 'step_us=%.1f fwd=%.1f bwd=%.1f' % (d['ms_per_step']*1e3, d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3))"; }
for suf in "" _xNO_RED _xNO_MATH _xNO_FLUSH _xNO_OCOPY _xNO_ZERO _xALL; do
  MOT_LIB_SUFFIX=$suf $B | pp "48k lib=$suf"
done
