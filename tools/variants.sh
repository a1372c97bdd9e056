#!/bin/bash
# bench the default workload (and the 64K x 1024 shape) with experiment builds of the library: tools/variants.sh suffix...
for sfx in "" "$@"; do
  for wl in mot-sum-124M-48k mot-sum-medium-64k; do
    MOT_LIB_SUFFIX=$sfx python bench.py --workload $wl --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-torch-gpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lib=$sfx', '$wl', 'step_us=%.1f' % (d['ms_per_step']*1e3), 'fwd=%.1f bwd=%.1f' % (d['kernel_ms']['fwd']*1e3, d['kernel_ms']['bwd_main']*1e3), 'frac=%.3f' % d['roofline']['frac'])"
  done
done
