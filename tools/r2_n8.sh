#!/bin/bash
# Round-2 8-GPU check, as the driver launches it (default flags incl. the e2e leg), then a short size sweep and one
# projection line.  usage: tools/r2_n8.sh [N] [tag]
N=${1:-8}; TAG=${2:-r2_n$N}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29571 bench.py --gpus $N --steps 100 --warmup 10 > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -1 $O/${TAG}_bench.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); i=d.get('dp') or {}
    print('step_us=%.1f' % (d['ms_per_step']*1e3), 'Mtok/s=%.1f' % (d['value']/1e6), 'e2e=%.1f' % ((d.get('e2e') or {}).get('value',0)/1e6), {k:(round(v,4) if isinstance(v,float) else v) for k,v in i.items()})
except Exception as e: print('parse error', e)
"
SWEEP_SIZES=65536,1048576 timeout 100 $TR --master-port 29572 tools/scale_sweep.py > $O/${TAG}_sweep.log 2>&1; grep "^|" $O/${TAG}_sweep.log
timeout 90 $TR --master-port 29573 bench.py --gpus $N --workload mot-proj-spt-64k --steps 20 --warmup 5 --no-e2e > $O/${TAG}_proj.log 2> $O/${TAG}_proj.err; echo "proj rc=$?"
tail -1 $O/${TAG}_proj.log | cut -c1-600
