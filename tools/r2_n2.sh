#!/bin/bash
# Round-2 multi-GPU evidence at N GPUs (default 2): exchange tests, default bench line (uniform / Zipf; peer bitmaps on / off),
# size sweep.  usage: tools/r2_n2.sh N tag     (logs in gpurun_out/)
N=${1:-2}; TAG=${2:-r2_n$N}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_dp.py -q -x -p no:cacheprovider > $O/${TAG}_test.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${TAG}_test.log)"
pp() { tail -1 $1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); i=d.get('dp') or {}
    print('$2', 'step_us=%.1f' % (d['ms_per_step']*1e3), 'Mtok/s=%.1f' % (d['value']/1e6), {k:(round(v,4) if isinstance(v,float) else v) for k,v in i.items()})
except Exception as e: print('$2', 'parse error', e)
"; }
B="bench.py --gpus $N --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-torch-gpu"
timeout 200 $TR --master-port 29562 $B > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; pp $O/${TAG}_bench.log "uniform"
timeout 200 $TR --master-port 29563 $B --dist zipf > $O/${TAG}_bench_zipf.log 2> $O/${TAG}_bench_zipf.err; pp $O/${TAG}_bench_zipf.log "zipf"
if [ "$N" -le 4 ]; then
MOT_DP_PEER_BITS=0 timeout 200 $TR --master-port 29564 $B > $O/${TAG}_bench_nopb.log 2> $O/${TAG}_bench_nopb.err; pp $O/${TAG}_bench_nopb.log "uniform peer_bits=0"
MOT_DP_PEER_BITS=0 timeout 200 $TR --master-port 29565 $B --dist zipf > $O/${TAG}_bench_zipf_nopb.log 2> $O/${TAG}_bench_zipf_nopb.err; pp $O/${TAG}_bench_zipf_nopb.log "zipf peer_bits=0"
fi
SWEEP_SIZES=${SWEEP_SIZES:-1024,16384,65536,262144,1048576} timeout 300 $TR --master-port 29561 tools/scale_sweep.py > $O/${TAG}_sweep.log 2>&1; grep "^|" $O/${TAG}_sweep.log
