#!/bin/bash
# one multi-GPU session: dp test, exchange sweep, timelines, bench with 1 / 2 / 4 slabs.  usage: tools/run_dp8.sh N tag
N=${1:-8}; TAG=${2:-r2_dp8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_dp.py -q -x -p no:cacheprovider > gpurun_out/${TAG}_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_test.log
timeout 240 $TR --master-port 29521 tools/dp_bench.py > gpurun_out/${TAG}_sweep.log 2>&1
for k in 1; do MOT_LIB_SUFFIX=_trace timeout 120 $TR --master-port 2953$k tools/dp_trace.py $k > gpurun_out/${TAG}_trace_$k.log 2>&1; done
for k in 1 2 4; do MOT_DP_SLABS=$k timeout 200 $TR --master-port 2954$k bench.py --gpus $N --steps 50 --warmup 10 --no-e2e > gpurun_out/${TAG}_bench_s$k.log 2> gpurun_out/${TAG}_bench_s$k.err; done
