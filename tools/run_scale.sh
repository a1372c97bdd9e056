#!/bin/bash
# usage: tools/run_scale.sh N tag : size sweep + default bench line (+ projection line) at N GPUs
N=${1:-8}; TAG=${2:-r2_scale}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29561 tools/scale_sweep.py > gpurun_out/${TAG}_sweep.log 2>&1
timeout 200 $TR --master-port 29562 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
timeout 200 $TR --master-port 29563 bench.py --gpus $N --steps 100 --warmup 10 --dist zipf --no-e2e > gpurun_out/${TAG}_bench_zipf.log 2> gpurun_out/${TAG}_bench_zipf.err
timeout 200 $TR --master-port 29564 bench.py --gpus $N --workload mot-proj-spt-64k --steps 30 --warmup 5 --no-e2e > gpurun_out/${TAG}_proj.log 2> gpurun_out/${TAG}_proj.err
timeout 200 $TR --master-port 29565 bench.py --gpus $N --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_ref.log 2> gpurun_out/${TAG}_ref.err
