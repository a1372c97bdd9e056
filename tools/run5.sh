set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
MOT_AR_SWEEP=1 timeout 200 $TR tools/test_nvls.py 2>&1 | grep -vE "^\*|OMP_NUM" 
timeout 200 $TR bench.py --gpus 2 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench2_own.log 2>&1; tail -1 gpurun_out/bench2_own.log | cut -c1-900
MOT_DP_NCCL=1 timeout 200 $TR bench.py --gpus 2 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench2_nccl.log 2>&1; tail -1 gpurun_out/bench2_nccl.log | cut -c1-400
