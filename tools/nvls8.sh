TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
MOT_AR_SWEEP=1 timeout 150 $TR tools/test_nvls.py 2>&1 | grep -vE "^\*|OMP_NUM"
timeout 150 $TR bench.py --gpus 8 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench8_own.log 2>&1; tail -1 gpurun_out/bench8_own.log | cut -c1-1100
MOT_DP_NCCL=1 timeout 150 $TR bench.py --gpus 8 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench8_nccl.log 2>&1; tail -1 gpurun_out/bench8_nccl.log | cut -c1-300
