#!/bin/bash
# Two-GPU evidence (gpurun --gpus 2 -- 'bash tools/final_profile_n2.sh'): NCCL parity test and the bench line at N = 2.
T=${1:-r2f}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/${T}_dist_nccl_2gpu.log 2>&1; tail -2 $O/${T}_dist_nccl_2gpu.log
timeout 900 $TR --master-port 29511 bench.py --gpus 2 > $O/${T}_scale_2gpu.json 2> $O/${T}_scale_2gpu.err; echo "bench n2 rc=$?"
# (An ncu launch-list pass of rank 0 inside the 2-rank job was tried once -- rank 0 under `ncu --metrics
#  gpu__time_duration.sum`, rank 1 plain -- and hung before the first kernel until its 600 s timeout, 23 GPU-minutes for no
#  data; the roofline at N > 1 therefore comes from the bench line's own device timers, which is what the contract asks.)
ls -la $O | grep ${T}_
