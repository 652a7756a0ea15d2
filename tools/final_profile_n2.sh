#!/bin/bash
# Two-GPU evidence (gpurun --gpus 2 -- 'bash tools/final_profile_n2.sh'): NCCL parity test, the bench line at N = 2 and
# -- after the same command exited 0 without ncu -- the launch list of rank 0 inside the 2-rank job.
T=${1:-r2f}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/${T}_dist_nccl_2gpu.log 2>&1; tail -2 $O/${T}_dist_nccl_2gpu.log
timeout 900 $TR --master-port 29511 bench.py --gpus 2 > $O/${T}_scale_2gpu.json 2> $O/${T}_scale_2gpu.err; echo "bench n2 rc=$?"
ARGS="bench.py --gpus 2 --steps 2 --warmup 3 --only --e2e-steps 0 --flag-rays 0 --cpu-rays 1000"
if timeout 300 $TR --master-port 29513 $ARGS > $O/${T}_n2_plain.log 2>&1; then
  timeout 600 $TR --no-python --master-port 29517 bash tools/rank0_ncu.sh $O/${T}_launches_c2_n2.csv $ARGS > $O/${T}_n2_ncu.log 2>&1; echo "ncu n2 rc=$?"
fi
ls -la $O | grep ${T}_
