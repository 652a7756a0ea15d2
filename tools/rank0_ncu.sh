#!/bin/bash
# Launch-list pass at N > 1: local rank 0 runs under `ncu --metrics gpu__time_duration.sum` (no replay, so the other
# ranks and NCCL are not disturbed beyond the serialisation of rank 0's kernels); the other ranks run plainly.
#   python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
#       bash tools/rank0_ncu.sh gpurun_out/r2f_launches_n2.csv bench.py --gpus 2 --steps 2 --warmup 3 --only ...
OUT="$1"; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT" python "$@"
else
  exec python "$@"
fi
