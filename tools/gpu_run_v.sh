#!/bin/bash
# round-2 session v (2 GPUs): why is the overlapped 2-GPU apply slow?  NCCL kernel footprint, serialised path, face slots
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/v_bench.log
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --quick --steps 300 --warmup 10 >> gpurun_out/v_bench.log 2>/dev/null
}
run A=0
run GDM_FUSED_OVERLAP=0
run NCCL_NTHREADS=128 NCCL_MAX_NCHANNELS=2
run NCCL_NTHREADS=64 NCCL_MAX_NCHANNELS=1
run NCCL_NTHREADS=128 NCCL_MAX_NCHANNELS=2 GDM_PERS_FACE_SLOTS=16
run NCCL_NTHREADS=128 NCCL_MAX_NCHANNELS=2 GDM_FUSED_OVERLAP=0
run GDM_PERS_FACE_SLOTS=40
