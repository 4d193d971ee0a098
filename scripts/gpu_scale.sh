#!/bin/bash
# bench.py under torchrun at N ranks (N = $1), both arms, as the driver launches them.
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/scale_ref_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/scale_n$N.err | tail -1 > gpurun_out/scale_n$N.json
cat gpurun_out/scale_n$N.json | cut -c1-260; tail -3 gpurun_out/scale_n$N.err
