#!/bin/bash
# 2-GPU sanity of the torchrun path with the final defaults (64-frame launch groups, 80 groups per step)
O=gpurun_out/r2n2; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench2.json 2> $O/bench2.err; echo "bench2 rc=$?"; tail -1 $O/bench2.json | cut -c1-900; tail -3 $O/bench2.err | cut -c1-200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/ref2.json 2> $O/ref2.err; echo "ref2 rc=$?"; tail -1 $O/ref2.json | cut -c1-300
