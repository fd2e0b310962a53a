#!/bin/bash
# 2-GPU run of the 1280x720 configuration (BASELINE.json configs[3] at the N this budget allows)
O=gpurun_out/r2n2hd; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --height 720 --width 1280 --chunks 16 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench2hd.json 2> $O/bench2hd.err; echo "rc=$?"; tail -1 $O/bench2hd.json | cut -c1-400
