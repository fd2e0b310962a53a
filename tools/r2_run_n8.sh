#!/bin/bash
# 8-GPU run: where does the end-to-end time go when 8 ranks share the host? (+ the scaling bench line)
O=gpurun_out/r2n8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --diagnose-e2e --chunks 8 --steps 5 --no-cpu-baseline > $O/diag8.json 2> $O/diag8.err; echo "diag8 rc=$?"; tail -1 $O/diag8.json | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --chunks 16 --steps 5 --no-cpu-baseline > $O/bench8.json 2> $O/bench8.err; echo "bench8 rc=$?"; tail -1 $O/bench8.json | cut -c1-700
timeout 200 python bench.py --diagnose-e2e --chunks 8 --steps 5 --no-cpu-baseline > $O/diag1.json 2> $O/diag1.err; echo "diag1 rc=$?"; tail -1 $O/diag1.json | cut -c1-900
