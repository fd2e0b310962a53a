#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
XFB_MS_DEBUG=64 timeout 200 python bench.py --no-cpu-baseline --chunks 2 --steps 3 --contexts 1 > $O/bench_msdbg.json 2> $O/bench_msdbg.err; echo "bench msdbg rc=$?"; grep "xfb" $O/bench_msdbg.err | cut -c1-700
