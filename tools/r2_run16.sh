#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
XFB_T2_DEBUG=1 timeout 200 python bench.py --no-cpu-baseline --chunks 2 --steps 3 --contexts 1 > $O/bench_dbg.json 2> $O/bench_dbg.err; echo "bench dbg rc=$?"; grep "xfb t2" $O/bench_dbg.err | cut -c1-420 | head -24
