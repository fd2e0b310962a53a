#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
XFB_MS_DEBUG=1 timeout 200 python bench.py --no-cpu-baseline --batch 32 --chunks 2 --steps 3 --contexts 1 > $O/bench_dbg.json 2> $O/bench_dbg.err; echo "rc=$?"; grep "xfb" $O/bench_dbg.err | grep -v "match_stream\|mma thread\|CTA(0,0)" | cut -c1-700 | head
