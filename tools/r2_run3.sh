#!/bin/bash
# round 2, GPU call 3: where do the cycles of conv_tc2 go?  in-kernel role counters + one ncu --set full capture
O=gpurun_out/r2c; mkdir -p $O
XFB_T2_DEBUG=1 timeout 200 python bench.py --no-cpu-baseline --chunks 2 --steps 3 --contexts 1 > $O/bench_dbg.json 2> $O/bench_dbg.err; echo "bench dbg rc=$?"; grep "xfb t2" $O/bench_dbg.err | cut -c1-420
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 21 -c 8 -o $O/conv2 python bench.py --chunks 1 --steps 1 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu.log
ls -la $O
