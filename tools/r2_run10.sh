#!/bin/bash
# ncu --set full with source for the three hot tcgen05 kernels of the current build
O=gpurun_out/r2j; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mm_kernel|conv_tc2_kernel" -s 22 -c 9 -o $O/hot python bench.py --chunks 1 --steps 1 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mm_kernel" -s 1 -c 1 -o $O/mm python bench.py --chunks 1 --steps 1 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la $O
