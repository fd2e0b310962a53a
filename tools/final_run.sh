#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): tests, smoke, bench, ncu launch list, ncu full captures, microbenchmarks.
O=gpurun_out/final; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-600 $O/bench_n1.json
timeout 100 python bench.py --contexts 1 --no-cpu-baseline > $O/bench_n1_ctx1.json 2> $O/bench_n1_ctx1.err; echo "bench ctx1 rc=$?"
XFB_MS_DEBUG=0 timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --contexts 1 2> $O/ms_debug.err > /dev/null; grep xfb $O/ms_debug.err
timeout 60 tools/tmem_bench > $O/tmem_bench.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 37 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ms_kernel -s 2 -c 2 -o $O/ms_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --contexts 1 > $O/ncu_ms.log 2>&1; echo "ncu ms rc=$?"
timeout 240 ncu --set full --clock-control none -k regex:"conv_tc_kernel|conv_small_kernel" -s 25 -c 25 -o $O/conv_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --contexts 1 > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ls -la $O
