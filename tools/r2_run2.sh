#!/bin/bash
# round 2, GPU call 2: first run of the persistent conv kernels (conv_tc2.cu): parity, then A/B timing against conv_tc.cu
O=gpurun_out/r2b; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_extract.py -m gpu -q -x -k "every_layer or dense_maps" > $O/pytest_layers.log 2>&1; echo "layers rc=$?"; tail -25 $O/pytest_layers.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest.log
timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench_v2.json 2> $O/bench_v2.err; echo "bench v2 rc=$?"; cut -c1-300 $O/bench_v2.json
XFB_CONV_TC=1 timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench_v1.json 2> $O/bench_v1.err; echo "bench v1 rc=$?"; cut -c1-300 $O/bench_v1.json
timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 --contexts 1 > $O/bench_v2_c1.json 2> $O/bench_v2_c1.err; echo "bench v2 c1 rc=$?"
timeout 200 python bench.py --no-cpu-baseline --chunks 4 --steps 5 --height 720 --width 1280 > $O/bench_v2_hd.json 2> $O/bench_v2_hd.err; echo "bench v2 hd rc=$?"; cut -c1-300 $O/bench_v2_hd.json
