#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest.log | cut -c1-300
XFB_MS_DEBUG=64 timeout 200 python bench.py --no-cpu-baseline --chunks 2 --steps 3 --contexts 1 > $O/bench_dbg.json 2> $O/bench_dbg.err; echo "bench dbg rc=$?"; grep "match_mutual" $O/bench_dbg.err | cut -c1-700
timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench_v2.json 2> $O/bench_v2.err; echo "bench v2 rc=$?"; cut -c1-200 $O/bench_v2.json
XFB_B1_FUSE=0 timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench_nofuse.json 2> $O/bench_nofuse.err; echo "bench nofuse rc=$?"; cut -c1-200 $O/bench_nofuse.json
timeout 200 python bench.py --no-cpu-baseline --chunks 4 --steps 5 --height 720 --width 1280 > $O/bench_v2_hd.json 2> $O/bench_v2_hd.err; echo "bench v2 hd rc=$?"; cut -c1-200 $O/bench_v2_hd.json
