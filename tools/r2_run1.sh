#!/bin/bash
# round 2, GPU call 1: the new parity tests + baselines (VGA sustained, HD) before the kernel work
O=gpurun_out/r2a; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 400 python bench.py --no-cpu-baseline > $O/bench_vga.json 2> $O/bench_vga.err; echo "bench rc=$?"; cut -c1-400 $O/bench_vga.json
timeout 400 python bench.py --no-cpu-baseline --height 720 --width 1280 --chunks 24 > $O/bench_hd.json 2> $O/bench_hd.err; echo "bench hd rc=$?"; cut -c1-400 $O/bench_hd.json
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; lscpu | head -30 >> $O/topo.txt
