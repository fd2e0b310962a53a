#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): tests, smoke, both bench arms, HD config, ncu launch list, ncu --set full of one
# launch group (summarised on the box: the .ncu-rep exceeds the 64 MiB return limit).
O=gpurun_out/final2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-300 $O/bench_n1.json
timeout 400 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"; cut -c1-200 $O/bench_ref.json
timeout 300 python bench.py --height 720 --width 1280 --chunks 32 --no-cpu-baseline > $O/bench_n1_hd.json 2> $O/bench_n1_hd.err; echo "bench hd rc=$?"; cut -c1-200 $O/bench_n1_hd.json
timeout 200 python bench.py --contexts 1 --no-cpu-baseline --chunks 16 > $O/bench_n1_ctx1.json 2> $O/bench_n1_ctx1.err; echo "bench ctx1 rc=$?"; cut -c1-200 $O/bench_n1_ctx1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --chunks 1 --steps 2 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none -s 160 -c 40 -o /tmp/step_full python bench.py --chunks 1 --steps 1 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/step_full.ncu-rep --page raw --csv > $O/step_full_raw.csv 2>/dev/null
du -sh $O
