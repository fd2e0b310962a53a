#!/bin/bash
# (1) matcher epilogue: share of the compares moved from the alu pipe to the fma pipe (XFB_MM_FMA_BITS) -- correctness + per-kernel time
# (2) evidence: both bench arms, HD config, ncu launch list, ncu --set full of one batch (summarised on the box: the .ncu-rep is too big to return)
O=gpurun_out/r2p; mkdir -p $O
run() {  # $1 = bits, $2 = run the parity tests
  (cd xfeatslam_b200/csrc && touch match_mutual.cu && make EXTRA=-DXFB_MM_FMA_BITS=$1 > /dev/null 2>&1) || { echo "build $1 failed"; return; }
  if [ "$2" = 1 ]; then timeout 300 python -m pytest tests/test_gpu_match.py tests/test_gpu_bench_path.py -q > $O/pytest_$1.log 2>&1; echo "bits=$1 pytest rc=$?"; tail -2 $O/pytest_$1.log | cut -c1-200; fi
  timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench_$1.json 2> $O/bench_$1.err; echo "bench rc=$?"
  python - <<PY
import json
l=json.load(open("$O/bench_$1.json"))
print("bits=$1 value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "match ms", {k:round(v,4) for k,v in l["roofline"]["kernel_ms_per_step"].items() if "match" in k or "mm" in k})
PY
}
run 0 0; run 26 0; run 30 1; run 21 1
(cd xfeatslam_b200/csrc && touch match_mutual.cu && make > /dev/null 2>&1)
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-300 $O/bench_n1.json
timeout 400 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"; cut -c1-200 $O/bench_ref.json
timeout 300 python bench.py --height 720 --width 1280 --chunks 32 --no-cpu-baseline > $O/bench_n1_hd.json 2> $O/bench_n1_hd.err; echo "bench hd rc=$?"; cut -c1-200 $O/bench_n1_hd.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --chunks 1 --steps 2 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none -s 160 -c 40 -o /tmp/step_full python bench.py --chunks 1 --steps 1 --warmup 3 --no-cpu-baseline --contexts 1 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py /tmp/step_full.ncu-rep > $O/step_full_summary.txt 2>&1
ncu -i /tmp/step_full.ncu-rep --page raw --csv > $O/step_full_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"nms_score_kernel|topk_kernel|describe_kernel" -s 6 -c 3 -o $O/post_full python bench.py --chunks 1 --steps 1 --warmup 1 --no-cpu-baseline --contexts 1 > $O/ncu_post.log 2>&1; echo "ncu post rc=$?"
du -sh $O; ls -la $O
