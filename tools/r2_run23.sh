#!/bin/bash
# matcher ring-reuse ordering fix: the 5-stage build exposed the race (top-8192 case); both builds must pass now
O=gpurun_out/r2A; mkdir -p $O
(cd xfeatslam_b200/csrc && touch match_mutual.cu && make EXTRA="-DXFB_MM_STAGES=5" > /dev/null 2>&1) || echo "build 5 failed"
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_bench_path.py -q -x > $O/pytest5_$i.log 2>&1; echo "stages=5 run $i pytest rc=$?"; tail -1 $O/pytest5_$i.log; done
(cd xfeatslam_b200/csrc && touch match_mutual.cu && make > /dev/null 2>&1)
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_match.py tests/test_gpu_bench_path.py tests/test_gpu_host_dropin.py -q -x > $O/pytest4_$i.log 2>&1; echo "stages=4 run $i pytest rc=$?"; tail -1 $O/pytest4_$i.log; done
timeout 300 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.load(open("$O/bench.json"))
k=l["roofline"]["kernel_ms_per_batch"]
print("value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "match", k["match_tile"])
PY
