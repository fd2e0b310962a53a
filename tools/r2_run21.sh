#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_match.py tests/test_gpu_bench_path.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
python - <<PY
import json
l=json.load(open("$O/bench.json"))
k=l["roofline"]["kernel_ms_per_batch"]
print("value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "sum ms/batch", round(sum(k.values()),3), "match", k["match_tile"], "prep_stats", k["prep_stats"])
PY
