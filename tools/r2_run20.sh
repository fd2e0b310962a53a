#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extract.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log | cut -c1-300
for bs in 32 64 96; do
ch=$((256 / bs)); if [ $ch -lt 2 ]; then ch=2; fi
timeout 300 python bench.py --no-cpu-baseline --batch $bs --chunks $ch --steps 5 > $O/bench_$bs.json 2> $O/bench_$bs.err; echo "bench batch=$bs rc=$?"; tail -2 $O/bench_$bs.err
python - <<PY
import json
l=json.load(open("$O/bench_$bs.json"))
k=l["roofline"]["kernel_ms_per_batch"]
print("batch=$bs value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "sum ms/batch", round(sum(k.values()),3), "match", k["match_tile"], "heatmap_out", k["heatmap_out"])
PY
done
