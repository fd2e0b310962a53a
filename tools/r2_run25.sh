#!/bin/bash
O=gpurun_out/r2C; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extract.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.load(open("$O/bench.json"))
k=l["roofline"]["kernel_ms_per_batch"]
print("value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "pyramid", k["pyramid_sum"], "sum", round(sum(k.values()),3))
PY
