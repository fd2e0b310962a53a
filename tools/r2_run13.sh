#!/bin/bash
# post-processing rewrite (separable NMS + compacted scoring, register-resident top-k sort, 2 keypoints / warp describe, coalesced matcher images)
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log | cut -c1-300
timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.load(open("$O/bench.json"))
print("value", round(l["value"]), "e2e", round(l["e2e"]["value"]))
print({k:round(v,4) for k,v in l["roofline"]["kernel_ms_per_batch"].items()})
PY
