#!/bin/bash
# conv_tc2 raw input ring (TMA tensor loads): parity + per-kernel time
O=gpurun_out/r2s; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_bench_path.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest.log | cut -c1-300
timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
l=json.load(open("$O/bench.json"))
print("value", round(l["value"]), "e2e", round(l["e2e"]["value"]))
print({k:round(v,4) for k,v in l["roofline"]["kernel_ms_per_batch"].items()})
PY
