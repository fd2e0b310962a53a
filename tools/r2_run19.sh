#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_match.py tests/test_gpu_bench_path.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log | cut -c1-300
timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.load(open("$O/bench.json"))
print("value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "match", l["roofline"]["kernel_ms_per_batch"]["match_tile"])
PY
# memory checker on the small-shape extract + match tests (new kernels of this round: NMS, top-k, describe, raw ring, drains)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_extract.py -q -x -k "every_layer or edge_cases or odd_sizes or topk_8192" > $O/memcheck_extract.log 2>&1; echo "memcheck extract rc=$?"; tail -4 $O/memcheck_extract.log | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_match.py -q -x > $O/memcheck_match.log 2>&1; echo "memcheck match rc=$?"; tail -4 $O/memcheck_match.log | cut -c1-200
