#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_match.py tests/test_gpu_bench_path.py tests/test_gpu_host_dropin.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log | cut -c1-300
XFB_MS_DEBUG=1 timeout 200 python bench.py --no-cpu-baseline --batch 32 --chunks 2 --steps 3 --contexts 1 > $O/bench_dbg.json 2> $O/bench_dbg.err; echo "rc=$?"; grep "xfb" $O/bench_dbg.err | grep -v "match_stream\|mma thread\|CTA(0,0)" | cut -c1-600 | head
timeout 300 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
python - <<PY
import json
l=json.load(open("$O/bench.json"))
k=l["roofline"]["kernel_ms_per_batch"]
print("value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "sum ms/batch", round(sum(k.values()),3), "match", k["match_tile"])
PY
