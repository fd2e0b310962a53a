#!/bin/bash
# FFMA2 in block1 + programmatic dependent launch of the conv_tc2 layers
O=gpurun_out/r2v; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_bench_path.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest.log | cut -c1-300
for pdl in 0 1; do
XFB_PDL=$pdl timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 > $O/bench_$pdl.json 2> $O/bench_$pdl.err; echo "bench pdl=$pdl rc=$?"; tail -2 $O/bench_$pdl.err
python - <<PY
import json
l=json.load(open("$O/bench_$pdl.json"))
k=l["roofline"]["kernel_ms_per_batch"]
print("pdl=$pdl value", round(l["value"]), "e2e", round(l["e2e"]["value"]), {n:round(v,4) for n,v in k.items() if n.startswith("block1")})
PY
done
XFB_PDL=1 timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 --contexts 1 > $O/bench_1c.json 2>/dev/null; python -c "
import json; print('pdl=1 ctx1', round(json.load(open('$O/bench_1c.json'))['value']))"
XFB_PDL=0 timeout 200 python bench.py --no-cpu-baseline --chunks 8 --steps 5 --contexts 1 > $O/bench_0c.json 2>/dev/null; python -c "
import json; print('pdl=0 ctx1', round(json.load(open('$O/bench_0c.json'))['value']))"
