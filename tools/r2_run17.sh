#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
run() {
  (cd xfeatslam_b200/csrc && touch conv_small.cu && make EXTRA="$1" > /dev/null 2>&1) || { echo "build $1 failed"; return; }
  timeout 200 python bench.py --no-cpu-baseline --chunks 4 --steps 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
  python - <<PY
import json
l=json.load(open("$O/bench.json"))
k=l["roofline"]["kernel_ms_per_batch"]
print("$1", "value", round(l["value"]), {n:round(v,4) for n,v in k.items() if n.startswith("block1")})
PY
}
run "-DXFB_B11_NW=8"
run "-DXFB_B11_NW=4"
run "-DXFB_B11_NW=4 -DXFB_B11_DIRECT=false"
run "-DXFB_B11_NW=8 -DXFB_B11_DIRECT=false"
run "-DXFB_B13_DIRECT=false -DXFB_B12_DIRECT=true"
(cd xfeatslam_b200/csrc && touch conv_small.cu && make > /dev/null 2>&1)
