XFB_MATCH_IMPL=1 timeout 100 python -m pytest tests/test_gpu_match.py tests/test_gpu_host_dropin.py -m gpu -q -x 2>&1 | tail -3
for cfg in "1 -" "1 0"; do set -- $cfg
  if [ "$2" = "-" ]; then unset XFB_MS_DEBUG; else export XFB_MS_DEBUG=$2; fi
  XFB_MATCH_IMPL=$1 timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s16_bench_$1_$2.json 2> gpurun_out/s16_bench_$1_$2.err || { echo "bench $cfg failed"; tail -3 gpurun_out/s16_bench_$1_$2.err; continue; }
  grep xfb gpurun_out/s16_bench_$1_$2.err
  python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[2], sys.argv[3], round(d['value']), round(d['e2e']['value']), {k:v for k,v in d['roofline']['kernel_ms_per_step'].items() if k.startswith('match')})" gpurun_out/s16_bench_$1_$2.json $1 $2; done
