timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/s21_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s21_pytest.log
timeout 100 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err || { echo "bench failed"; tail -3 gpurun_out/s21_bench.err; }
python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(round(d['value']), round(d['e2e']['value']), d['gpu_launches'], {k:v for k,v in d['roofline']['kernel_ms_per_step'].items() if k.startswith('match') or k.startswith('block3') or k.startswith('block2')})" gpurun_out/s21_bench.json
