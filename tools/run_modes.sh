timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/s19_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s19_pytest.log
for nc in 2; do
timeout 100 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --contexts $nc > gpurun_out/s19_bench_$nc.json 2> gpurun_out/s19_bench_$nc.err || { echo "bench $nc failed"; tail -3 gpurun_out/s19_bench_$nc.err; continue; }
python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[2], round(d['value']), round(d['e2e']['value']), d['gpu_launches'], d['roofline']['kernel_ms_per_step'])" gpurun_out/s19_bench_$nc.json $nc
done
