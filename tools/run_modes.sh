# validation of the streaming matcher: frames-path probe, then the GPU test suite, then the bench
timeout 150 python tools/ms_probe.py > gpurun_out/s17_probe.log 2>&1; rc=$?; tail -12 gpurun_out/s17_probe.log
if [ $rc -ne 0 ]; then echo "probe failed rc=$rc -- skipping the rest"; exit 0; fi
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s17_pytest.log
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err || { echo "bench failed"; tail -3 gpurun_out/s17_bench.err; exit 0; }
python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(round(d['value']), round(d['e2e']['value']), {k:v for k,v in d['roofline']['kernel_ms_per_step'].items() if k.startswith('match')})" gpurun_out/s17_bench.json
