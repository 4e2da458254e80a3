set -x
timeout 900 python -m pytest tests/test_gpu_prepared.py tests/test_gpu_wire.py -x -q > gpurun_out/pytest_prep.log 2>&1; tail -15 gpurun_out/pytest_prep.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_prep.json 2> gpurun_out/bench_prep.err; tail -3 gpurun_out/bench_prep.err; python -c "
import json; d=json.load(open('gpurun_out/bench_prep.json')); print(round(d['value']), d['roofline']['frac']); print(d['extra']['groth16_prepared_2e18']); print(d['extra']['workloads']['groth16_2e18'])"
