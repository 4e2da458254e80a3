set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_split14.json 2> gpurun_out/bench_split.err; python -c "import json; d=json.load(open('gpurun_out/bench_split14.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'])"
PHASE_MODES="2 3" bash tools/_run_var.sh
