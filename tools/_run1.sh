set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for m in 2 3 1; do BNP_PHASE_MODE=$m python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_phase$m.json 2> gpurun_out/bench_phase$m.err; cat gpurun_out/bench_phase$m.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'])"; done
BNP_PHASE_MODE=3 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
