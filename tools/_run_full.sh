# usage (on the GPU box): TAG=name bash tools/_run_full.sh  -- the whole gpu test suite, smoke, the default bench line and the reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -2 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['parity'], d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_ref_${TAG}.json | cut -c1-400
