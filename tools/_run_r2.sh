# usage (on the GPU box): TAG=name [NCU=1] [TESTS=1] bash tools/_run_r2.sh
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu --no-extras > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python -c "import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('${TAG}', round(d['value']), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['ms_per_step'])"
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnp_vm_kernel -c 1 -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu --no-extras > gpurun_out/ncu_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}.log
fi
