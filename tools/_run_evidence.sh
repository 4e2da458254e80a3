# usage (on the GPU box): TAG=r2_final bash tools/_run_evidence.sh
# the whole gpu test suite, smoke, the default bench line, the reference arm, the ncu launch list of the bench command and
# one --set full capture of the sequencer kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['parity']['ok'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnp_vm_kernel -c 1 -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu --no-extras > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
