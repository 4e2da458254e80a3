set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; cat gpurun_out/bench_head.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_head.json 2>> gpurun_out/bench_head.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_head.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnp_vm_kernel -c 1 -o gpurun_out/prof_vm_head -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
