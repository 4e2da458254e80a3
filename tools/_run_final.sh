# round-2 evidence run (on the GPU box): tests, the driver's bench invocation, the reference arm, the ncu launch list of the
# bench command and one --set full capture of the dominant kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -2 gpurun_out/bench_r2.err; head -c 600 gpurun_out/bench_r2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2.json 2>> gpurun_out/bench_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnp_vm_kernel -c 1 -o gpurun_out/prof_r2 -f python bench.py --steps 1 --warmup 0 --no-cpu --no-extras > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python -c "import __graft_entry__ as g; g.smoke()"
