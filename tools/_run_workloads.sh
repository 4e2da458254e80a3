for w in miller final_exp groth16; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_wl_$w.json 2> gpurun_out/bench_wl_$w.err
  python -c "import json; d=json.load(open('gpurun_out/bench_wl_$w.json')); print('$w', round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']), d['config']['batch_per_gpu'])" || tail -3 gpurun_out/bench_wl_$w.err
done
# batch-size sweep of the headline workload (2^14 .. 2^20 per GPU)
for b in 16384 262144 1048576; do
  python bench.py --batch $b --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python -c "import json; d=json.load(open('gpurun_out/bench_b$b.json')); print('batch $b', round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']))" || tail -3 gpurun_out/bench_b$b.err
done
