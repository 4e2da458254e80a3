# BASELINE.json configs 3-5 at their full per-GPU sizes on one GPU (device-resident + e2e), no CPU leg
run() { # tag workload batch steps
  python bench.py --workload $2 --batch $3 --steps $4 --warmup 3 --no-cpu > gpurun_out/bench_cfg_$1.json 2> gpurun_out/bench_cfg_$1.err
  python -c "import json; d=json.load(open('gpurun_out/bench_cfg_$1.json')); print('$1', round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']), d['config']['batch_per_gpu'])" || tail -3 gpurun_out/bench_cfg_$1.err
}
run miller_2e20 miller 1048576 3
run final_exp_2e20 final_exp 1048576 3
run groth16_2e18 groth16 262144 3
run pairing_2e22 pairing 4194304 2
