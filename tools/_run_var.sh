# usage (on the GPU box): VARIANTS="a b" bash tools/_run_var.sh  -- benches build_var/<name>.so (kernel-only + e2e, parity-gated, no cpu leg)
mkdir -p gpurun_out
echo "--- $(date)" >> gpurun_out/var_results.txt
for v in ${VARIANTS:-$(ls build_var | sed 's/\.so$//')}; do
  r=$(BNP_LIB=$PWD/build_var/$v.so timeout 300 python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu --no-extras 2>gpurun_out/var_$v.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']), d.get('parity',{}).get('ok'))" 2>&1 | tail -1)
  echo "$v : $r" | tee -a gpurun_out/var_results.txt
done
