set -x
b() { # name, env...
  name=$1; shift
  r=$(env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-extras 2>gpurun_out/var_$name.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']), d.get('parity',{}).get('ok'))" 2>&1 | tail -1)
  echo "$name : $r" | tee -a gpurun_out/tmem_results.txt
}
echo "--- $(date)" >> gpurun_out/tmem_results.txt
if [ -n "$MAIN" ]; then b main; fi
for v in ${VARIANTS}; do
  b $v BNP_LIB=$PWD/build_var/$v.so
done
if [ -n "$MAIN" ]; then b main_again; fi
