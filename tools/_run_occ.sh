run() { # lib threads
  r=$(BNP_LIB=$PWD/$1 BNP_THREADS=$2 python bench.py --steps 6 --warmup 3 --no-cpu 2>gpurun_out/occ.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']))" 2>&1 | tail -1)
  echo "$1 threads=$2 : $r" | tee -a gpurun_out/occ_results.txt
}
run build_var/libbnp_s12_t96.so 96
run build_var/libbnp_s11_t64.so 64
run build_var/libbnp_s10_t64.so 64
run plonky2_bn254_pairing_b200/libbnp.so 64
