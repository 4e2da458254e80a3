for lib in build_var/libbnp_stag0.so plonky2_bn254_pairing_b200/libbnp.so build_var/libbnp_stag1500.so; do
 for b in 37888 65536; do
  r=$(BNP_LIB=$PWD/$lib python bench.py --batch $b --steps 6 --warmup 3 --no-cpu 2>gpurun_out/stag.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4))" 2>&1 | tail -1)
  echo "$lib batch=$b : $r" | tee -a gpurun_out/stag_results.txt
 done
done
