for lib in build_var/libbnp_ph8.so build_var/libbnp_ph6.so plonky2_bn254_pairing_b200/libbnp.so; do
 for b in 65536 16384; do
  r=$(BNP_LIB=$PWD/$lib python bench.py --batch $b --steps 6 --warmup 3 --no-cpu 2>gpurun_out/ph.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4))" 2>&1 | tail -1)
  echo "$lib batch=$b : $r" | tee -a gpurun_out/ph_results.txt
 done
done
