# usage: bash tools/_run_var.sh  -- benches every build_var/*.so (kernel-only + e2e, no cpu leg)
for lib in build_var/*.so; do
  for pm in ${PHASE_MODES:-1}; do
    r=$(BNP_LIB=$PWD/$lib BNP_PHASE_MODE=$pm python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu 2>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']))" 2>&1 | tail -1)
    echo "$lib phase_mode=$pm : $r" | tee -a gpurun_out/var_results.txt
  done
done
