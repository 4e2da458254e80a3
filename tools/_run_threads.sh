for t in ${THREADS:-64 256 512}; do
  for pm in ${PHASE_MODES:-1}; do
    r=$(BNP_THREADS=$t BNP_PHASE_MODE=$pm python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu 2>gpurun_out/thr.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['value']))" 2>&1 | tail -1)
    echo "threads=$t phase_mode=$pm : $r" | tee -a gpurun_out/thr_results.txt
  done
done
