for b in 37888 65536 75776 113664 1048576; do
  for pm in 2 3; do
    r=$(BNP_PHASE_MODE=$pm python bench.py --batch $b --steps 5 --warmup 3 --no-cpu 2>gpurun_out/tail.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4))" 2>&1 | tail -1)
    echo "batch=$b phase_mode=$pm : $r" | tee -a gpurun_out/tail_results.txt
  done
done
