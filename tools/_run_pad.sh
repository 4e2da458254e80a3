for pad in 0 4096 10240 18432 30720; do
  r=$(BNP_SMEM_PAD=$pad python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/pad.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4))" 2>&1 | tail -1)
  echo "pad=$pad : $r" | tee -a gpurun_out/pad_results.txt
done
