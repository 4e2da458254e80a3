# usage (on the GPU box): LIBS="a b" BATCHES="50000 65536" bash tools/_run_batches.sh  -- build_var/<lib>.so at several batch sizes
mkdir -p gpurun_out
for b in $BATCHES; do for v in $LIBS; do
  r=$(BNP_LIB=$PWD/build_var/$v.so timeout 300 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu --no-extras 2>gpurun_out/bt_$v.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4), d.get('parity',{}).get('ok'))" 2>&1 | tail -1)
  echo "batch $b $v : $r" | tee -a gpurun_out/batch_results.txt
done; done
