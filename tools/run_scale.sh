# usage: tools/run_scale.sh "<N list>" [parity]   (under gpurun --gpus N): the weak- and strong-scaling bench lines at each N;
# with "parity" also the multi-GPU parity scripts at test and bench size (planar strips / ranges / Delaunay, fracture network).
# Logs land in gpurun_out/ (copy what should be judged into profiles/).
for n in $1; do
  run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"
  if [ "$2" = "parity" ]; then
    echo "== parity N=$n"
    {
      $run tests/mgpu_check.py --mode weak 2>&1 | grep -E "^rank|Error|error|assert" | tail -$n
      $run tests/mgpu_check.py --mode strong --nx 96 --ny 80 2>&1 | grep -E "^rank|Error|error|assert" | tail -$n
      $run tests/mgpu_check.py --mode delaunay --nx 60 --ny 50 --rows-per-tile 48 2>&1 | grep -E "^rank|Error|error|assert" | tail -$n
      $run tests/mgpu_check.py --mode weak --nx 2048 --ny 1024 --rows-per-tile 336 2>&1 | grep -E "^rank|Error|error|assert" | tail -$n
      $run tests/mgpu_check.py --mode strong --nx 2048 --ny 1024 --rows-per-tile 336 2>&1 | grep -E "^rank|Error|error|assert" | tail -$n
      $run tests/mgpu_fracture_check.py 2>&1 | grep -E "^rank|Error|error|assert" | tail -$n
      $run tests/mgpu_fracture_check.py --nx 1024 --ny 586 --rows-per-tile 336 2>&1 | grep -E "^rank|Error|error|assert" | tail -$n
    } | tee gpurun_out/mgpu_parity_n$n.txt
  fi
  for sc in weak strong; do
    echo "== bench $sc N=$n"
    $run bench.py --gpus $n --steps 20 --warmup 5 --scaling $sc 2>gpurun_out/bench_${sc}_$n.err | tee gpurun_out/bench_${sc}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['scaling'], d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'el/s', '%.3e' % d['value'], 'e2e', '%.3e' % d['e2e']['value'], d['config'].get('exchange'))"
  done
done
