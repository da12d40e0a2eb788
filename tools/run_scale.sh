# usage: tools/run_scale.sh "<N list>"   (under gpurun --gpus N): multi-GPU parity at test and bench size, then
# the weak- and strong-scaling bench lines at each N.  Logs land in gpurun_out/.
for n in $1; do
  run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"
  echo "== parity N=$n (small: weak, strong, delaunay; bench size: weak 2048x1024 per rank, strong 2048x1024)"
  $run tests/mgpu_check.py --mode weak 2>&1 | grep -E "rank|Error|error|assert" | tail -$n
  $run tests/mgpu_check.py --mode strong --nx 96 --ny 80 2>&1 | grep -E "rank|Error|error|assert" | tail -$n
  $run tests/mgpu_check.py --mode delaunay --nx 60 --ny 50 --rows-per-tile 48 2>&1 | grep -E "rank|Error|error|assert" | tail -$n
  $run tests/mgpu_check.py --mode weak --nx 2048 --ny 1024 --rows-per-tile 336 2>&1 | grep -E "rank|Error|error|assert" | tail -$n
  $run tests/mgpu_check.py --mode strong --nx 2048 --ny 1024 --rows-per-tile 336 2>&1 | grep -E "rank|Error|error|assert" | tail -$n
  echo "== bench weak N=$n"
  $run bench.py --gpus $n --steps 20 --warmup 5 2>/dev/null | tee gpurun_out/bench_weak_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['scaling'], d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'el/s', '%.3e' % d['value'], 'e2e', '%.3e' % d['e2e']['value'])"
  echo "== bench strong N=$n"
  $run bench.py --gpus $n --steps 20 --warmup 5 --scaling strong 2>/dev/null | tee gpurun_out/bench_strong_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['scaling'], d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'el/s', '%.3e' % d['value'], 'e2e', '%.3e' % d['e2e']['value'])"
done
