# usage: tools/run_scale.sh "<N list>"  -- weak-scaling bench lines + multi-GPU parity check
for n in $1; do
  echo "== mgpu_check N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 96 40 2>&1 | grep -a "rank\|Error\|error" | head -10
  echo "== bench N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_$n.err | tee gpurun_out/scale_$n.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('n',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
