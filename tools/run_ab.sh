n=$1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/ab.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],d['config'].get('exchange'))"; }
echo "== no exchange, reserve 0";  TFEM_DEBUG_NO_EXCHANGE=1 TFEM_RESERVE_CTAS=0 run
echo "== no exchange, reserve 24"; TFEM_DEBUG_NO_EXCHANGE=1 TFEM_RESERVE_CTAS=24 run
echo "== peer single, reserve 24"; run
echo "== peer single, reserve 8"; TFEM_RESERVE_CTAS=8 run
echo "== peer single, reserve 48"; TFEM_RESERVE_CTAS=48 run
