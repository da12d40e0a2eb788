run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('kernel_ms',round(d['roofline']['kernel_ms'],4))"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== default"; run
for r in 176 192; do echo "== default rows $r"; run --rows-per-tile $r; done
for v in $VARIANTS; do echo "== $v"; TFEM_B200_LIB=variants/lib$v.so run; done
echo "== timing"; TFEM_B200_LIB=variants/libtiming.so python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e 2>/dev/null | grep -a "^cta" | head -6
