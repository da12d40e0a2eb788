run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('kernel_ms',round(d['roofline']['kernel_ms'],4))"; }
echo "== default"; run
for v in $VARIANTS; do echo "== $v"; TFEM_B200_LIB=variants/lib$v.so run; done
