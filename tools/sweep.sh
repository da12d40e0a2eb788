#!/bin/bash
# usage: tools/sweep.sh "<rows list>"   -- kernel time and roofline fraction per rows-per-tile
for r in $1; do
  python bench.py --steps 20 --warmup 5 --rows-per-tile $r --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); t=d['config']['tile_plan']
print('rows',$r,'kernel_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'tiles',t['tiles'],'halo',t['halo_factor'],'maxE',t['max_elem'],'maxV',t['max_vert'],'clk',d['clocks'])"
done
