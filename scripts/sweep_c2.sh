#!/bin/bash
for rows in 8 4 2; do for xc in 12 16 24 32 48; do
  python bench.py --steps 100 --no-cpu-baseline --no-e2e --rows $rows --xchunk $xc 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rows=$rows xchunk=$xc', round(d['value'],2), 'E ms', round(d['roofline']['ms_per_launch'],4), 'H ms', round(d['roofline']['ms_per_launch_yee_H'],4))"
done; done
