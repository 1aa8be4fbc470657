#!/bin/bash
for b in 16 24 32; do echo -n "batch $b: "; python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), [round(q['ms'],2) for q in d['roofline_passes']])"; done
