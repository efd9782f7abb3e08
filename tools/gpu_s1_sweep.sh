#!/bin/bash
for wl in ${WLS:-c1k1 c2k1}; do for R in ${RS:-64 128 256 512}; do CLOTHGPU_S1_ROWS=$R python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl R=$R', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; done; done
