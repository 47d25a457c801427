#!/bin/bash
for w in 2 4 8; do BOTGAT_ROWWISE_BWD=1 python tools/rank_bench.py --world $w; done
for w in 8; do BOTGAT_ROWWISE_BWD=1 BOTGAT_RW_BULK=0 python tools/rank_bench.py --world $w; done
for w in 8; do BOTGAT_ROWWISE=1 python tools/rank_bench.py --world $w; done
python bench.py --shape proteins --no-cpu-baseline --no-skew --no-e2e --no-parity --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('proteins', round(d['ms_per_step'],3), {k: v['avg_ms'] for k, v in d['kernels'].items()})"
