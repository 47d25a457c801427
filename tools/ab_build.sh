#!/bin/bash
# A/B of two builds of the library on the same box: bash tools/ab_build.sh variants/libX.so [shape]
var=$1; shape=${2:-proteins}
for i in 1 2; do for lib in $var ""; do
BOTGAT_LIB=${lib:+$PWD/$lib} python bench.py --shape $shape --no-cpu-baseline --no-skew --no-e2e --no-parity --steps 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$shape', '$lib' or 'current', round(d['ms_per_step'],3), {k: v['avg_ms'] for k, v in d['kernels'].items()})"
done; done
