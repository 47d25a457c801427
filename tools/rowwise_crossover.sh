#!/bin/bash
# Where the all-heads-per-row kernels pay: head-major (BOTGAT_ROWWISE=0) against all-heads-per-row (=1)
#   usage: bash tools/rowwise_crossover.sh size     products-like layer (H = 4, D = 120, 25 edges per node), growing node count
#          bash tools/rowwise_crossover.sh degree   fixed node count, growing average degree, products and proteins geometry
run() {  # shape nodes edges
  for rw in 0 1; do
    BOTGAT_ROWWISE=$rw python bench.py --shape $1 --nodes $2 --edges $3 --no-cpu-baseline --no-skew --no-e2e --no-parity --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernels']; print('$1', 'nodes', $2, 'deg', $3//$2, 'rowwise', $rw, 'step', round(d['ms_per_step'],3), 'fwd', k['gat_fwd']['avg_ms'], 'src', k['gat_bwd_src']['avg_ms'])"
  done
}
if [ "${1:-size}" = "size" ]; then
  for n in 150000 300000 450000 600000 900000; do run products $n $((n * 25)); done
else
  for deg in 60 120 240; do run products 300000 $((300000 * deg)); done
  for deg in 25 75 150; do run proteins 132534 $((132534 * deg)); done
fi
