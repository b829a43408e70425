for s in 0.5 0.58 0.63 0.7 0.8; do
  echo "shrink $s"
  WTP_GRID2_SHRINK=$s timeout 300 python scripts/graded_probe.py 2>&1 | tail -1 | python -c "
import sys, ast
d=ast.literal_eval(sys.stdin.read()); print({k:d[k] for k in ('ms_query','ms_sort','n_ring_expanded','n_leftover_sparse','n_leftover_dense','n_leftover_other')})"
done
