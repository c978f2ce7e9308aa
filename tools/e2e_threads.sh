#!/bin/bash
for t in 6 8 12 16 6; do
  echo "== OM_STAGE_THREADS=$t"
  OM_STAGE_THREADS=$t python tools/e2e_breakdown.py 2>&1 | grep -E "DeviceMesh|get points|get cells|optimize_points_cells" | tail -8 | awk '{printf "%s  ", $0} END {print ""}'
done
