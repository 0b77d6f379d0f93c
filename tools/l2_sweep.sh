#!/bin/bash
# IDWT L2-residency experiment: sub-group size of the last two levels x eviction hints (cfg2, 120 pictures)
cd "$(dirname "$0")/.."
for H in 0 1 3; do
  for G in 0 1 2 3 4 6 8 12; do
    echo -n "hints=$H group=$G: "
    VC2B_IDWT_L2_HINTS=$H VC2B_IDWT_L2_GROUP=$G python tools/bench_stage.py idwt --workload cfg2 --pictures 120 --iters 20 | tail -1
  done
done
