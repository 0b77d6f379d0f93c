#!/bin/bash
# Device-timed stage micro-benchmarks over the BASELINE configurations (tools/bench_stage.py).
cd "$(dirname "$0")/.."
for wl in cfg1 cfg2 cfg3 cfg4; do
  for st in unpack idwt; do python tools/bench_stage.py $st --workload $wl --pictures 64 2>&1 | tail -1; done
done
# small pictures need larger batches to fill the GPU (a lane per slice: 32 x 24 x 148 slices resident)
for st in unpack idwt; do python tools/bench_stage.py $st --workload cfg3 --pictures 256 2>&1 | tail -1; done
for st in unpack idwt; do python tools/bench_stage.py $st --workload cfg1 --pictures 192 2>&1 | tail -1; done
for st in unpack idwt; do python tools/bench_stage.py $st --workload cfg2 --pictures 120 2>&1 | tail -1; done
# the int32 planes (vc2b_unpack / vc2b_idwt) on the headline configuration
for st in unpack idwt; do python tools/bench_stage.py $st --workload cfg2 --pictures 120 --int32-plane 2>&1 | tail -1; done
# encoder stages (cfg5's 12-bit coefficients need the int32 plane for the decode that sets the stage up)
for st in dwt quant pack; do python tools/bench_stage.py $st --workload cfg2 --pictures 64 2>&1 | tail -1; done
for st in dwt quant pack; do python tools/bench_stage.py $st --workload cfg5 --pictures 8 --int32-plane 2>&1 | tail -1; done
