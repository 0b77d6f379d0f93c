#!/bin/bash
# Device-timed stage micro-benchmarks over the BASELINE configurations (tools/bench_stage.py).
cd "$(dirname "$0")/.."
for wl in cfg1 cfg2 cfg3 cfg4; do
  for st in unpack idwt; do python tools/bench_stage.py $st --workload $wl --pictures 64 2>&1 | tail -1; done
done
# small pictures need larger batches to fill the GPU: the LD path is latency-bound at 64 fields
# (DC prediction alone is a dependent chain of ~0.13 ms whatever the number of fields)
for st in unpack idwt; do python tools/bench_stage.py $st --workload cfg3 --pictures 256 2>&1 | tail -1; done
for st in unpack idwt; do python tools/bench_stage.py $st --workload cfg1 --pictures 192 2>&1 | tail -1; done
for wl in cfg2 cfg5; do
  for st in dwt quant pack; do python tools/bench_stage.py $st --workload $wl --pictures 8 2>&1 | tail -1; done
done
