#!/bin/bash
set -u
mkdir -p gpurun_out
for sh in 256x8 512x4 128x16; do
  COMBAT_TC_DBG=1 timeout 300 python scripts/bench_conv.py --only $sh --variant "fwd bn only" > gpurun_out/dbg_pair_$sh.txt 2>&1
  COMBAT_TC_DBG=1 timeout 300 python scripts/bench_conv.py --only $sh --variant "fwd res+out+bn" >> gpurun_out/dbg_pair_$sh.txt 2>&1
  COMBAT_TC_DBG=1 COMBAT_NO_PAIR=1 timeout 300 python scripts/bench_conv.py --only $sh --variant "fwd bn only" > gpurun_out/dbg_nopair_$sh.txt 2>&1
  cat gpurun_out/dbg_pair_$sh.txt gpurun_out/dbg_nopair_$sh.txt
done
COMBAT_TC_DBG=1 timeout 300 python scripts/bench_conv.py --only 64x32 --variant "fwd bn only"
