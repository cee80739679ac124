#!/bin/bash
for lib in dynamicppr_b200/lib/libdppr.so build/variants/libdppr_u8b3.so build/variants/libdppr_u8b2.so build/variants/libdppr_u2b4.so build/variants/libdppr_u4b3.so; do
  for s in 1 4; do echo "=== $lib sources=$s"; DPPR_LIB=$PWD/$lib DPPR_CTAS_PER_SM=8 timeout 600 python scripts/probe.py --shape youtube --sources $s --batches 30 --show 0 2>&1 | grep -E "mean ms|per batch"; done
done
