#!/bin/bash
for lib in libdppr.so libdppr_b1.so libdppr_b2.so; do
for args in "--shape youtube --batches 50" "--shape livejournal --scale 0.25 --batches 10"; do
  echo "=== $lib dense-kernel, never entering: $args"; DPPR_LIB=$PWD/dynamicppr_b200/lib/$lib DPPR_DENSE_DIV=1e-6 DPPR_DENSE_MIN_EDGES=0 timeout 120 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
done; done
echo "=== reference: non-dense kernel"; for args in "--shape youtube --batches 50" "--shape livejournal --scale 0.25 --batches 10"; do DPPR_DENSE_DIV=0 timeout 120 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"; done
