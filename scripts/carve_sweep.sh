#!/bin/bash
# Developer sweep: block-kernel time vs shared-memory carve-out (percent of 228 KB) for the library and its _variants builds.
for so in "" $@; do
  for c in -1 100 86 72 58 44; do
    if [ "$c" = "-1" ]; then unset PCC_THR_CARVEOUT; else export PCC_THR_CARVEOUT=$c; fi
    echo -n "carve=$c "; PCC_SO=$so python scripts/probe.py 10000000 surface 16 0 2>&1 | tail -1
  done
done
