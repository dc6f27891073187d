#!/bin/bash
# Developer helper: ncu launch list of the k=16 headline call for each given build; prints the block / retry / ring kernel times.
for so in "$@"; do
  PCC_SO=$so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/kt.csv python scripts/probe.py 10000000 surface 16 0 > /dev/null 2>&1
  echo "== $so"; python scripts/launch_times.py /tmp/kt.csv 40 | grep -E "knn_thr|knn_rings|knn_fast" | tail -3
done
