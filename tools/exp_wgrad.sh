#!/bin/bash
# variants of the row-ring wgrad kernel (profiling knobs), row kernel only
for v in "" "B200_WGRAD_NSTACK=0" "B200_WGRAD_KHM1=1" "B200_WGRAD_DEBUG=1" "B200_WGRAD_DEBUG=2" "B200_WGRAD_DEBUG=3"; do
  echo "== variant: $v"
  env $v python tools/bench_wgrad.py 10 row
done
