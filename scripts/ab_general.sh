#!/bin/bash
# time the general LM kernel of every variants/*.so (and the default build) on C4a / C4b shapes
IFS=";" read -ra CFGS <<< "${AB_CFGS:-imu_lidar 8192 20 20;twist 8192 15 12;imu_lidar 65536 20 20}"
for cfg in "${CFGS[@]}"; do
  for lib in default variants/*.so; do
    set -- $cfg
    if [ "$lib" = default ]; then unset UWBGO_LIB; else export UWBGO_LIB=$PWD/$lib; fi
    out=$(timeout 120 python scripts/profile_solve.py --kind $1 --windows $2 --poses $3 --iters $4 --reps 3 2>&1 | grep -E "kernel ms|digest" | tail -2 | tr '\n' ' ')
    echo "$lib $cfg :: $out"
  done
done
