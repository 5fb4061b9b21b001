#!/bin/bash
# time the linearise stage (C3 shape) of every variants/*.so and the default build
for lib in default variants/*.so; do
  if [ "$lib" = default ]; then unset UWBGO_LIB; else export UWBGO_LIB=$PWD/$lib; fi
  out=$(timeout 120 python scripts/profile_solve.py --stage linearize --reps 5 2>&1 | grep -E "kernel ms" | tail -3 | tr '\n' ' ')
  echo "$lib :: $out"
done
