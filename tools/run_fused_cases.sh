#!/bin/bash
# every fused parity case in its own process
for kind in mass stiffness advection; do
for c in "3 35 47 20 dirichlet" "3 33 43 9 none" "3 40 12 17 left" "1 37 35 11 dirichlet" "1 5 4 6 none" "5 35 29 14 dirichlet" "5 13 12 27 none" "3 8 8 8 dirichlet"; do
  for lz in "" 8; do
    timeout 120 python tools/fused_case.py $c $kind $lz 2>&1 | grep -E "CASE|rror" | tail -2
  done
done
done
