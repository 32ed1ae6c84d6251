#!/bin/bash
# tools/build_variants.sh name "-DXN_..." [name "-D..."]...  -> build/variants/libxnart_<name>.so (tuning experiments; XNART_LIB selects one)
cd "$(dirname "$0")/../xna-ray-trace_b200/csrc"
mkdir -p ../../build/variants
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  ( make -B OUT=../../build/variants/libxnart_$n.so LOG=../../build/variants/$n.log EXTRA="$f" > /dev/null 2>&1; grep -A2 "k_pathILb0" ../../build/variants/$n.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $n" ) &
done
wait
