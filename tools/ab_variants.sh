#!/bin/bash
# Builds differently compiled copies of the product library for A/B runs on the GPU box (YB_LIB_PATH selects one):
#   tools/ab_variants.sh name1 "-DFLAG=..." name2 "-D..." ...
set -e
cd "$(dirname "$0")/../yacrd_b200/csrc"
while [ $# -ge 2 ]; do
  make -s variant VARIANT="$1" EXTRA="$2" &
  shift 2
done
wait
ls -la ../libyacrd_b200_*.so
