#!/bin/bash
# tools/ab2.sh <script.py> <variant>... : runs a diagnostic script against library variants
scr=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then lib=""; else lib="wgsparkl_b200/_variants/lib_$v.so"; fi
  echo "== $v"
  B200MPM_LIB=$lib python $scr 2>&1 | tail -9
done
