#!/bin/bash
# A/B timing of library variants: tools/ab.sh <config> <variant>...   ("main" = the in-tree library)
cfg=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then lib=""; else lib="wgsparkl_b200/_variants/lib_$v.so"; fi
  echo "== $v"
  B200MPM_VERBOSE=1 B200MPM_LIB=$lib timeout 120 python tools/run_config.py $cfg 3 2>&1 | grep -E "frame 2|update rigid|finite|k_g2p"
done
