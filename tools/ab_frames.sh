#!/bin/bash
# Frame-by-frame A/B of library variants: tools/ab_frames.sh <config> <frames> <variant>...   ("main" = in-tree)
cfg=$1; shift; frames=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then lib=""; else lib="wgsparkl_b200/_variants/lib_$v.so"; fi
  echo "== $v"
  B200MPM_LIB=$lib timeout 300 python tools/run_config.py $cfg $frames 2>&1 | grep -E "^frame" | awk '{printf "%s ", $5} END {print ""}' | tr -d '('
done
