#!/bin/bash
# A/B timing of environment switches: tools/ab_env.sh <config> "VAR=1" "VAR2=x" ...   ("-" = no extra variable)
cfg=$1; shift
for v in "$@"; do
  echo "== $v"
  if [ "$v" = "-" ]; then python tools/run_config.py $cfg 3 2>&1 | grep -E "frame 2|update rigid"; else env $v python tools/run_config.py $cfg 3 2>&1 | grep -E "frame 2|update rigid"; fi
done
