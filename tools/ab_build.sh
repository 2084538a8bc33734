#!/bin/bash
# A/B of compile-time switches on ONE box: rebuild libagb.so without / with each given nvcc flag, alternately, and time the
# graph-replayed C2 step (tools/ab_graph.py) plus the attention op after every build.
#   tools/ab_build.sh [pairs] -DAGB_X -DAGB_Y=3 ...
cd "$(dirname "$0")/.."
PAIRS=$1; shift
for i in $(seq 1 $PAIRS); do
  for arm in default "$@"; do
    if [ "$arm" = default ]; then python -m affectgpt_b200.build --force > /dev/null; else AGB_NVCC_EXTRA=$arm python -m affectgpt_b200.build --force > /dev/null; fi
    echo "== build: $arm"
    timeout 300 python tools/ab_graph.py X=1 --rounds 16 2>&1 | tail -1
    timeout 100 python tools/attn_bench.py 2>&1 | tail -2
  done
done
python -m affectgpt_b200.build --force > /dev/null
