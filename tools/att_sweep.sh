#!/bin/bash
# timing of the tcgen05 attention kernel at the lang-config shapes, all softmax variants
for v in ${VARS:-0 1 2}; do
  echo "== SS_ATT_VAR=$v"
  SS_ATT_VAR=$v python tools/run_attention.py
  SS_ATT_VAR=$v ATT_N=120000 ATT_D=32 python tools/run_attention.py
  SS_ATT_VAR=$v ATT_H=2 ATT_D=16 python tools/run_attention.py
  SS_ATT_VAR=$v ATT_N=34000 ATT_H=16 ATT_D=16 python tools/run_attention.py
done
