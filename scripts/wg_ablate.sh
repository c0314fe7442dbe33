#!/bin/bash
# k_wgrad_tc under its ablation switches (B2NO_WG_DEBUG: 1 no X-lo pass, 2 no G conversion, 4 no MMAs)
for d in 0 1 2 4 3 7; do
  echo "== B2NO_WG_DEBUG=$d"; B2NO_WG_DEBUG=$d timeout 120 python scripts/prof_layer.py time 2>&1 | grep -E "^wgrad"
done
