#!/bin/bash
# per-kernel durations (ncu, serialised) of the tile kernel under the B2NO_TC_DEBUG ablation switches
for d in ${DBG_LIST:-0 1 2 3}; do
  B2NO_TC_DEBUG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pw_tc|k_inv_h" -s 4 -c 2 --csv \
     python scripts/prof_layer.py ${DBG_WHAT:-inv} 3 2>/dev/null | grep -E "k_pw_tc|k_inv_h" | awk -F'","' -v d=$d '{printf "debug=%s %s %s us\n", d, substr($5,1,40), $NF/1000}' | tr -d '"'
done
