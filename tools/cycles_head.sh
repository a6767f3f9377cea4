#!/bin/bash
# cycle-accurate A/B of head_fwd_kernel (clock independent): elapsed SM cycles + tensor-pipe activity
for cfg in "$@"; do
  ncu --metrics sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
      --clock-control none -k regex:head_fwd_kernel -s 2 -c 1 --csv python tools/profile_head_only.py $cfg 2>/dev/null \
      | grep -E "head_fwd" | awk -F'","' -v c="$cfg" '{printf "%s | %s = %s %s\n", c, $(NF-2), $NF, $(NF-1)}' | tr -d '"'
done
