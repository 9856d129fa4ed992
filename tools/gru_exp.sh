#!/bin/bash
# Timing experiments on the fused GRU recurrence (debug flags of ha2g_gru_seq_fwd_tc2_dbg, read from HA2G_GRU_DBGFLAGS):
#   0 = production path, 1 = skip the y / saved-gate copy-out, 3 = skip their shared-memory staging too.
# Results are only meaningful as timings (flags 1 and 3 do not write y / gates).
for f in 0 1 3; do
  echo "== flags $f"
  HA2G_GRU_DBGFLAGS=$f timeout 100 python tools/time_gru_tc.py 2>&1 | grep -E "tc2 gates|step period|mma issue|gate math|y/gates|wait for" | head -8
done
