for f in 0 4 8; do echo "== flags $f"; HA2G_GRU_DBGFLAGS=$f timeout 100 python tools/time_gru_tc.py 2>&1 | grep -E "tc2 gates|step period|mma issue|wait for|max \|y" | head -5; done
