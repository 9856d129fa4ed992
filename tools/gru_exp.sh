for f in 0 1 3; do echo "== flags $f"; HA2G_GRU_DBGFLAGS=$f timeout 100 python tools/time_gru_tc.py 2>&1 | head -11; done
