# phase-cost probes of k_count_tc (OCBNN_COUNT_DBG bit mask, see svgd_tc.cu: count_dbg; 8 = the real kernel, gather skipped like in the other probes)
for d in ${COUNT_DBG_LIST:-8 9 10 12 15}; do echo "dbg=$d"; OCBNN_COUNT_DBG=$d timeout 60 python scripts/svgd_probe.py W4 ${COUNT_DBG_N:-32768} 2>&1 | grep -o "bandwidth [0-9.]* us"; done
