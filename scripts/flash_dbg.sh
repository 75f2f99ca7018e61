# phase-cost probes of k_phi_flash (OCBNN_FLASH_DBG bit mask, see svgd_tc.cu: flash_dbg)
for d in ${FLASH_DBG_LIST:-0 7 1}; do echo "dbg=$d"; OCBNN_FLASH_DBG=$d timeout 120 python scripts/svgd_probe.py W4 ${FLASH_DBG_N:-32768} 2>&1 | grep -o "phi [0-9.]* us"; done
