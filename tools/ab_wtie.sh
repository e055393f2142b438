# A/B of the warp tie-score list-length threshold (recount: MSSPE_WTIE_MAX, incremental: MSSPE_WTIE_MAX_INC)
cfg=${1:-cfg2}
for w in 0 128 256 1024 4096; do echo "WTIE $w"; MSSPE_WTIE_MAX=$w MSSPE_WTIE_MAX_INC=$w python tools/dbg_phases.py $cfg 2>&1 | grep -E "^mode|worker|persistent|incremental" | sed -n "5,7p;10,11p"; done
