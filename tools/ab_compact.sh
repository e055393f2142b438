# how many stream compactions pay at a configuration: MSSPE_COMPACT_MIN sweep (recount kernel)
cfg=${1:-cfg2}
for cm in 1048576 2097152 4194304 8388608 33554432; do echo "COMPACT_MIN $cm"; MSSPE_COMPACT_MIN=$cm python tools/dbg_phases.py $cfg 2>&1 | grep -E "^mode 0|compactions" | tail -2; done
