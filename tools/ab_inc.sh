# same-box A/B of the incremental kernel (libbase.so vs the current build): cfg2, cfg3 and a 25,000 x 30 kb input
for L in libbase libodmsspe_b200; do echo $L; export MSSPE_LIB=/root/repo/open-msspe-design_b200/$L.so
  python tools/dbg_phases.py cfg2 2>&1 | grep -E "incremental" | tail -1 | cut -c1-230
  python tools/dbg_phases.py cfg3 2>&1 | grep -E "incremental" | tail -1 | cut -c1-230
  MSSPE_DEBUG_TIMERS=1 python tools/prof_modes.py 25000 1000 2>&1 | grep -E "incremental|identical" | tail -3 | cut -c1-230
done
