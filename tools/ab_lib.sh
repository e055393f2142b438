# same-box A/B of two builds of the library: tools/ab_lib.sh <cfg>  (libbase.so = the build to compare against)
cfg=${1:-cfg2}
echo BASE; MSSPE_LIB=/root/repo/open-msspe-design_b200/libbase.so python tools/dbg_phases.py $cfg 2>&1 | grep -E "^mode|worker|persistent|incremental" | sed -n "5,7p;10,11p"
echo NEW; python tools/dbg_phases.py $cfg 2>&1 | grep -E "^mode|worker|persistent|incremental" | sed -n "5,7p;10,11p"
