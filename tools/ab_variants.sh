run() {
  label=$1; shift
  for c in $CFGS; do env "$@" python tools/profile_kernel.py --cfg $c --frames 262144 --iters 10 2>&1 | tail -1 | sed "s/^/$label /"; done
}
CFGS=${CFGS:-"1 3 8"}
run wtw-compute
run wtw-table SBX_LIB_PATH=variants/libsbx_wtwtab.so
run wtw-compute
run wtw-table SBX_LIB_PATH=variants/libsbx_wtwtab.so
