# A/B of library variants (tools/build_variant.sh) on the GPU box: label, then env assignments
run() {
  label=$1; shift
  for c in $CFGS; do env "$@" python tools/profile_kernel.py --cfg $c --frames 262144 --iters 10 2>&1 | tail -1 | sed "s/^/$label /"; done
}
CFGS=${CFGS:-"1 2 8"}
run main
run sp1 SBX_LIB_PATH=variants/libsbx_sp1.so
run sp2 SBX_LIB_PATH=variants/libsbx_sp2.so
