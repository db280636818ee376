# A/B of library variants (tools/build_variant.sh) on the GPU box: label, then env assignments
run() {
  label=$1; shift
  for c in $CFGS; do env "$@" python tools/profile_kernel.py --cfg $c --frames 262144 --iters 10 2>&1 | tail -1 | sed "s/^/$label /"; done
}
CFGS=${CFGS:-"1 2"}
run q2-u8r4 SBX_Q2=1
run q2-u4r2 SBX_Q2=1 SBX_LIB_PATH=variants/libsbx_u4r2.so
run q2-u4r4 SBX_Q2=1 SBX_LIB_PATH=variants/libsbx_u4r4.so
run q2-u8r2 SBX_Q2=1 SBX_LIB_PATH=variants/libsbx_u8r2.so
CFGS="7 8"
run base SBX_Q2=0
run q2 SBX_Q2=1
