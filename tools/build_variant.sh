#!/bin/bash
# Builds a variant of the library for A/B measurements: tools/build_variant.sh NAME [-DMACRO=VALUE ...]
#   -> variants/libsbx_NAME.so; select it at run time with SBX_LIB_PATH=variants/libsbx_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
python - "$name" "$@" <<'PY'
import importlib.util, os, sys
spec = importlib.util.spec_from_file_location("sbx_build", "spectrogram-rs_b200/build.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
name, flags = sys.argv[1], sys.argv[2:]
print(mod.build(force=True, extra_flags=flags, lib=os.path.abspath(f"variants/libsbx_{name}.so"),
                obj_dir=os.path.abspath(f"variants/_build_{name}")))
PY
