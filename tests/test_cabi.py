"""CPU: the C-ABI library loads, exports every symbol include/sbx.h declares, its struct layout matches
the binding, and it fails loudly (no CPU fallback) when no GPU is present.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sbx.h")


@pytest.fixture(scope="module")
def binding():
    import spectrogram_rs_b200 as s
    return s.binding


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"SBX_API\s+[\w\s\*]+?\b(sbx_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared_symbols()
    for must in ("sbx_plan_create", "sbx_plan_destroy", "sbx_process_magnitudes", "sbx_process_column",
                 "sbx_process_batch", "sbx_process_batch_device", "sbx_colorize_magnitudes", "sbx_get_bin_map",
                 "sbx_live_push", "sbx_live_pop_column", "sbx_num_input_samples", "sbx_num_output_frequencies"):
        assert must in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol(binding):
    lib = binding.load()
    declared = _declared_symbols()
    assert sorted(binding.SYMBOLS) == declared, "binding.SYMBOLS and include/sbx.h disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", binding.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (sbx_\w+)", out))
    assert set(declared) <= exported
    assert exported == set(declared), f"undeclared exports: {exported - set(declared)}"
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sbx_abi_version() == 1 and lib.sbx_build_arch() == b"sm_100a"


def test_library_is_sm100a_and_has_no_cufft_dependency(binding):
    out = subprocess.run(["cuobjdump", "--list-elf", binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "cufft" not in ldd.lower() and "torch" not in ldd.lower()


def test_struct_layout_matches_the_header(binding):
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "sbx.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(sbx_config), offsetof(sbx_config, sample_rate),
         offsetof(sbx_config, channels), offsetof(sbx_config, f_min), offsetof(sbx_config, db_min),
         offsetof(sbx_config, interp), offsetof(sbx_config, device), offsetof(sbx_config, lut_rgb));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "layout.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "layout")
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        vals = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    S = binding.SbxConfig
    assert vals == [C.sizeof(S), S.sample_rate.offset, S.channels.offset, S.f_min.offset, S.db_min.offset,
                    S.interp.offset, S.device.offset, S.lut_rgb.offset]


def test_defaults_are_the_reference_literals(binding):
    lib = binding.load()
    cfg = binding.SbxConfig()
    lib.sbx_config_default(C.byref(cfg))
    assert (cfg.struct_size, cfg.abi_version) == (C.sizeof(binding.SbxConfig), 1)
    assert (cfg.sample_rate, cfg.window_n, cfg.pad_factor, cfg.hop, cfg.rows_h) == (48000.0, 4096, 2, 512, 1024)
    assert (cfg.f_min, cfg.f_max) == (32.0, 22030.0)                         # simple_spectrogram.rs:107
    assert (cfg.db_min, cfg.db_max) == (-70.0, -10.0)                        # colorscheme.rs:16-17
    assert abs(cfg.db_eps - 1e-7) < 1e-12 and cfg.interp == 0 and cfg.scheme == 0
    assert tuple(cfg.lut_rgb[:3]) == (0, 0, 4)                               # magma
    assert lib.sbx_samples_from_seconds(0.05, 48000.0) == 2400               # fft.rs:19
    assert lib.sbx_samples_from_seconds(2.0 / 1024.0, 48000.0) == 93         # audio_transform.rs:35
    assert lib.sbx_samples_from_seconds(-1.0, 48000.0) == 0
    lut = (C.c_uint8 * 768)()
    assert lib.sbx_builtin_lut(b"viridis", lut) == 0 and tuple(lut[:3]) == (68, 1, 84)
    assert lib.sbx_builtin_lut(b"nope", lut) == -1


def test_null_and_invalid_arguments_return_codes(binding):
    lib = binding.load()
    h = C.c_void_p()
    assert lib.sbx_plan_create(None, C.byref(h)) == -1
    cfg = binding.SbxConfig()
    lib.sbx_config_default(C.byref(cfg))
    cfg.struct_size = 12
    assert lib.sbx_plan_create(C.byref(cfg), C.byref(h)) == -1 and not h.value
    assert b"size/version" in lib.sbx_last_error(None)
    lib.sbx_config_default(C.byref(cfg))
    cfg.window_n = 5462                                                       # 2 * 2731: a prime factor > 7, too long for Bluestein
    assert lib.sbx_plan_create(C.byref(cfg), C.byref(h)) == -2 and b"prime factor" in lib.sbx_last_error(None)
    lib.sbx_config_default(C.byref(cfg))
    cfg.window_n = 131072
    assert lib.sbx_plan_create(C.byref(cfg), C.byref(h)) == -2
    lib.sbx_config_default(C.byref(cfg))
    cfg.channels = 3                                                          # audio_input_list_model.rs:73-75
    assert lib.sbx_plan_create(C.byref(cfg), C.byref(h)) == -1
    lib.sbx_plan_destroy(None)
    lib.sbx_live_destroy(None)
    assert lib.sbx_num_input_samples(None) == 0 and lib.sbx_num_frames(None, 10) == 0
    assert lib.sbx_process_magnitudes(None, None, 0, None) == -1
    assert lib.sbx_live_push(None, None, 0, 2) == 0


def test_no_cpu_fallback_without_a_gpu():
    """In the CPU-only container plan creation must fail with SBX_ERR_CUDA, not compute anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import spectrogram_rs_b200 as s
    with pytest.raises(s.SbxError) as e:
        s.Plan()
    assert e.value.status == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(s.SbxError) as e:                                      # the several-device entry point likewise
        s.MultiPlan(2)
    assert e.value.status == -3 and "no CPU fallback" in str(e.value) and "device 0" in str(e.value)
    lib = s.binding.load()
    assert lib.sbx_multi_device_count(None) == 0 and lib.sbx_multi_plan(None, 0) is None
    assert lib.sbx_multi_process_batch(None, None, 0, None, 0, None) == -1
    assert lib.sbx_multi_last_error(None) is not None
    # the built-in gradients are host functions: available without a GPU
    assert s.builtin_lut("turbo").shape == (256, 3) and s.builtin_gradient("red_blue", 1024).shape == (1024, 3)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under spectrogram-rs_b200/ may import, include or link it."""
    pkg = os.path.join(ROOT, "spectrogram-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("liboracle", "libemu", "stft_oracle", "from oracle", "import oracle", "oracle/", "orc_"):
                    assert needle not in text, f"{f} references the oracle ({needle})"


def _build_cpp_example(tmpdir):
    host = os.path.join(ROOT, "spectrogram-rs_b200", "host")
    exe = os.path.join(str(tmpdir), "example_stream")
    lib_dir = os.path.join(ROOT, "spectrogram-rs_b200")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(host, "example_stream.cpp"),
                    "-L", lib_dir, "-lsbx_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    return exe


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(binding, tmp_path):
    """spectrogram-rs_b200/host/audio_transform.hpp (AudioTransform / AudioStreamTransform<T> / HeapRb mirror)."""
    binding.load()
    exe = _build_cpp_example(tmp_path)
    import torch
    out = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stderr
    else:
        assert out.returncode == 2 and "no CPU fallback" in out.stderr


# ---- INTEGRATION.md's Rust binding vs include/sbx.h --------------------------------------------------------------
_C_TO_RUST = {"float": "f32", "double": "f64", "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64",
              "int32_t": "i32", "sbx_status": "i32", "size_t": "usize", "void": "c_void", "char": "c_char",
              "sbx_plan": "SbxPlan", "sbx_config": "SbxConfig", "sbx_multi": "SbxMulti", "sbx_live": "SbxLive"}


def _rust_type_of_c(decl):
    """'const float *pcm' -> '*const f32'; 'uint8_t lut_rgb[768]' -> '*mut u8'; 'uint32_t n' -> 'u32'."""
    decl = decl.strip()
    array = decl.endswith("]")
    decl = re.sub(r"\[[^\]]*\]$", "", decl).strip()
    stars = decl.count("*") + (1 if array else 0)
    words = decl.replace("*", " ").split()
    const = "const" in words
    words = [w for w in words if w != "const"]
    base = words[0]                                     # words[1:] is the parameter name (absent for `void`)
    rust = _C_TO_RUST[base]
    for level in range(stars):
        rust = ("*const " if const and level == 0 else "*mut ") + rust
    return rust


def _header_signatures():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    sigs = {}
    for ret, name, args in re.findall(r"SBX_API\s+([\w\s\*]+?)\b(sbx_\w+)\s*\(([^)]*)\)\s*;", text):
        params = [] if args.strip() in ("", "void") else [_rust_type_of_c(a) for a in args.split(",")]
        ret = ret.strip()
        sigs[name] = (params, None if ret == "void" else _rust_type_of_c(ret + " _"))
    return sigs


def test_integration_md_rust_block_matches_the_header():
    """The Rust shim cannot be compiled in this image (no rustc): keep its extern "C" block and its #[repr(C)]
    config struct consistent with include/sbx.h by parsing both."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', md, flags=re.S).group(1)
    block = re.sub(r"//[^\n]*", "", block)
    rust = {}
    for name, args, ret in re.findall(r"fn\s+(sbx_\w+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        params = [a.split(":", 1)[1].strip() for a in args.split(",") if a.strip()]
        rust[name] = (params, ret.strip() if ret else None)
    assert len(rust) >= 15
    header = _header_signatures()
    for name, sig in rust.items():
        assert name in header, f"{name} is not declared in include/sbx.h"
        assert sig == header[name], f"{name}: INTEGRATION.md {sig} vs sbx.h {header[name]}"
    # every sbx_* call the shim makes is declared in its extern block
    code = "\n".join(re.findall(r"```rust(.*?)```", md, flags=re.S))
    assert set(re.findall(r"\b(sbx_\w+)\s*\(", code)) <= set(rust)
    # struct SbxConfig: same fields, order and types as sbx_config
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct sbx_config \{(.*?)\} sbx_config;", text, flags=re.S).group(1)
    c_fields = []
    for ctype, names in re.findall(r"(\w+)\s+([^;]+);", body):
        for nm in names.split(","):
            nm = nm.strip()
            m = re.match(r"(\w+)\[(\d+)\]", nm)
            c_fields.append((m.group(1), f"[{_C_TO_RUST[ctype]}; {m.group(2)}]") if m else (nm, _C_TO_RUST[ctype]))
    rs_body = re.search(r"pub struct SbxConfig \{(.*?)\n\}", md, flags=re.S).group(1)
    rs_fields = [(n, t.strip()) for n, t in re.findall(r"(\w+)\s*:\s*([^,\n]+),", rs_body)]
    assert rs_fields == c_fields
