"""ctypes binding of libsbx_b200.so (include/sbx.h).  Loading fails loudly if the library has not
been built; there is no Python or CPU implementation of the transform behind it."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SBX_LIB_PATH") or os.path.join(HERE, "libsbx_b200.so")   # override: kernel-variant experiments

SBX_OK, SBX_NOT_ENOUGH_SAMPLES = 0, 1
SBX_ERR_INVALID_ARGUMENT, SBX_ERR_UNSUPPORTED, SBX_ERR_CUDA, SBX_ERR_OUT_OF_MEMORY, SBX_ERR_BUFFER_TOO_SMALL = -1, -2, -3, -4, -5
INTERP_CUBIC, INTERP_COSINE = 0, 1
SCHEME_MONO, SCHEME_STEREO = 0, 1
ABI_VERSION = 1


class SbxConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("abi_version", C.c_uint32), ("sample_rate", C.c_float),
                ("window_n", C.c_uint32), ("pad_factor", C.c_uint32), ("hop", C.c_uint32),
                ("rows_h", C.c_uint32), ("channels", C.c_uint32), ("f_min", C.c_double), ("f_max", C.c_double),
                ("db_min", C.c_float), ("db_max", C.c_float), ("db_eps", C.c_float), ("interp", C.c_uint32),
                ("scheme", C.c_uint32), ("device", C.c_int32), ("reserved", C.c_uint32),
                ("lut_rgb", C.c_uint8 * 768)]


# every symbol include/sbx.h declares: name -> (restype, argtypes)
_f32p, _u8p, _u32p, _u64p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
_vp = C.c_void_p
SYMBOLS = {
    "sbx_config_default": (None, [C.POINTER(SbxConfig)]),
    "sbx_builtin_lut": (C.c_int32, [C.c_char_p, _u8p]),
    "sbx_builtin_gradient": (C.c_int32, [C.c_char_p, C.c_uint32, _vp]),
    "sbx_builtin_gradient_count": (C.c_uint32, []),
    "sbx_builtin_gradient_name": (C.c_char_p, [C.c_uint32, C.POINTER(C.c_int32)]),
    "sbx_plan_set_pan_gradient": (C.c_int32, [_vp, _vp, C.c_uint32]),
    "sbx_samples_from_seconds": (C.c_uint32, [C.c_float, C.c_float]),
    "sbx_plan_create": (C.c_int32, [C.POINTER(SbxConfig), C.POINTER(_vp)]),
    "sbx_plan_destroy": (None, [_vp]),
    "sbx_last_error": (C.c_char_p, [_vp]),
    "sbx_sample_rate": (C.c_float, [_vp]),
    "sbx_num_input_samples": (C.c_uint32, [_vp]),
    "sbx_num_output_frequencies": (C.c_uint32, [_vp]),
    "sbx_fft_length": (C.c_uint32, [_vp]),
    "sbx_rows": (C.c_uint32, [_vp]),
    "sbx_hop": (C.c_uint32, [_vp]),
    "sbx_channels": (C.c_uint32, [_vp]),
    "sbx_num_frames": (C.c_uint64, [_vp, C.c_uint64]),
    "sbx_process_magnitudes": (C.c_int32, [_vp, _vp, C.c_uint32, _vp]),
    "sbx_process_column": (C.c_int32, [_vp, _vp, C.c_uint32, _vp]),
    "sbx_colorize_magnitudes": (C.c_int32, [_vp, _vp, C.c_uint64, _vp]),
    "sbx_process_batch": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "sbx_process_batch_device": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "sbx_process_batch_device_planar": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "sbx_magnitudes_batch_device": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "sbx_magnitudes_batch_device_f16": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "sbx_palette_table": (C.c_int32, [_vp, C.c_uint32, _vp]),
    "sbx_blit_columns_device": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _u32p]),
    "sbx_set_stream": (C.c_int32, [_vp, _vp]),
    "sbx_synchronize": (C.c_int32, [_vp]),
    "sbx_launch_count": (C.c_uint64, [_vp]),
    "sbx_kernel_path": (C.c_char_p, [_vp]),
    "sbx_get_bin_map": (C.c_int32, [_vp, _f32p, _f32p, _u32p, _f32p]),
    "sbx_get_color_thresholds": (C.c_int32, [_vp, _f32p]),
    "sbx_get_window": (C.c_int32, [_vp, _f32p]),
    "sbx_live_create": (C.c_int32, [_vp, C.c_uint32, C.POINTER(_vp)]),
    "sbx_live_destroy": (None, [_vp]),
    "sbx_live_push": (C.c_uint32, [_vp, _vp, C.c_uint32, C.c_uint32]),
    "sbx_live_pop_column": (C.c_int32, [_vp, _vp]),
    "sbx_live_pop_magnitudes": (C.c_int32, [_vp, _vp]),
    "sbx_live_buffered": (C.c_uint32, [_vp]),
    "sbx_live_measure": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "sbx_multi_create": (C.c_int32, [C.POINTER(SbxConfig), C.POINTER(C.c_int32), C.c_uint32, C.POINTER(_vp)]),
    "sbx_multi_destroy": (None, [_vp]),
    "sbx_multi_last_error": (C.c_char_p, [_vp]),
    "sbx_multi_device_count": (C.c_uint32, [_vp]),
    "sbx_multi_plan": (_vp, [_vp, C.c_uint32]),
    "sbx_multi_time_block": (C.c_int32, [_vp, C.c_uint64, C.c_uint32, _u64p, _u64p, _u64p, _u64p]),
    "sbx_shard_time_block": (C.c_int32, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _u64p, _u64p, _u64p]),
    "sbx_multi_process_batch": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "sbx_multi_process_batch_planar": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "sbx_multi_launch_count": (C.c_uint64, [_vp]),
    "sbx_host_register": (C.c_int32, [_vp, C.c_size_t]),
    "sbx_host_unregister": (C.c_int32, [_vp]),
    "sbx_host_alloc": (_vp, [C.c_size_t]),
    "sbx_host_free": (None, [_vp]),
    "sbx_abi_version": (C.c_uint32, []),
    "sbx_build_arch": (C.c_char_p, []),
}

_lib = None


def load():
    """Loads libsbx_b200.so.  Raises if it is missing: build it with `python spectrogram-rs_b200/build.py`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is not built (run `python spectrogram-rs_b200/build.py`); "
                "the transform has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if L.sbx_abi_version() != ABI_VERSION:
            raise RuntimeError("libsbx_b200.so ABI version mismatch")
        _lib = L
    return _lib
