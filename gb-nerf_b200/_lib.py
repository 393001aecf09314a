"""ctypes binding of libgbnerf.so — the reference-side stub of INTEGRATION.md, verbatim.

The library is the product: if it is missing or a call fails this module raises; nothing here (or anywhere in
the package) falls back to PyTorch or to the CPU oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GBNERF_LIB: another build of the same library (tools: libgbnerf_diag.so, csrc/build.py --diag); never a fallback
LIB_PATH = os.environ.get("GBNERF_LIB") or os.path.join(_HERE, "libgbnerf.so")

PRECISION = {"bf16": 0, "tf32": 1}
PACK_BWD_BF16 = 2

_p, _i64, _i, _f, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/gbnerf.h one to one
SIGNATURES = {
    "gbn_version": (_i, []),
    "gbn_last_error_string": (C.c_char_p, []),
    "gbn_kernel_launches": (C.c_ulonglong, []),
    "gbn_pack_rays": (_i, [_p, _i, _p, _i, _p, _i64, _p, _i64, _p, _i, _i, C.c_double, _i, _i, _i, _i, _i, _i, _f, _f, _i64,
                           _p, _p]),
    "gbn_zvals_stratified": (_i, [_p, _p, _i64, _i64, _i, _i, _p, _p, _p]),
    "gbn_encode_points": (_i, [_p, _p, _p, _i64, _p, _i64, _i, _p, _p]),
    "gbn_composite_forward": (_i, [_p, _p, _p, _i64, _p, _i64, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "gbn_composite_backward": (_i, [_p, _p, _p, _i64, _p, _i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "gbn_sample_pdf": (_i, [_p, _p, _p, _i64, _i, _i, _p, _p]),
    "gbn_searchsorted_right": (_i, [_p, _p, _i64, _i, _i, _p, _p]),
    "gbn_sample_pdf_merge": (_i, [_p, _p, _p, _i64, _i, _i, _p, _p, _p, _p]),
    "gbn_sample_pdf_merge_ex": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _p]),
    "gbn_mlp_packed_bytes": (_sz, [_i]),
    "gbn_mlp_prepack_weights": (_i, [C.POINTER(_p), _p, _i, _p]),
    "gbn_mlp_workspace_bytes": (_sz, [_i64]),
    "gbn_mlp_forward": (_i, [_p, _i, _p, _p, _p, _i64, _p, _p, _i64, _i, _p, _p, _p, _p]),
    "gbn_mlp_stash_bytes": (_sz, [_i64]),
    "gbn_mlp_backward_data": (_i, [_p, _p, _i64, _p, _p, _p, _p]),
    "gbn_mlp_wgrad_workspace_bytes": (_sz, [_i64]),
    "gbn_mlp_backward_weights": (_i, [_p, _p, _p, _p, _i64, _i64, _i, C.POINTER(_p), _p, _p]),
    "gbn_mlp_set_trace": (_i, [_p, _i]),
    "gbn_watchdog_report": (_i, [_p, _i]),
    "gbn_debug_ts_plan": (_i, [_i, _p, _i, _p, _i, _p]),
    "gbn_mlp_forward_embedded": (_i, [_p, _i, _p, _i64, _p, _p, _p, _p]),
    "gbn_mlp_variant": (_i, []),
    "gbn_normals_forward": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "gbn_normals_backward": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "gbn_tcnn_table_bytes": (_sz, []),
    "gbn_tcnn_grid_params": (_sz, []),
    "gbn_tcnn_prepack": (_i, [_p, _p, _p, _p, _p]),
    "gbn_tcnn_forward": (_i, [_p, _p, _p, _p, _i64, _p, _p, _i64, _i, _p, _p, _p]),
    "gbn_tcnn_backward": (_i, [_p, _p, _p, _p, _i64, _p, _p, _i64, _i, _p, _p, _f, _p, _p, _p, _p, _p]),
    "gbn_adam_step_repack": (_i, [_p, _p, _p, _p, C.c_double, C.c_double, C.c_double, C.c_double, _i64, _p, _p, _p]),
    "gbn_adam_tick": (_i, [_p, _p, C.c_double, C.c_double, _p, _p]),
    "gbn_adam_step_repack_dev": (_i, [_p, _p, _p, _p, _p, C.c_double, C.c_double, C.c_double, _p, _p, _p]),
    "gbn_loss_seed": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _f, _p, _p, _p, _p, _p]),
}


def kernel_launches():
    """Kernels the library has launched in this process so far (counted at every launch site in csrc/; bench.py
    reports the difference over its timed region as ``gpu_launches``)."""
    return int(load().gbn_kernel_launches())


def watchdog_report():
    """First barrier-watchdog expiry of the MLP kernels in this process (host memory: readable after the CUDA context
    died), or None.  See gbn_watchdog_report in include/gbnerf.h."""
    buf = (C.c_uint32 * 256)()
    n = load().gbn_watchdog_report(buf, 256)
    if n < 4 or buf[3] == 0:
        return None
    return {"code": hex(buf[0]), "cta": buf[1], "thread": buf[2],
            "barriers_from_end": [hex(buf[8 + 2 * i] | (buf[9 + 2 * i] << 32)) for i in range(32)],
            "warp_waits": {w: (hex(buf[128 + w]), buf[160 + w]) for w in range(16) if buf[128 + w]}}


_lib = None


class GbnError(RuntimeError):
    pass


def load():
    """Load (once) and type the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GbnError(f"{LIB_PATH} is missing: build it with `python gb-nerf_b200/csrc/build.py` "
                       "(or __graft_entry__.build()); there is no fallback path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header ever drift apart
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an int-returning entry point and raise on a non-zero code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.gbn_last_error_string().decode(errors="replace")
        if rc == 1:
            raise ValueError(f"{name}: {msg}")
        raise GbnError(f"{name} failed (code {rc}): {msg}")
