"""ctypes binding of libcvb200.so (include/cvb200.h) + the in-tree build recipe.

There is no CPU fallback: if the shared library is missing or cannot be loaded the
import of any model module raises, and every call checks the C status code.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcvb200.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp"))) + \
        [os.path.join(INCLUDE, "cvb200.h")]


def source_hash():
    """12 hex digits over the library's sources (csrc/ + include/): names the build that a profile / traffic file belongs to"""
    import hashlib
    h = hashlib.sha256()
    for fn in sources():
        h.update(os.path.basename(fn).encode())
        with open(fn, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:12]


def build(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU (a few seconds)."""
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in sources())
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH, os.path.join(CSRC, "cvb200.cu")] + \
        sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cpp"))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol include/cvb200.h declares
SIGNATURES = {
    "cvb_last_error": (ctypes.c_char_p, []),
    "cvb_version": (ctypes.c_int, []),
    "cvb_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "cvb_destroy": (ctypes.c_int, [c_vp]),
    "cvb_num_variables": (ctypes.c_int, [c_vp]),
    "cvb_variable_info": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                         ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(c_i64)]),
    "cvb_num_parameters": (c_i64, [c_vp]),
    "cvb_set_variable": (ctypes.c_int, [c_vp, ctypes.c_char_p, ctypes.c_int, c_vp, c_i64]),
    "cvb_get_variable": (ctypes.c_int, [c_vp, ctypes.c_char_p, ctypes.c_int, c_vp, c_i64]),
    "cvb_init_weights": (ctypes.c_int, [c_vp, ctypes.c_uint64]),
    "cvb_set_step": (ctypes.c_int, [c_vp, c_i64]),
    "cvb_get_step": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "cvb_set_compute_mode": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "cvb_predict_host": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cvb_predict_host_f16": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cvb_predict_device": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cvb_predict_device_x": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_vp, c_vp, c_vp]),
    "cvb_predict_submit": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "cvb_predict_collect": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cvb_predict_host_counts_i16": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cvb_predict_host_counts_u8": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cvb_loss_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "cvb_train_step_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.c_float, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_uint64, ctypes.c_int, c_vp]),
    "cvb_set_dropout_fc5": (ctypes.c_int, [c_vp, ctypes.c_float]),
    "cvb_loss_host_x": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, c_i64, c_vp]),
    "cvb_train_step_host_x": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, c_i64, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_float, ctypes.c_uint64, ctypes.c_int, c_vp]),
    "cvb_set_train_mode": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "cvb_grad_buffer": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64)]),
    "cvb_nccl_unique_id": (ctypes.c_int, [c_vp]),
    "cvb_allreduce_init": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int]),
    "cvb_allreduce_attach": (ctypes.c_int, [c_vp, c_vp]),
    "cvb_allreduce_gradients": (ctypes.c_int, [c_vp]),
    "cvb_get_gradient": (ctypes.c_int, [c_vp, ctypes.c_char_p, c_vp, c_i64]),
    "cvb_apply_adam": (ctypes.c_int, [c_vp, ctypes.c_float, ctypes.c_float, c_vp]),
    "cvb_parse_tensor_text": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, c_i64, ctypes.c_int, c_vp, c_vp,
                                             ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "cvb_pack_counts": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp, c_vp,
                                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "cvb_blosc_info": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "cvb_blosc_decompress": (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "cvb_pileup_create": (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.POINTER(c_vp)]),
    "cvb_pileup_destroy": (ctypes.c_int, [c_vp]),
    "cvb_pileup_set_threads": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "cvb_pileup_feed": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int]),
    "cvb_pileup_ready": (c_i64, [c_vp]),
    "cvb_pileup_take": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, ctypes.POINTER(c_i64)]),
    "cvb_pileup_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "cvb_pileup_format_rows": (c_i64, [ctypes.c_char_p, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64]),
    "cvb_candidates_create": (ctypes.c_int, [ctypes.c_char_p, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, ctypes.c_int,
                                             ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64,
                                             ctypes.POINTER(c_vp)]),
    "cvb_candidates_destroy": (ctypes.c_int, [c_vp]),
    "cvb_candidates_set_threads": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "cvb_candidates_feed": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int]),
    "cvb_candidates_pending_bytes": (c_i64, [c_vp]),
    "cvb_candidates_pending": (c_i64, [c_vp]),
    "cvb_candidates_take": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "cvb_candidates_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "cvb_tensor_text_positions": (c_i64, [c_vp, c_vp, c_i64, c_vp, c_i64]),
    "cvb_vcf_records": (c_i64, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp, c_i64]),
    "cvb_sam_view": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, c_i64, c_i64, c_vp,
                                    ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "cvb_blosc_compress_bound": (c_i64, [c_i64]),
    "cvb_blosc_compress": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "cvb_crc32c": (ctypes.c_uint32, [ctypes.c_uint32, c_vp, c_i64]),
    "cvb_alloc_pinned": (ctypes.c_int, [c_i64, ctypes.POINTER(c_vp)]),
    "cvb_free_pinned": (ctypes.c_int, [c_vp]),
    "cvb_debug_read": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_i64]),
    "cvb_profile_begin": (ctypes.c_int, [c_vp]),
    "cvb_profile_read": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i64)]),
    "cvb_kernel_launches": (c_i64, [c_vp]),
}


def load():
    """Load (never build) the library and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libcvb200.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` -- there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)      # plain CDLL releases the GIL around every call (SURVEY 8b threading)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise RuntimeError("libcvb200: " + load().cvb_last_error().decode("utf-8", "replace"))
