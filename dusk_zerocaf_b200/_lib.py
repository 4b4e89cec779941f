"""ctypes loader for libzerocaf_b200.so (the C ABI declared in include/zerocaf_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a context is created,
this raises.  Nothing here imports anything under oracle/.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZC_LIB_PATH: an alternative build of the same library (A/B runs of kernel variants); default: the in-tree build
SO_PATH = os.environ.get("ZC_LIB_PATH") or os.path.join(_HERE, "libzerocaf_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "zerocaf_b200.h")

_u64p = ctypes.c_void_p   # pointers are passed as integers (host numpy .ctypes.data or device data_ptr())
_lib = None


class ZerocafError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"zerocaf_b200 status {status}: {msg}")
        self.status = status


def header_symbols():
    """Every function name declared in include/zerocaf_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zc_[a-z0-9_]+)\s*\(", text)))


def build(force=False):
    """Compile the library in-tree with csrc/Makefile (nvcc, sm_100a only)."""
    import subprocess
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "-s", "clean"])
    subprocess.check_call(["make", "-C", csrc, "-s", "-j4"])
    return SO_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `make -C dusk_zerocaf_b200/csrc` (or __graft_entry__.build()). "
            "zerocaf_b200 has no CPU fallback.")
    L = ctypes.CDLL(SO_PATH, mode=ctypes.RTLD_GLOBAL)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32
    L.zc_version.restype = ctypes.c_char_p
    L.zc_last_error_string.restype = ctypes.c_char_p
    L.zc_last_error_string.argtypes = [vp]
    L.zc_ctx_launch_count.restype = ctypes.c_uint64
    L.zc_ctx_launch_count.argtypes = [vp]
    L.zc_ctx_create.argtypes = [i32, vp, ctypes.POINTER(vp)]
    L.zc_ctx_destroy.argtypes = [vp]
    L.zc_ctx_sync.argtypes = [vp]
    L.zc_host_alloc.argtypes = [sz, ctypes.POINTER(vp)]
    L.zc_host_free.argtypes = [vp]
    L.zc_host_register.argtypes = [vp, sz]
    L.zc_host_unregister.argtypes = [vp]
    for mod in ("fe", "scalar"):
        for op in ("mul", "add", "sub"):
            for suf in ("", "_dev"):
                getattr(L, f"zc_{mod}_{op}_batch{suf}").argtypes = [vp, vp, vp, vp, sz]
        for op in ("square", "neg"):
            for suf in ("", "_dev"):
                getattr(L, f"zc_{mod}_{op}_batch{suf}").argtypes = [vp, vp, vp, sz]
    for suf in ("", "_dev"):
        getattr(L, f"zc_fe_mul_square_batch{suf}").argtypes = [vp, vp, vp, vp, vp, sz]
        for op in ("add", "sub"):
            getattr(L, f"zc_point_{op}_batch{suf}").argtypes = [vp, vp, vp, vp, sz]
        for op in ("double", "neg"):
            getattr(L, f"zc_point_{op}_batch{suf}").argtypes = [vp, vp, vp, sz]
        getattr(L, f"zc_point_scalar_mul_batch{suf}").argtypes = [vp, vp, vp, vp, sz, i32]
        getattr(L, f"zc_ristretto_eq_batch{suf}").argtypes = [vp, vp, vp, vp, sz]
        getattr(L, f"zc_msm{suf}").argtypes = [vp, vp, vp, sz, i32, vp]
        for nm in ("zc_fe_invert_batch", "zc_point_to_affine_batch", "zc_ristretto_compress_batch"):
            getattr(L, nm + suf).argtypes = [vp, vp, vp, sz]
        getattr(L, "zc_point_is_valid_batch" + suf).argtypes = [vp, vp, vp, sz]
        getattr(L, "zc_basepoint_mul_batch" + suf).argtypes = [vp, vp, vp, sz]
        getattr(L, "zc_ristretto_elligator_batch" + suf).argtypes = [vp, vp, vp, sz]
        getattr(L, "zc_ristretto_from_uniform_bytes_batch" + suf).argtypes = [vp, vp, vp, sz]
        getattr(L, "zc_ristretto_decompress_batch" + suf).argtypes = [vp, vp, vp, vp, sz]
        for mod in ("fe", "scalar"):
            getattr(L, f"zc_{mod}_pow_batch{suf}").argtypes = [vp, vp, vp, vp, sz]
            getattr(L, f"zc_{mod}_half_batch{suf}").argtypes = [vp, vp, vp, sz]
            getattr(L, f"zc_{mod}_to_bytes_batch{suf}").argtypes = [vp, vp, vp, sz]
        getattr(L, "zc_fe_from_bytes_batch" + suf).argtypes = [vp, vp, vp, sz]
        getattr(L, "zc_scalar_from_bytes_batch" + suf).argtypes = [vp, vp, vp, vp, sz]
        getattr(L, "zc_scalar_window_naf_batch" + suf).argtypes = [vp, vp, i32, vp, sz]
        getattr(L, "zc_fe_sqrt_ratio_i_batch" + suf).argtypes = [vp, vp, vp, vp, vp, sz]
    L.zc_msm_sharded_dev.argtypes = [vp, vp, vp, sz, i32, vp]
    L.zc_msm_sharded.argtypes = [vp, vp, vp, sz, i32, vp]
    u64p = ctypes.POINTER(ctypes.c_uint64)
    for suf in ("", "_dev"):
        getattr(L, "zc_fe_mul_square_batch_packed" + suf).argtypes = [vp, vp, vp, vp, vp, sz]
        getattr(L, "zc_fe_div_batch" + suf).argtypes = [vp, vp, vp, vp, sz]
        getattr(L, "zc_scalar_into_bits_batch" + suf).argtypes = [vp, vp, vp, sz]
        for kind in ("fe", "scalar", "point"):
            getattr(L, f"zc_{kind}_check_canonical_batch{suf}").argtypes = [vp, vp, sz, u64p]
    L.zc_ctx_set_validation.argtypes = [vp, i32]
    L.zc_msm_plan_query.argtypes = [i32, i32, i32, ctypes.POINTER(i32)]
    L.zc_msm_generators_create_dev.argtypes = [vp, vp, sz, i32, i32, i32, i32, ctypes.POINTER(vp)]
    L.zc_msm_generators_destroy.argtypes = [vp, vp]
    L.zc_msm_generators_info.argtypes = [vp, ctypes.POINTER(sz), ctypes.POINTER(i32), ctypes.POINTER(sz)]
    L.zc_msm_gen_dev.argtypes = [vp, vp, vp, i32, vp]
    L.zc_msm_gen_partial_dev.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.zc_msm_gen_sharded_dev.argtypes = [vp, vp, vp, i32, vp]
    L.zc_msm_partial_dev.argtypes = [vp, vp, vp, sz, i32, i32, i32, vp]
    L.zc_point_fold_dev.argtypes = [vp, vp, sz, vp]
    L.zc_ctx_set_nccl.argtypes = [vp, vp, i32, i32]
    L.zc_nccl_unique_id.argtypes = [vp]
    L.zc_peer_mailbox_create.argtypes = [vp, vp]
    L.zc_peer_mailbox_connect.argtypes = [vp, vp, i32, i32]
    L.zc_peer_mailbox_ptr.argtypes = [vp, ctypes.POINTER(vp)]
    L.zc_peer_mailbox_connect_local.argtypes = [vp, ctypes.POINTER(vp), i32, i32]
    L.zc_nccl_comm_init.argtypes = [vp, i32, i32, ctypes.POINTER(vp)]
    L.zc_nccl_comm_destroy.argtypes = [vp]
    _lib = L
    return L
