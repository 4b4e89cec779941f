"""Context = one device + one stream + scratch arenas (zc_ctx in include/zerocaf_b200.h)."""
import ctypes

import numpy as np

from ._lib import ZerocafError, lib


def _ptr(x):
    """Host numpy array or torch tensor (host or device) -> raw address."""
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if x is None:
        return None
    return int(x)


class Context:
    def __init__(self, device=0, stream=None):
        self._L = lib()
        h = ctypes.c_void_p()
        st = self._L.zc_ctx_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(h))
        if st != 0:
            raise ZerocafError(st, "zc_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h
        self.device = int(device)
        self._comm = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.zc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st):
        if st != 0:
            raise ZerocafError(st, self._L.zc_last_error_string(self._h).decode())

    def call(self, name, *args):
        self.check(getattr(self._L, name)(self._h, *[_ptr(a) if not isinstance(a, int) else a for a in args]))

    def sync(self):
        self.check(self._L.zc_ctx_sync(self._h))

    def set_validation(self, on=True):
        """zc_ctx_set_validation: the hot-path entry points check their inputs (limbs < 2^52, value < modulus) and return
        ZC_ERR_NONCANONICAL (status 4) instead of a silently wrong result."""
        self.check(self._L.zc_ctx_set_validation(self._h, 1 if on else 0))

    @property
    def launches(self):
        return int(self._L.zc_ctx_launch_count(self._h))

    # ---- NCCL plumbing for the sharded MSM: unique id travels over torch.distributed ------------------------
    def init_nccl(self, rank, nranks, broadcast_bytes):
        """broadcast_bytes(buf: bytes|None) -> bytes: rank 0 passes the id, everyone gets it back."""
        buf = (ctypes.c_uint8 * 128)()
        if rank == 0:
            st = self._L.zc_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p))
            if st != 0:
                raise ZerocafError(st, "zc_nccl_unique_id failed (libnccl not loadable)")
            idb = broadcast_bytes(bytes(buf))
        else:
            idb = broadcast_bytes(None)
        buf2 = (ctypes.c_uint8 * 128).from_buffer_copy(idb)
        comm = ctypes.c_void_p()
        st = self._L.zc_nccl_comm_init(ctypes.cast(buf2, ctypes.c_void_p), int(rank), int(nranks), ctypes.byref(comm))
        if st != 0:
            raise ZerocafError(st, "zc_nccl_comm_init failed")
        self._comm = comm
        self.check(self._L.zc_ctx_set_nccl(self._h, comm, int(rank), int(nranks)))


    def init_peer_mailboxes(self, rank, nranks, allgather_bytes):
        """NVLink peer-memory exchange for the sharded MSM.  allgather_bytes(b: bytes) -> list[bytes] in rank order
        (the host framework's transport, e.g. torch.distributed.all_gather_object)."""
        buf = (ctypes.c_uint8 * 64)()
        self.check(self._L.zc_peer_mailbox_create(self._h, ctypes.cast(buf, ctypes.c_void_p)))
        handles = allgather_bytes(bytes(buf))
        assert len(handles) == nranks and all(len(h) == 64 for h in handles)
        allb = (ctypes.c_uint8 * (64 * nranks)).from_buffer_copy(b"".join(handles))
        self.check(self._L.zc_peer_mailbox_connect(self._h, ctypes.cast(allb, ctypes.c_void_p), int(rank), int(nranks)))


    @staticmethod
    def connect_local(contexts):
        """Several contexts of THIS process (several GPUs with peer access, or several streams of one GPU) as the ranks of
        one sharded MSM: zc_peer_mailbox_ptr + zc_peer_mailbox_connect_local, rank = position in the list."""
        n = len(contexts)
        ptrs = (ctypes.c_void_p * n)()
        for r, cx in enumerate(contexts):
            scratch = (ctypes.c_uint8 * 64)()
            cx.check(cx._L.zc_peer_mailbox_create(cx._h, ctypes.cast(scratch, ctypes.c_void_p)))
            p = ctypes.c_void_p()
            cx.check(cx._L.zc_peer_mailbox_ptr(cx._h, ctypes.byref(p)))
            ptrs[r] = p.value
        for r, cx in enumerate(contexts):
            cx.check(cx._L.zc_peer_mailbox_connect_local(cx._h, ptrs, r, n))

    # ---- fixed generators --------------------------------------------------------------------------------------
    def msm_generators(self, points_dev_ptr, n, kind=1, window_bits=16, rank=0, nranks=1):
        """zc_msm_generators_create_dev: an opaque handle owning the prepared operands (kind 1, ZC_GEN_PREPARED) or the
        pre-scaled fixed-base tables (kind 2, ZC_GEN_FIXED_BASE) of a resident point set.  The points are read during
        creation only."""
        return Generators(self, points_dev_ptr, n, kind, window_bits, rank, nranks)


GEN_PREPARED, GEN_FIXED_BASE = 1, 2


class Generators:
    """Handle of zc_msm_generators (include/zerocaf_b200.h): close() (or the owning context's close) releases it."""

    def __init__(self, ctx, points_dev_ptr, n, kind, window_bits, rank, nranks):
        self.ctx = ctx
        h = ctypes.c_void_p()
        ctx.check(ctx._L.zc_msm_generators_create_dev(ctx._h, points_dev_ptr, int(n), int(kind), int(window_bits), int(rank),
                                                      int(nranks), ctypes.byref(h)))
        self._g = h
        self.n, self.kind, self.window_bits = int(n), int(kind), int(window_bits)

    @property
    def device_bytes(self):
        b = ctypes.c_size_t()
        self.ctx._L.zc_msm_generators_info(self._g, None, None, ctypes.byref(b))
        return int(b.value)

    def msm(self, scalars_dev_ptr, out_dev_ptr, window_bits=None):
        self.ctx.check(self.ctx._L.zc_msm_gen_dev(self.ctx._h, self._g, scalars_dev_ptr, int(window_bits or self.window_bits), out_dev_ptr))

    def msm_partial(self, scalars_dev_ptr, out_dev_ptr, rank, nranks, window_bits=None):
        self.ctx.check(self.ctx._L.zc_msm_gen_partial_dev(self.ctx._h, self._g, scalars_dev_ptr, int(window_bits or self.window_bits),
                                                          int(rank), int(nranks), out_dev_ptr))

    def msm_sharded(self, scalars_dev_ptr, out_dev_ptr, window_bits=None):
        self.ctx.check(self.ctx._L.zc_msm_gen_sharded_dev(self.ctx._h, self._g, scalars_dev_ptr, int(window_bits or self.window_bits), out_dev_ptr))

    def close(self):
        if getattr(self, "_g", None) and getattr(self.ctx, "_h", None):
            self.ctx.check(self.ctx._L.zc_msm_generators_destroy(self.ctx._h, self._g))
        self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = None


def default_context():
    global _default
    if _default is None:
        _default = Context(0)
    return _default
