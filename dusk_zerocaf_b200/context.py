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


    # ---- fixed generators --------------------------------------------------------------------------------------
    def msm_prepare_points(self, points_dev_ptr, n):
        """zc_msm_prepare_points_dev: cache the MSM operands of a resident point set (Z = 1, cached form)."""
        self.check(self._L.zc_msm_prepare_points_dev(self._h, points_dev_ptr, int(n)))

    def msm_prepare_fixed_base(self, points_dev_ptr, n, window_bits=16, rank=0, nranks=1):
        """zc_msm_prepare_fixed_base_dev: pre-scaled rows 2^(c w) P_i for the windows rank owns (one merged bucket set and
        no doubling chain in later MSM calls of this shape)."""
        self.check(self._L.zc_msm_prepare_fixed_base_dev(self._h, points_dev_ptr, int(n), int(window_bits), int(rank), int(nranks)))

    def msm_forget_points(self):
        self.check(self._L.zc_msm_forget_points(self._h))


_default = None


def default_context():
    global _default
    if _default is None:
        _default = Context(0)
    return _default
