"""kube_throttler_b200 -- B200-native throttle-admission hot path of everpeace/kube-throttler.

The product is the C-ABI shared library `libkt_b200.so` (include/kt_b200.h; CUDA kernels in csrc/).
This Python module is a thin ctypes binding used by the tests and bench.py -- the same calls a Go
plugin would make through cgo.  There is no CPU fallback anywhere in this package:
if the library or a GPU is missing, `Engine(...)` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from . import abi
from .abi import PassResult, PodCols, Snapshot  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KT_B200_LIB") or os.path.join(_HERE, "libkt_b200.so")  # override: kernel-variant experiments only
_lib = None


class KtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kt error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in ("kt_engine.cu", "kt_tables.cc", "kt_host.cc", "kt_kernels.cuh", "kt_tables.h", "kt_json.h", "kt_quantity.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "kt_b200.h"))
    srcs.append(os.path.join(_HERE, "..", "include", "kt_host.h"))
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-B", "NVCCFLAGS_EXTRA="], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return LIB_PATH


EXPORTS = [
    "kt_create", "kt_destroy", "kt_last_error", "kt_version", "kt_set_stream", "kt_sync", "kt_enable_timing",
    "kt_enable_trace", "kt_get_trace", "kt_host_alloc", "kt_host_alloc_upload", "kt_host_free", "kt_upload_pods", "kt_upload_pods_compact", "kt_upload_pods_packed", "kt_set_async_uploads", "kt_update_pod_rows", "kt_upload_namespaces",
    "kt_upload_throttles", "kt_upload_status", "kt_set_reserved", "kt_evaluate", "kt_get_reconcile",
    "kt_match_words", "kt_get_match_bitmap", "kt_get_match_rows", "kt_get_check", "kt_set_sparse_check", "kt_get_check_sparse", "kt_get_check_rows", "kt_get_changed", "kt_get_reconcile_rows", "kt_step_submit", "kt_step_wait", "kt_admit_queue", "kt_get_timing", "kt_comm_unique_id",
    "kt_comm_init", "kt_comm_destroy", "kt_debug_compile_tables",
]


def lib():
    """Load libkt_b200.so (must have been built: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KtError(abi.ERR_STATE, f"{LIB_PATH} is missing -- build it first (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.kt_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(abi.Limits)]
        L.kt_destroy.argtypes = [vp]
        L.kt_destroy.restype = None
        L.kt_last_error.argtypes = [vp]
        L.kt_last_error.restype = C.c_char_p
        L.kt_version.restype = C.c_char_p
        L.kt_set_stream.argtypes = [vp, vp]
        L.kt_sync.argtypes = [vp]
        L.kt_enable_timing.argtypes = [vp, C.c_int]
        L.kt_enable_trace.argtypes = [vp, C.c_int]
        L.kt_get_trace.argtypes = [vp, vp, C.c_int64, vp]
        L.kt_get_trace.restype = C.c_int64
        L.kt_host_alloc.argtypes = [C.c_size_t]
        L.kt_host_alloc.restype = vp
        L.kt_host_alloc_upload.argtypes = [C.c_size_t]
        L.kt_host_alloc_upload.restype = vp
        L.kt_host_free.argtypes = [vp]
        L.kt_host_free.restype = None
        L.kt_upload_pods.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp]
        L.kt_set_async_uploads.argtypes = [vp, C.c_int]
        L.kt_upload_pods_compact.argtypes = [vp, C.c_int, C.c_int64, C.c_int32, vp, vp, vp, vp, vp]
        L.kt_upload_pods_packed.argtypes = [vp, C.c_int, C.c_int64, C.POINTER(abi.PackedPodsStruct)]
        L.kt_set_sparse_check.argtypes = [vp, C.c_int64]
        L.kt_get_check_sparse.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(C.c_int64)]
        L.kt_step_submit.argtypes = [vp, C.c_int64, C.POINTER(abi.PackedPodsStruct), C.c_int64, C.POINTER(abi.PackedPodsStruct), C.c_int64, C.c_uint32]
        L.kt_step_wait.argtypes = [vp, C.POINTER(abi.StepResult)]
        L.kt_admit_queue.argtypes = [vp, C.c_int64, C.c_int64, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.kt_get_check_rows.argtypes = [vp, C.c_int64, vp, vp, vp]
        L.kt_get_changed.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int64), vp]
        L.kt_get_reconcile_rows.argtypes = [vp, C.c_int64, vp, C.POINTER(abi.ReconcileOut)]
        L.kt_update_pod_rows.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp]
        L.kt_upload_namespaces.argtypes = [vp, C.c_int32, vp]
        L.kt_upload_throttles.argtypes = [vp, C.c_int32, C.POINTER(abi.ThrottleCols), C.POINTER(abi.SelectorTable)]
        L.kt_upload_status.argtypes = [vp, C.POINTER(abi.StatusCols)]
        L.kt_set_reserved.argtypes = [vp, vp, vp, vp]
        L.kt_evaluate.argtypes = [vp, C.c_int64, C.c_uint32]
        L.kt_get_reconcile.argtypes = [vp, C.POINTER(abi.ReconcileOut)]
        L.kt_match_words.argtypes = [vp]
        L.kt_match_words.restype = C.c_int32
        L.kt_get_match_bitmap.argtypes = [vp, C.c_int, vp]
        L.kt_get_check.argtypes = [vp, vp, vp]
        L.kt_get_timing.argtypes = [vp, C.POINTER(abi.Timing)]
        L.kt_comm_unique_id.argtypes = [vp]
        L.kt_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
        L.kt_comm_destroy.argtypes = [vp]
        L.kt_debug_compile_tables.argtypes = [C.POINTER(abi.Limits), C.c_int32, C.POINTER(abi.ThrottleCols), C.POINTER(abi.SelectorTable), C.c_int32, vp,
                                              vp, vp, vp, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


class Pinned:
    """Pinned (cudaHostAlloc) host buffer exposed as a numpy array: `.array`.  Keep the object alive
    while the array is in use."""

    def __init__(self, shape, dtype, upload_only: bool = False):
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 1)
        self._ptr = (lib().kt_host_alloc_upload if upload_only else lib().kt_host_alloc)(nbytes)  # upload_only: write-combined
        if not self._ptr:
            raise KtError(abi.ERR_CUDA, "kt_host_alloc failed")
        self._buf = (C.c_char * nbytes).from_address(self._ptr)
        self.array = np.frombuffer(self._buf, dtype=dtype, count=count).reshape(shape)

    def __del__(self):
        try:
            if self._ptr:
                lib().kt_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


class Engine:
    """One context == one GPU (one row shard in a multi-GPU run)."""

    def __init__(self, R: int, L: int, LN: int, device: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        self.R, self.Lslots, self.LN = R, L, LN
        lim = abi.Limits(abi.ABI_VERSION, R, L, LN)
        rc = self._L.kt_create(C.byref(self._h), device, C.byref(lim))
        if rc != 0:
            self._h = C.c_void_p()
            raise KtError(rc, "kt_create failed (no CUDA device? bad limits?) -- there is no CPU fallback")
        self.m = 0
        self.n = [0, 0]
        self._keep = []

    # -- plumbing -----------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise KtError(rc, self._L.kt_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.kt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: Optional[int]):
        self._ck(self._L.kt_set_stream(self._h, cuda_stream))

    def enable_timing(self, on: bool = True):
        self._ck(self._L.kt_enable_timing(self._h, int(on)))

    def enable_trace(self, on: bool = True):
        self._ck(self._L.kt_enable_trace(self._h, int(on)))

    def trace(self):
        """(rows[n][32] = ticket, sm, start_ns, end_ns, 12 stage stamps (4.. per role, see kt_kernels.cuh), 16 cycle counts; roles[4] = tiles per role)
        of the last fused pass."""
        roles = np.zeros(4, np.uint32)
        n = self._L.kt_get_trace(self._h, None, 0, roles.ctypes.data)
        rows = np.zeros((int(roles.sum()), 32), np.uint64)
        n = self._L.kt_get_trace(self._h, rows.ctypes.data, rows.shape[0], roles.ctypes.data)
        if n < 0:
            self._ck(int(n))
        return rows[: int(n)], roles

    def sync(self):
        self._ck(self._L.kt_sync(self._h))

    # -- uploads ------------------------------------------------------------------------------
    def upload_pods(self, kind: int, pods: PodCols):
        self._ck(self._L.kt_upload_pods(self._h, kind, pods.n, abi.ptr(pods.labels), abi.ptr(pods.req), abi.ptr(pods.present),
                                        abi.ptr(pods.flags), abi.ptr(pods.ns_id)))
        self.n[kind] = pods.n

    def set_async_uploads(self, on: bool = True):
        """Uploads return once queued; the host buffers must stay untouched until sync() / a getter returns."""
        self._ck(self._L.kt_set_async_uploads(self._h, int(on)))

    def upload_pods_compact(self, kind: int, cp: "abi.CompactPodCols"):
        """kt_upload_pods_compact: 56 instead of 108 bytes per row over the host link; expanded to the same int64 columns on the device."""
        self._ck(self._L.kt_upload_pods_compact(self._h, kind, cp.n, cp.val_bits, abi.ptr(cp.labels32), abi.ptr(cp.req32), abi.ptr(cp.req_shift),
                                                abi.ptr(cp.present), abi.ptr(cp.meta)))
        self.n[kind] = cp.n

    def upload_pods_packed(self, kind: int, pk: "abi.PackedPodCols"):
        """kt_upload_pods_packed: 36 bytes per row (L=8, R=4): 16-bit label-pair indices, presence inside the meta word."""
        st = pk.struct()
        self._ck(self._L.kt_upload_pods_packed(self._h, kind, pk.n, C.byref(st)))
        self.n[kind] = pk.n

    def update_pod_rows(self, kind: int, rows: np.ndarray, pods: PodCols):
        rows = np.ascontiguousarray(rows, np.int64)
        self._ck(self._L.kt_update_pod_rows(self._h, kind, rows.shape[0], abi.ptr(rows), abi.ptr(pods.labels), abi.ptr(pods.req),
                                            abi.ptr(pods.present), abi.ptr(pods.flags), abi.ptr(pods.ns_id)))

    def upload_namespaces(self, ns_labels: np.ndarray):
        self._ck(self._L.kt_upload_namespaces(self._h, ns_labels.shape[1], abi.ptr(ns_labels)))

    def upload_throttles(self, snap: Snapshot):
        cols, sel = snap.throttle_cols(), snap.selector_table()
        self._ck(self._L.kt_upload_throttles(self._h, snap.m, C.byref(cols), C.byref(sel)))
        self.m = snap.m

    def upload_status(self, snap: Snapshot):
        st = snap.status_cols()
        self._ck(self._L.kt_upload_status(self._h, C.byref(st)))

    def set_reserved(self, snap: Snapshot):
        self._ck(self._L.kt_set_reserved(self._h, abi.ptr(snap.reserved), abi.ptr(snap.reserved_present), abi.ptr(snap.reserved_cnt)))

    def upload_snapshot(self, snap: Snapshot):
        """Everything a pass needs, in dependency order."""
        snap.normalize()
        self.upload_namespaces(snap.ns_labels)
        self.upload_throttles(snap)
        if snap.reserved is not None:
            self.set_reserved(snap)
        if snap.status is not None:
            self.upload_status(snap)
        self.upload_pods(abi.PODS_RUNNING, snap.running)
        self.upload_pods(abi.PODS_PENDING, snap.pending)

    # -- the pass -------------------------------------------------------------------------------
    def evaluate(self, now: int, flags: int = abi.EVAL_FRESH_STATUS):
        self._ck(self._L.kt_evaluate(self._h, now, flags))

    @property
    def words_per_row(self) -> int:
        return int(self._L.kt_match_words(self._h))

    def timing(self) -> abi.Timing:
        t = abi.Timing()
        self._ck(self._L.kt_get_timing(self._h, C.byref(t)))
        return t

    # -- downloads ------------------------------------------------------------------------------
    def download(self, out: Optional[PassResult] = None, bitmaps: bool = True) -> PassResult:
        W = self.words_per_row
        if out is None:
            z = np.zeros
            m, R = self.m, self.R
            N, P = self.n
            out = PassResult(W, z((R, m), np.int64), z(m, np.uint32), z(m, np.int64), z(m, np.uint32), z((R, m), np.int64),
                             z(m, np.uint32), z(m, np.int64), z(m, np.uint8), z((N if bitmaps else 0, W), np.uint32),
                             z((P if bitmaps else 0, W), np.uint32), z((P, 2 * W), np.uint32), z(P, np.uint8))
        rec = out.reconcile_out()
        self._ck(self._L.kt_get_reconcile(self._h, C.byref(rec)))
        if bitmaps:
            if self.n[0]:
                self._ck(self._L.kt_get_match_bitmap(self._h, abi.PODS_RUNNING, abi.ptr(out.run_bitmap)))
            if self.n[1]:
                self._ck(self._L.kt_get_match_bitmap(self._h, abi.PODS_PENDING, abi.ptr(out.pend_bitmap)))
        self._ck(self._L.kt_get_check(self._h, abi.ptr(out.codes), abi.ptr(out.admit)))
        return out

    def get_check(self, codes: Optional[np.ndarray], admit: Optional[np.ndarray]):
        self._ck(self._L.kt_get_check(self._h, abi.ptr(codes), abi.ptr(admit)))

    def set_sparse_check(self, cap_entries: int):
        """Later passes also append every non-zero code word to a device list (0: off)."""
        self._ck(self._L.kt_set_sparse_check(self._h, cap_entries))

    def get_check_sparse(self, admit: Optional[np.ndarray], entries: np.ndarray) -> int:
        """admit[p] and the non-zero code words as rows {pending row, word index, codes}; returns their total number
        (more than entries.shape[0]: read the dense rows with get_check instead)."""
        n = C.c_int64(0)
        self._ck(self._L.kt_get_check_sparse(self._h, abi.ptr(admit), abi.ptr(entries), entries.shape[0], C.byref(n)))
        return int(n.value)

    # -- one end-to-end step in one call ----------------------------------------------------------
    def step_submit(self, running, pending, now: int, flags: int = abi.EVAL_FRESH_STATUS):
        """Queue packed uploads (abi.PackedPodCols or a prepared (n, struct) pair; None keeps the resident rows) + the pass + the
        result copies; returns at once.  The host columns must stay untouched until step_wait()."""
        def prep(x):
            if x is None:
                return 0, None
            if isinstance(x, tuple):
                return x
            return x.n, x.struct()
        (nr, sr), (np_, sp) = prep(running), prep(pending)
        self._ck(self._L.kt_step_submit(self._h, nr, C.byref(sr) if sr is not None else None, np_, C.byref(sp) if sp is not None else None, now, flags))
        if running is not None:
            self.n[abi.PODS_RUNNING] = nr
        if pending is not None:
            self.n[abi.PODS_PENDING] = np_

    def step_wait(self) -> "abi.StepResult":
        """Block until the submitted step's results have landed; the returned struct points into the library's pinned block
        (valid until the next step_submit): .admit [n_pending] u8, .entries [n_sparse][3] u32, .status columns."""
        res = abi.StepResult()
        self._ck(self._L.kt_step_wait(self._h, C.byref(res)))
        return res

    def admit_queue(self, first: int, count: int, flags: int = 0):
        """Queue-ordered greedy admission of the pending rows [first, first + count) on the device: (rounds, admitted)."""
        rounds, admitted = C.c_int32(0), C.c_int64(0)
        self._ck(self._L.kt_admit_queue(self._h, first, count, flags, C.byref(rounds), C.byref(admitted)))
        return int(rounds.value), int(admitted.value)

    def get_check_rows(self, rows: np.ndarray):
        """(codes[k][2W], admit[k]) of the listed pending rows."""
        rows = np.ascontiguousarray(rows, np.int64)
        codes = np.zeros((rows.shape[0], 2 * self.words_per_row), np.uint32)
        admit = np.zeros(rows.shape[0], np.uint8)
        self._ck(self._L.kt_get_check_rows(self._h, rows.shape[0], abi.ptr(rows), abi.ptr(codes), abi.ptr(admit)))
        return codes, admit

    def get_changed(self):
        """(sorted indices of the throttles whose status this pass changes, flags[m]) -- the device-side diff against kt_upload_status."""
        idx = np.zeros(max(self.m, 1), np.int32)
        flags = np.zeros(max(self.m, 1), np.uint8)
        n = C.c_int64(0)
        self._ck(self._L.kt_get_changed(self._h, abi.ptr(idx), idx.shape[0], C.byref(n), abi.ptr(flags)))
        return np.sort(idx[: int(n.value)]), flags[: self.m]

    def get_reconcile_rows(self, idx: np.ndarray) -> PassResult:
        """Status columns of the listed throttles only (shaped for len(idx) throttles)."""
        idx = np.ascontiguousarray(idx, np.int32)
        k, R, z = idx.shape[0], self.R, np.zeros
        out = PassResult(self.words_per_row, z((R, k), np.int64), z(k, np.uint32), z(k, np.int64), z(k, np.uint32), z((R, k), np.int64), z(k, np.uint32),
                         z(k, np.int64), z(k, np.uint8), z((0, 1), np.uint32), z((0, 1), np.uint32), z((0, 1), np.uint32), z(0, np.uint8))
        rec = out.reconcile_out()
        self._ck(self._L.kt_get_reconcile_rows(self._h, k, abi.ptr(idx), C.byref(rec)))
        return out

    def get_reconcile(self, out: PassResult):
        rec = out.reconcile_out()
        self._ck(self._L.kt_get_reconcile(self._h, C.byref(rec)))

    # -- multi-GPU --------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = lib().kt_comm_unique_id(buf)
        if rc != 0:
            raise KtError(rc, "kt_comm_unique_id failed (libnccl.so.2 missing?)")
        return bytes(buf)

    def comm_init(self, uid: bytes, nranks: int, rank: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self._L.kt_comm_init(self._h, buf, nranks, rank))

    def comm_destroy(self):
        self._ck(self._L.kt_comm_destroy(self._h))


def evaluate_snapshot(snap: Snapshot, flags: int = abi.EVAL_FRESH_STATUS, device: int = 0) -> PassResult:
    """Convenience: upload + one pass + download on one GPU."""
    eng = Engine(snap.R, snap.L, snap.LN, device)
    try:
        eng.upload_snapshot(snap)
        eng.evaluate(snap.now, flags)
        return eng.download()
    finally:
        eng.close()
