"""ctypes binding of libmmr_b200.so (C ABI: include/mmr_b200.h).

PyTorch is used for device memory and streams only; every compute call goes through the
C ABI with raw device pointers.  There is NO CPU fallback: if the library is missing or no
B200 is visible, calls raise ``NativeUnavailableError``.
"""

from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess
import threading

_PKG_DIR = pathlib.Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libmmr_b200.so"
CSRC_DIR = _PKG_DIR / "csrc"

LAYOUT_CELL_MAJOR = 0
LAYOUT_REFERENCE = 1
FIELD_B = 0
FIELD_H = 1


class NativeUnavailableError(RuntimeError):
    """libmmr_b200.so cannot be loaded or no CUDA device is usable."""


class NativeError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, status, message):
        super().__init__(f"libmmr_b200 status {status}: {message}")
        self.status = status


_c_double_p = ctypes.c_void_p  # device pointers travel as integers
_SIGNATURES = {
    "mmr_abi_version": (ctypes.c_int, []),
    "mmr_last_error": (ctypes.c_char_p, []),
    "mmr_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "mmr_ctx_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "mmr_ctx_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "mmr_ctx_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "mmr_ctx_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "mmr_ctx_profile": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
         ctypes.POINTER(ctypes.c_int64)],
    ),
    "mmr_ctx_profile_records": (
        ctypes.c_int64,
        [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int64],
    ),
    "mmr_ctx_profile_timeline": (
        ctypes.c_int64,
        [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
         ctypes.c_int64],
    ),
    "mmr_fp64_peaks": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "mmr_assemble": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, _c_double_p, _c_double_p]
        + [ctypes.c_int64] * 7
        + [_c_double_p, ctypes.c_double, ctypes.c_int, _c_double_p, ctypes.c_int64],
    ),
    "mmr_assemble_matched": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, _c_double_p, _c_double_p, ctypes.c_void_p,
         _c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)],
    ),
    "mmr_assemble_matched_rows": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, _c_double_p, _c_double_p, ctypes.c_void_p,
         _c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
         ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)],
    ),
    "mmr_lu_factor": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, ctypes.c_int64, ctypes.c_void_p,
         ctypes.POINTER(ctypes.c_int)],
    ),
    "mmr_lu_solve": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, ctypes.c_int64, ctypes.c_void_p, _c_double_p,
         ctypes.c_int64],
    ),
    "mmr_lu_diag_block": (
        ctypes.c_int,
        [ctypes.c_void_p, _c_double_p, ctypes.c_int64, ctypes.c_int, _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p,
         ctypes.POINTER(ctypes.c_int)],
    ),
    "mmr_lu_diag_block_clocks": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]),
    "mmr_dgemm_nn": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, _c_double_p,
         ctypes.c_int64, _c_double_p, ctypes.c_int64, ctypes.c_double, _c_double_p, ctypes.c_int64],
    ),
    "mmr_block_trsm": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, ctypes.c_int64, _c_double_p,
         ctypes.c_int64, ctypes.c_int],
    ),
    "mmr_gmres": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p,
         ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
         ctypes.POINTER(ctypes.c_double)],
    ),
    "mmr_gmres_bj": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p,
         ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
         ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)],
    ),
    "mmr_matvec": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, _c_double_p, ctypes.c_int64, _c_double_p,
         _c_double_p],
    ),
    "mmr_field": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
         ctypes.c_int64, _c_double_p, ctypes.c_int, _c_double_p],
    ),
    "mmr_field_currents": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, _c_double_p, ctypes.c_int64, _c_double_p, ctypes.c_int64,
         _c_double_p, ctypes.c_int, _c_double_p],
    ),
    "mmr_nccl_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "mmr_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "mmr_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "mmr_comm_p2p_disable": (ctypes.c_int, [ctypes.c_void_p]),
    "mmr_comm_p2p_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "mmr_comm_p2p_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "mmr_gmres_dist": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_double_p, ctypes.c_int64,
         _c_double_p, _c_double_p, ctypes.c_double, ctypes.c_int, ctypes.c_int,
         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)],
    ),
    "mmr_lu_solve_dist": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_double_p, ctypes.c_int64,
         _c_double_p, ctypes.POINTER(ctypes.c_int)],
    ),
    "mmr_lu_factor_dist": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_double_p, ctypes.c_int64,
         ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)],
    ),
    "mmr_lu_trisolve_dist": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_double_p, ctypes.c_int64,
         ctypes.c_void_p, _c_double_p],
    ),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lib_lock = threading.Lock()


def build(verbose=False):
    """Compile libmmr_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(CSRC_DIR), "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True, check=False)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        msg = f"building libmmr_b200.so failed (exit {res.returncode})"
        raise RuntimeError(msg)
    return LIB_PATH


def load_library():
    """dlopen the C-ABI library and attach prototypes (no GPU needed for this step)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            msg = (
                f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C magpylib_material_response_b200/csrc`. "
                "There is no CPU fallback for the demagnetization path."
            )
            raise NativeUnavailableError(msg)
        try:
            lib = ctypes.CDLL(str(LIB_PATH))
        except OSError as e:
            msg = f"cannot load {LIB_PATH}: {e}"
            raise NativeUnavailableError(msg) from e
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def _check(lib, status):
    if status != 0:
        raise NativeError(status, lib.mmr_last_error().decode(errors="replace"))


def _ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


class Context:
    """One C-ABI context = (device, stream).  Thin 1:1 wrapper over the exported functions."""

    def __init__(self, device=0, stream=None):
        import torch

        if not torch.cuda.is_available():
            msg = (
                "No CUDA device is visible: the demagnetization path runs only on a B200 (sm_100a) "
                "through libmmr_b200.so and has no CPU fallback."
            )
            raise NativeUnavailableError(msg)
        self.lib = load_library()
        self.torch = torch
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index or 0)
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        handle = ctypes.c_void_p()
        _check(
            self.lib,
            self.lib.mmr_ctx_create(
                self.device.index, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(handle)
            ),
        )
        self.handle = handle
        self.nranks = 1
        self.rank = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mmr_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ---------------------------------------------------------------------
    def to_device(self, arr):
        import numpy as np

        t = self.torch.from_numpy(np.array(arr, dtype=np.float64, order="C", copy=True))
        return t.to(self.device, non_blocking=False)

    def empty(self, *shape, dtype=None):
        return self.torch.empty(*shape, dtype=dtype or self.torch.float64, device=self.device)

    def sync(self):
        _check(self.lib, self.lib.mmr_ctx_sync(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.mmr_ctx_launch_count(self.handle))

    PROF_KINDS = ("assemble", "gemm", "panel", "trsm", "laswp", "matvec", "panel_gemm", "trisolve", "comm")

    def set_profiling(self, on=True, kinds=None):
        """Record CUDA-event pairs around the native launches (clears earlier records).  kinds: only these
        kernel classes (names of PROF_KINDS) -- every recorded launch costs two event records of host time."""
        flag = int(bool(on))
        if on and kinds is not None:
            flag = 0
            for name in kinds:
                flag |= 1 << (self.PROF_KINDS.index(name) + 1)
        _check(self.lib, self.lib.mmr_ctx_set_profiling(self.handle, flag))

    def profile(self):
        """{kind: (ms, units, launches)} for every kernel class recorded since set_profiling."""
        out = {}
        for k, name in enumerate(self.PROF_KINDS):
            ms, units, cnt = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int64(0)
            _check(self.lib, self.lib.mmr_ctx_profile(self.handle, k, ctypes.byref(ms), ctypes.byref(units),
                                                      ctypes.byref(cnt)))
            out[name] = (ms.value, units.value, cnt.value)
        return out

    def profile_timeline(self, cap=200000):
        """[(kind name, start ms, duration ms)] of every record since set_profiling(True), launch order."""
        kinds = (ctypes.c_int * cap)()
        t0 = (ctypes.c_double * cap)()
        dt = (ctypes.c_double * cap)()
        n = self.lib.mmr_ctx_profile_timeline(self.handle, kinds, t0, dt, cap)
        tag = ("", "@side", "@comm", "@aux")
        return [(self.PROF_KINDS[kinds[i] & 0xFF] + tag[(kinds[i] >> 8) & 3], t0[i], dt[i]) for i in range(n)]

    def profile_records(self, kind, cap=100000):
        buf = (ctypes.c_double * cap)()
        n = self.lib.mmr_ctx_profile_records(self.handle, self.PROF_KINDS.index(kind), buf, cap)
        return list(buf[:n])

    def fp64_peaks(self):
        """(FMA-pipe TFLOP/s, DMMA tensor-pipe TFLOP/s) of this device, measured now."""
        fma, mma = ctypes.c_double(0), ctypes.c_double(0)
        _check(self.lib, self.lib.mmr_fp64_peaks(self.handle, ctypes.byref(fma), ctypes.byref(mma)))
        return fma.value, mma.value

    # ---- kernel 1 --------------------------------------------------------------------
    def assemble(self, pos, quat, dim, out, ld, *, chi=None, max_dist=0.0, layout=LAYOUT_CELL_MAJOR,
                 src_range=None, tgt_first=0, n_tgt_local=None, tgt_block=None, tgt_nparts=1,
                 tgt_part=0):
        n = pos.shape[0]
        s0, s1 = src_range if src_range is not None else (0, n)
        if n_tgt_local is None:
            n_tgt_local = n
        if tgt_block is None:
            tgt_block = max(1, n_tgt_local)
        _check(
            self.lib,
            self.lib.mmr_assemble(
                self.handle, n, _ptr(pos), _ptr(quat), _ptr(dim), s0, s1, tgt_first, n_tgt_local,
                tgt_block, tgt_nparts, tgt_part, _ptr(chi), float(max_dist), layout, _ptr(out), ld,
            ),
        )

    def assemble_matched(self, pos, quat, dim, cls, out, ld, *, chi=None, table_cap_slots=1 << 24,
                         n_tgt_local=None, tgt_block=None, tgt_nparts=1, tgt_part=0):
        """pairs_matching assembly; with n_tgt_local / tgt_block / tgt_nparts / tgt_part only the rows of
        this rank's block-cyclic target set (multi-GPU)."""
        n = pos.shape[0]
        nuniq = ctypes.c_int64(0)
        if n_tgt_local is None:
            n_tgt_local, tgt_block = n, max(1, n)
        status = self.lib.mmr_assemble_matched_rows(
            self.handle, n, _ptr(pos), _ptr(quat), _ptr(dim), _ptr(cls), _ptr(chi), _ptr(out), ld,
            table_cap_slots, n_tgt_local, tgt_block, tgt_nparts, tgt_part, ctypes.byref(nuniq),
        )
        if status == -3:
            return None
        _check(self.lib, status)
        return int(nuniq.value)

    # ---- kernel 2a -------------------------------------------------------------------
    def lu_factor(self, Q, ld, ipiv):
        info = ctypes.c_int(0)
        _check(self.lib, self.lib.mmr_lu_factor(self.handle, Q.shape[0], _ptr(Q), ld, _ptr(ipiv),
                                                ctypes.byref(info)))
        return info.value

    def lu_solve(self, LU, ld, ipiv, B, nrhs=1):
        _check(self.lib, self.lib.mmr_lu_solve(self.handle, LU.shape[0], _ptr(LU), ld, _ptr(ipiv),
                                               _ptr(B), nrhs))

    def lu_diag_block(self, A, ld, nb, S, lds, Linv, Uinv):
        """No-pivot LU of the leading nb x nb block + both inverses; returns True if the speculation is rejected."""
        rej = ctypes.c_int(0)
        _check(self.lib, self.lib.mmr_lu_diag_block(self.handle, _ptr(A), ld, nb, _ptr(S), lds, _ptr(Linv), _ptr(Uinv),
                                                    ctypes.byref(rej)))
        return bool(rej.value)

    def lu_diag_block_clocks(self):
        buf = (ctypes.c_longlong * 30)()
        _check(self.lib, self.lib.mmr_lu_diag_block_clocks(self.handle, buf))
        return list(buf)

    def dgemm_nn(self, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
        _check(self.lib, self.lib.mmr_dgemm_nn(self.handle, m, n, k, alpha, _ptr(A), lda, _ptr(B), ldb,
                                               beta, _ptr(C), ldc))

    def block_trsm(self, nb0, nb1, Linv0, Linv1, L10, ldl, B, ldb, ncols):
        """B <- [L00 0; L10 L11]^-1 B in one launch (column-major operands, see include/mmr_b200.h)."""
        _check(self.lib, self.lib.mmr_block_trsm(self.handle, nb0, nb1, _ptr(Linv0), _ptr(Linv1), _ptr(L10), ldl,
                                                 _ptr(B), ldb, ncols))

    # ---- kernel 2b -------------------------------------------------------------------
    def gmres(self, Q, ld, b, x, tol=1e-10, restart=200, maxit=2000, raise_on_fail=True):
        iters = ctypes.c_int(0)
        relres = ctypes.c_double(0.0)
        status = self.lib.mmr_gmres(self.handle, b.shape[0], _ptr(Q), ld, _ptr(b), _ptr(x), tol, restart,
                                    maxit, ctypes.byref(iters), ctypes.byref(relres))
        if status != 0 and (raise_on_fail or status != -4):
            _check(self.lib, status)
        return iters.value, relres.value

    def gmres_bj(self, Q, ld, b, x, blocks, tol=1e-10, restart=200, maxit=2000, raise_on_fail=True):
        """GMRES with a block-Jacobi right preconditioner; blocks = [(first row, rows), ...]."""
        iters = ctypes.c_int(0)
        relres = ctypes.c_double(0.0)
        nb = len(blocks)
        r0 = (ctypes.c_int64 * max(nb, 1))(*[int(b0) for b0, _ in blocks])
        rn = (ctypes.c_int64 * max(nb, 1))(*[int(n) for _, n in blocks])
        status = self.lib.mmr_gmres_bj(self.handle, b.shape[0], _ptr(Q), ld, _ptr(b), _ptr(x), tol, restart, maxit,
                                       nb, r0, rn, ctypes.byref(iters), ctypes.byref(relres))
        if status != 0 and (raise_on_fail or status != -4):
            _check(self.lib, status)
        return iters.value, relres.value

    def matvec(self, Q, ld, x, y, nrows=None, ncols=None):
        nrows = Q.shape[0] if nrows is None else nrows
        ncols = x.shape[0] if ncols is None else ncols
        _check(self.lib, self.lib.mmr_matvec(self.handle, nrows, ncols, _ptr(Q), ld, _ptr(x), _ptr(y)))

    # ---- field -----------------------------------------------------------------------
    def field(self, pos, quat, dim, pol, obs, out, field=FIELD_B):
        _check(self.lib, self.lib.mmr_field(self.handle, pos.shape[0], _ptr(pos), _ptr(quat), _ptr(dim),
                                            _ptr(pol), obs.shape[0], _ptr(obs), field, _ptr(out)))

    def field_currents(self, obs, seg, loop, out, field=FIELD_B):
        """seg: (nseg,7) device tensor or None; loop: (nloop,9) device tensor or None."""
        nseg = 0 if seg is None else seg.shape[0]
        nloop = 0 if loop is None else loop.shape[0]
        _check(self.lib, self.lib.mmr_field_currents(self.handle, obs.shape[0], _ptr(obs), nseg,
                                                     _ptr(seg) if nseg else None, nloop,
                                                     _ptr(loop) if nloop else None, field, _ptr(out)))

    # ---- multi-GPU -------------------------------------------------------------------
    def comm_init_from_torch(self):
        """Create the context's NCCL communicator; the 128-byte id is broadcast with
        torch.distributed (the process group is plumbing only)."""
        import torch.distributed as dist

        from .distributed import broadcast_bytes

        world, rank = dist.get_world_size(), dist.get_rank()
        idbuf = (ctypes.c_char * 128)()
        if rank == 0:
            _check(self.lib, self.lib.mmr_nccl_unique_id(ctypes.cast(idbuf, ctypes.c_void_p)))
        idbuf = ctypes.create_string_buffer(broadcast_bytes(bytes(idbuf), src=0), 128)
        _check(self.lib, self.lib.mmr_comm_init(self.handle, ctypes.cast(idbuf, ctypes.c_void_p), world, rank))
        self.nranks, self.rank = world, rank
        # peer-memory buffers of the triangular sweeps (CUDA IPC): handles travel over torch.distributed
        self.p2p = False
        if world <= 16 and os.environ.get("MMR_P2P", "1") != "0":
            mine = (ctypes.c_char * 64)()
            ok = self.lib.mmr_comm_p2p_export(self.handle, ctypes.cast(mine, ctypes.c_void_p)) == 0
            table = [None] * world
            dist.all_gather_object(table, bytes(mine) if ok else None)
            if all(h is not None for h in table):
                blob = ctypes.create_string_buffer(b"".join(table), 64 * world)
                ok = self.lib.mmr_comm_p2p_open(self.handle, ctypes.cast(blob, ctypes.c_void_p), world, rank) == 0
            else:
                ok = False
            # every rank must take the same path
            votes = [None] * world
            dist.all_gather_object(votes, bool(ok))
            self.p2p = all(votes)
            if not self.p2p:
                self.lib.mmr_comm_p2p_disable(self.handle)

    def gmres_dist(self, N, nrows_local, tgt_block, Qlocal, ld, b, x, tol=1e-10, restart=200, maxit=2000):
        iters = ctypes.c_int(0)
        relres = ctypes.c_double(0.0)
        _check(self.lib, self.lib.mmr_gmres_dist(self.handle, N, nrows_local, tgt_block, _ptr(Qlocal), ld,
                                                 _ptr(b), _ptr(x), tol, restart, maxit,
                                                 ctypes.byref(iters), ctypes.byref(relres)))
        return iters.value, relres.value

    def lu_solve_dist(self, N, nrows_local, tgt_block, Qlocal, ld, b):
        info = ctypes.c_int(0)
        _check(self.lib, self.lib.mmr_lu_solve_dist(self.handle, N, nrows_local, tgt_block, _ptr(Qlocal),
                                                    ld, _ptr(b), ctypes.byref(info)))
        return info.value


    def lu_factor_dist(self, N, nrows_local, tgt_block, Qlocal, ld, ipiv):
        info = ctypes.c_int(0)
        _check(self.lib, self.lib.mmr_lu_factor_dist(self.handle, N, nrows_local, tgt_block, _ptr(Qlocal), ld,
                                                     _ptr(ipiv), ctypes.byref(info)))
        return info.value

    def lu_trisolve_dist(self, N, nrows_local, tgt_block, Qlocal, ld, ipiv, b):
        _check(self.lib, self.lib.mmr_lu_trisolve_dist(self.handle, N, nrows_local, tgt_block, _ptr(Qlocal), ld,
                                                       _ptr(ipiv), _ptr(b)))


_default_ctx = {}


def default_context(device=None):
    """Process-wide context per device on torch's current stream."""
    import torch

    if not torch.cuda.is_available():
        msg = (
            "No CUDA device is visible: the demagnetization path runs only on a B200 (sm_100a) "
            "through libmmr_b200.so and has no CPU fallback."
        )
        raise NativeUnavailableError(msg)
    idx = torch.cuda.current_device() if device is None else int(device)
    # one context per (device, torch current stream): the torch copies / allocations around a native call
    # run on the current stream, so the kernels must be enqueued there too (inside `with
    # torch.cuda.stream(s)` the legacy default stream would be unordered with them)
    stream = torch.cuda.current_stream(idx)
    key = (idx, stream.cuda_stream)
    ctx = _default_ctx.get(key)
    if ctx is None:
        ctx = Context(idx, stream)
        if key[1] != 0:
            # communicator state is per process group, not per stream: share the default-stream context's
            base = _default_ctx.get((idx, 0))
            if base is not None and base.nranks > 1:
                msg = "multi-GPU calls must be made on the default stream (the NCCL communicator lives in that context)"
                raise NativeError(-1, msg)
        _default_ctx[key] = ctx
    return ctx
