"""
ctypes binding of liblwb200.so (include/lwb200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is called
without a CUDA device, the call raises.  Host-only entry points (Fock rank/unrank/count) work
without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LWB_LIB_PATH") or os.path.join(_HERE, "liblwb200.so")  # override: experiment builds

LWB_C128, LWB_C64 = 0, 1
LWB_E_ARG, LWB_E_PHOTONS, LWB_E_LIMIT, LWB_E_CUDA, LWB_E_NOMEM, LWB_E_COMM = -1, -2, -3, -4, -5, -6


class BackendError(Exception):
    """Mirror of lightworks.emulator.utils.exceptions.BackendError (emulator/utils/exceptions.py:34);
    rebound to the reference's class by lightworks_b200.install() when lightworks is importable."""


class PhotonNumberError(Exception):
    """Mirror of lightworks.sdk.utils.exceptions.PhotonNumberError (sdk/utils/exceptions.py:109)."""


_exc = {"BackendError": BackendError, "PhotonNumberError": PhotonNumberError}

# every symbol include/lwb200.h declares: name -> (restype, argtypes)
_vp, _i, _u64, _dbl = C.c_void_p, C.c_int, C.c_uint64, C.c_double
_pi, _pu64, _pdbl = C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_double)
SIGNATURES = {
    "lwb_create": (_i, [_i, C.POINTER(_vp)]),
    "lwb_destroy": (_i, [_vp]),
    "lwb_last_error": (C.c_char_p, []),
    "lwb_version": (_i, []),
    "lwb_set_stream": (_i, [_vp, _vp, _i]),
    "lwb_synchronize": (_i, [_vp]),
    "lwb_device_info": (_i, [_vp, _pi, _pi, C.c_char_p, _i]),
    "lwb_set_tuning": (_i, [_i, _i]),
    "lwb_set_option": (_i, [C.c_char_p, _i]),
    "lwb_multiplicity_plan": (_i, [_i, _i, _pi, _pi, _pi, _pi, _pi, _pdbl]),
    "lwb_launch_count": (_i, [_vp, _pu64]),
    "lwb_slos_plan_check": (_i, [_i, _i, _pu64, _pu64, _pu64]),
    "lwb_fock_count": (_i, [_i, _i, _pu64]),
    "lwb_fock_rank": (_i, [_i, _vp, _pu64]),
    "lwb_fock_unrank": (_i, [_i, _i, _u64, _vp]),
    "lwb_slos_rank": (_i, [_i, _vp, _pu64]),
    "lwb_slos_unrank": (_i, [_i, _i, _u64, _vp]),
    "lwb_fock_basis": (_i, [_vp, _i, _i, _u64, _u64, _vp]),
    "lwb_fock_basis_dev": (_i, [_vp, _i, _i, _u64, _u64, _vp]),
    "lwb_perm_matrices": (_i, [_vp, _vp, _u64, _i, _i, _vp]),
    "lwb_perm_matrices_dev": (_i, [_vp, _vp, _u64, _i, _i, _vp]),
    "lwb_perm_amplitudes": (_i, [_vp, _vp, _i, _vp, _i, _vp, _u64, _i, _i, _vp]),
    "lwb_perm_amplitudes_batch": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _u64, _i, _i, _vp]),
    "lwb_perm_amplitudes_dev": (_i, [_vp, _vp, _i, _vp, _i, _vp, _u64, _i, _i, _i, _vp]),
    "lwb_perm_basis_probs": (_i, [_vp, _vp, _i, _vp, _u64, _u64, _i, _vp]),
    "lwb_perm_basis_probs_dev": (_i, [_vp, _vp, _i, _vp, _i, _u64, _u64, _i, _vp]),
    "lwb_perm_single": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "lwb_perm_single_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "lwb_slos_amplitudes": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "lwb_slos_amplitudes_dev": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "lwb_slos_basis_probs": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "lwb_slos_basis_probs_dev": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "lwb_pdist": (_i, [_vp, _vp, _i, _i, _vp, _dbl, _i, _i, _vp, _vp, _u64, _pu64]),
    "lwb_sample": (_i, [_vp, _vp, _u64, _dbl, _vp, _u64, _vp, _pdbl]),
    "lwb_sample_dev": (_i, [_vp, _vp, _u64, _dbl, _vp, _u64, _vp, _vp]),
    "lwb_basis_sample": (_i, [_vp, _i, _vp, _i, _vp, _dbl, _vp, _u64, _vp, _pdbl]),
    "lwb_basis_sample_selected": (_i, [_vp, _i, _vp, _i, _vp, _dbl, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _u64,
                                       _vp, _pdbl]),
    "lwb_keyspace_size": (_i, [_i, _i, _pu64]),
    "lwb_keyspace_index": (_i, [_i, _vp, _pu64]),
    "lwb_keyspace_state": (_i, [_i, _i, _u64, _vp]),
    "lwb_keyspace_states": (_i, [_i, _i, _vp, _u64, _vp]),
    "lwb_slos_states": (_i, [_i, _i, _vp, _u64, _vp]),
    "lwb_dist_create": (_i, [_vp, _i, _i, _vp, _pi]),
    "lwb_dist_from_circuit": (_i, [_vp, _vp, _i, _i, _vp, _dbl, _i, _i, _pi]),
    "lwb_dist_convolve": (_i, [_vp, _i, _i, _pi]),
    "lwb_dist_axpy": (_i, [_vp, _i, _dbl, _i]),
    "lwb_dist_axpy_many": (_i, [_vp, _i, _i, _vp, _vp]),
    "lwb_dist_info": (_i, [_vp, _i, _pi, _pi, _pu64, C.POINTER(_vp)]),
    "lwb_dist_download": (_i, [_vp, _i, _vp]),
    "lwb_dist_sample": (_i, [_vp, _i, _vp, _u64, _vp, _pdbl]),
    "lwb_dist_free": (_i, [_vp, _i]),
    "lwb_source_statistics": (_i, [_i, _vp, _dbl, _dbl, _dbl, _dbl, _u64, _vp, _vp, _vp, _pu64, _pi]),
    "lwb_comm_unique_id": (_i, [_vp]),
    "lwb_comm_init_rank": (_i, [_vp, _i, _i, _vp, C.POINTER(_vp)]),
    "lwb_comm_init_local": (_i, [_i, C.POINTER(_vp)]),
    "lwb_comm_destroy": (_i, [_vp]),
    "lwb_comm_info": (_i, [_vp, _pi, _pi, _pi, _pi]),
    "lwb_comm_set_option": (_i, [_vp, C.c_char_p, _i]),
    "lwb_comm_handle": (_i, [_vp, _i, C.POINTER(_vp)]),
    "lwb_shard_cuts": (_i, [_i, _i, _i, _vp, _vp]),
    "lwb_comm_shard": (_i, [_vp, _i, _i, _i, _pu64, _pu64]),
    "lwb_comm_perm_basis_probs": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _pdbl]),
    "lwb_comm_perm_basis_probs_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, C.POINTER(_vp), C.POINTER(_vp)]),
    "lwb_comm_pdist": (_i, [_vp, _vp, _i, _i, _vp, _dbl, _i, _vp, _vp, _u64, _pu64]),
    "lwb_comm_perm_single": (_i, [_vp, _vp, _i, _i, _vp]),
    "lwb_comm_perm_single_dev": (_i, [_vp, _vp, _i, _i, _vp]),
    "lwb_dists_from_circuit": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _dbl, _i, _i, _vp]),
    "lwb_debug_pending_cuda_error": (_i, [C.c_char_p, _i]),
    "lwb_fp64_peak": (_i, [_vp, _pdbl]),
    "lwb_fp64_probe": (_i, [_vp, _i, _pdbl]),
}

_cdll = None
_lock = threading.Lock()


def cdll() -> C.CDLL:
    """Load liblwb200.so (once).  Raises BackendError if it has not been built."""
    global _cdll
    with _lock:
        if _cdll is None:
            if not os.path.exists(LIB_PATH):
                raise _exc["BackendError"](
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(make -C lightworks_b200/csrc).  lightworks_b200 has no CPU fallback."
                )
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _cdll = lib
    return _cdll


def _raise(rc: int):
    msg = cdll().lwb_last_error().decode(errors="replace")
    if rc == LWB_E_PHOTONS:
        raise _exc["PhotonNumberError"](msg)
    if rc in (LWB_E_ARG, LWB_E_LIMIT):
        raise ValueError(msg)
    raise _exc["BackendError"](msg)


def check(rc: int):
    if rc != 0:
        _raise(rc)


def ptr(a) -> C.c_void_p:
    """Raw pointer of a numpy array, torch tensor (host or device) or int address."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())  # torch tensor


# ---- host-only helpers (no GPU needed) ------------------------------------------------------------
def fock_count(m: int, n: int) -> int:
    out = C.c_uint64(0)
    check(cdll().lwb_fock_count(m, n, C.byref(out)))
    return int(out.value)


def _u8(state) -> np.ndarray:
    return np.ascontiguousarray(state, dtype=np.uint8)


def fock_rank(state) -> int:
    s = _u8(state)
    out = C.c_uint64(0)
    check(cdll().lwb_fock_rank(len(s), ptr(s), C.byref(out)))
    return int(out.value)


def fock_unrank(m: int, n: int, rank: int) -> np.ndarray:
    s = np.zeros(m, dtype=np.uint8)
    check(cdll().lwb_fock_unrank(m, n, rank, ptr(s)))
    return s


def slos_rank(state) -> int:
    s = _u8(state)
    out = C.c_uint64(0)
    check(cdll().lwb_slos_rank(len(s), ptr(s), C.byref(out)))
    return int(out.value)


def slos_unrank(m: int, n: int, rank: int) -> np.ndarray:
    s = np.zeros(m, dtype=np.uint8)
    check(cdll().lwb_slos_unrank(m, n, rank, ptr(s)))
    return s


def slos_plan_check(m: int, k: int) -> dict:
    """Host-only self-check of the SLOS prefix-unit plan (csrc/slos_units.cu) of layer (m modes, k photons)."""
    u, i, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    check(cdll().lwb_slos_plan_check(m, k, C.byref(u), C.byref(i), C.byref(c)))
    return {"units": int(u.value), "items": int(i.value), "children": int(c.value)}


def keyspace_size(n_modes: int, kmax: int) -> int:
    """Number of occupation vectors of n_modes modes with 0..kmax photons (the dense distribution layout)."""
    out = C.c_uint64(0)
    check(cdll().lwb_keyspace_size(n_modes, kmax, C.byref(out)))
    return int(out.value)


def keyspace_index(state) -> int:
    s = _u8(state)
    out = C.c_uint64(0)
    check(cdll().lwb_keyspace_index(len(s), ptr(s), C.byref(out)))
    return int(out.value)


def keyspace_state(n_modes: int, kmax: int, index: int) -> np.ndarray:
    s = np.zeros(n_modes, dtype=np.uint8)
    check(cdll().lwb_keyspace_state(n_modes, kmax, index, ptr(s)))
    return s


def keyspace_states(n_modes: int, kmax: int, indices) -> np.ndarray:
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    out = np.zeros((len(idx), n_modes), dtype=np.uint8)
    check(cdll().lwb_keyspace_states(n_modes, kmax, ptr(idx), len(idx), ptr(out)))
    return out


def slos_states(m: int, n: int, ranks) -> np.ndarray:
    """Occupation vectors of the states at the given positions of the SLOS (dict insertion) order of (m, n)."""
    idx = np.ascontiguousarray(ranks, dtype=np.uint64)
    out = np.zeros((len(idx), m), dtype=np.uint8)
    check(cdll().lwb_slos_states(m, n, ptr(idx), len(idx), ptr(out)))
    return out


def source_classes(state, purity=1.0, brightness=1.0, indistinguishability=1.0, probability_threshold=1e-9):
    """lwb_source_statistics as arrays: (annotated, group_occ uint8[K, m], single_occ uint8[K, m], probs float64[K]) in
    the reference's dict order (include/lwb200.h): group_occ = photons per mode sharing label 0, single_occ = photons
    per mode carrying a label of their own."""
    s = _u8(state)
    m = len(s)
    cap = 1 << 16  # a too small capacity costs a second enumeration
    while True:
        g = np.zeros((cap, m), dtype=np.uint8)
        b = np.zeros((cap, m), dtype=np.uint8)
        p = np.zeros(cap, dtype=np.float64)
        cnt, ann = C.c_uint64(0), C.c_int(0)
        rc = cdll().lwb_source_statistics(m, ptr(s), float(purity), float(brightness), float(indistinguishability),
                                          float(probability_threshold), cap, ptr(g), ptr(b), ptr(p), C.byref(cnt),
                                          C.byref(ann))
        if rc == LWB_E_LIMIT and cnt.value > cap:
            cap = int(cnt.value)
            continue
        check(rc)
        break
    k = int(cnt.value)
    return bool(ann.value), g[:k], b[:k], p[:k]


def source_statistics(state, purity=1.0, brightness=1.0, indistinguishability=1.0, probability_threshold=1e-9):
    """Source(purity, brightness, indistinguishability, probability_threshold)._build_statistics(state)
    (emulator/components/source.py:134-161) enumerated natively on the host.  Returns (annotated, entries) with
    entries = [(state, p), ...] in the reference's dict order; a state is a tuple of per-mode label tuples
    (AnnotatedState.s, labels renumbered by first appearance as _remap_distribution does) when `annotated`,
    else a plain occupation tuple (State.s)."""
    ann, g, b, p = source_classes(state, purity, brightness, indistinguishability, probability_threshold)
    m = g.shape[1]
    if not ann:
        return False, [(tuple(int(v) for v in row), float(q)) for row, q in zip(g, p)]
    out = []
    for gi, bi, q in zip(g.tolist(), b.tolist(), p.tolist()):
        nxt, zero, modes = 0, None, []
        for j in range(m):
            labels = []
            if gi[j]:  # the shared label gets its number at its first appearance
                if zero is None:
                    zero, nxt = nxt, nxt + 1
                labels += [zero] * gi[j]
            if bi[j]:
                labels += range(nxt, nxt + bi[j])
                nxt += bi[j]
            modes.append(tuple(labels))
        out.append((tuple(modes), q))
    return True, out


class DeviceDistribution:
    """A dense distribution over the key space K(n_modes, kmax) resident on the GPU (one slot of a Context).
    `a @ b` is the merge-convolution of probability_distribution.py:150-158; freed with the object."""

    def __init__(self, ctx: "Context", slot: int):
        self.ctx, self.slot = ctx, slot
        nm, km, sz = C.c_int(0), C.c_int(0), C.c_uint64(0)
        check(cdll().lwb_dist_info(ctx.handle, slot, C.byref(nm), C.byref(km), C.byref(sz), None))
        self.n_modes, self.kmax, self.size = nm.value, km.value, int(sz.value)

    def __matmul__(self, other: "DeviceDistribution") -> "DeviceDistribution":
        out = C.c_int(-1)
        check(cdll().lwb_dist_convolve(self.ctx.handle, self.slot, other.slot, C.byref(out)))
        return DeviceDistribution(self.ctx, out.value)

    def add_scaled(self, alpha: float, x: "DeviceDistribution") -> None:
        check(cdll().lwb_dist_axpy(self.ctx.handle, self.slot, float(alpha), x.slot))

    def add_scaled_many(self, alphas, xs) -> None:
        """self += alphas[j] * xs[j] for every j, in that order, as ONE launch (lwb_dist_axpy_many)."""
        if not len(xs):
            return
        a = np.ascontiguousarray(alphas, dtype=np.float64)
        slots = np.ascontiguousarray([x.slot for x in xs], dtype=np.int32)
        if len(a) != len(slots):
            raise ValueError("one coefficient per distribution")
        check(cdll().lwb_dist_axpy_many(self.ctx.handle, self.slot, len(slots), ptr(a), ptr(slots)))

    def to_numpy(self) -> np.ndarray:
        out = np.zeros(self.size, dtype=np.float64)
        check(cdll().lwb_dist_download(self.ctx.handle, self.slot, ptr(out)))
        return out

    def sample(self, uniform):
        uniform = np.ascontiguousarray(uniform, dtype=np.float64)
        idx = np.zeros(len(uniform), dtype=np.uint64)
        total = C.c_double(0)
        check(cdll().lwb_dist_sample(self.ctx.handle, self.slot, ptr(uniform), len(uniform), ptr(idx), C.byref(total)))
        return idx, float(total.value)

    def device_ptr(self) -> int:
        p = C.c_void_p(0)
        check(cdll().lwb_dist_info(self.ctx.handle, self.slot, None, None, None, C.byref(p)))
        return int(p.value)

    def free(self):
        if self.slot >= 0 and self.ctx.handle:
            cdll().lwb_dist_free(self.ctx.handle, self.slot)
        self.slot = -1

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- device context ---------------------------------------------------------------------------------
class Context:
    """One lwb_handle (one GPU).  Thin, typed wrappers over the C ABI with numpy host buffers."""

    def __init__(self, device: int = 0, _borrowed=None):
        self._owned = _borrowed is None
        if _borrowed is not None:  # a handle owned by a communicator (lwb_comm_handle)
            self._h = _borrowed
        else:
            self._h = C.c_void_p(0)
            check(cdll().lwb_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h and self._owned:
            cdll().lwb_destroy(self._h)
        self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        """Enqueue on a caller-owned stream (raw cudaStream_t; 0 = legacy default stream);
        None restores the handle's own stream."""
        if cuda_stream is None:
            check(cdll().lwb_set_stream(self._h, C.c_void_p(0), 1))
        else:
            check(cdll().lwb_set_stream(self._h, C.c_void_p(cuda_stream), 0))

    def synchronize(self):
        check(cdll().lwb_synchronize(self._h))

    def device_info(self) -> dict:
        sm, khz = C.c_int(0), C.c_int(0)
        name = C.create_string_buffer(128)
        check(cdll().lwb_device_info(self._h, C.byref(sm), C.byref(khz), name, 128))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "name": name.value.decode()}

    def launch_count(self) -> int:
        out = C.c_uint64(0)
        check(cdll().lwb_launch_count(self._h, C.byref(out)))
        return int(out.value)

    def fp64_peak(self) -> float:
        out = C.c_double(0)
        check(cdll().lwb_fp64_peak(self._h, C.byref(out)))
        return float(out.value)

    def fp64_probe(self, mode: int) -> float:
        out = C.c_double(0)
        check(cdll().lwb_fp64_probe(self._h, mode, C.byref(out)))
        return float(out.value)

    # -- host-buffer entry points --
    def fock_basis(self, m: int, n: int, r0: int = 0, cnt: int | None = None) -> np.ndarray:
        cnt = fock_count(m, n) - r0 if cnt is None else cnt
        out = np.zeros((cnt, m), dtype=np.uint8)
        check(cdll().lwb_fock_basis(self._h, m, n, r0, cnt, ptr(out)))
        return out

    def perm_matrices(self, A, dtype: int = LWB_C128) -> np.ndarray:  # noqa: N803
        A = np.ascontiguousarray(A, dtype=np.complex128)  # noqa: N806
        if A.ndim != 3 or A.shape[1] != A.shape[2]:
            raise ValueError("expected an array of shape (B, n, n)")
        out = np.zeros(A.shape[0], dtype=np.complex128)
        check(cdll().lwb_perm_matrices(self._h, ptr(A), A.shape[0], A.shape[1], dtype, ptr(out)))
        return out

    def perm_amplitudes(self, U, ins, outs, dtype: int = LWB_C128, probs: bool = False) -> np.ndarray:  # noqa: N803
        U = np.ascontiguousarray(U, dtype=np.complex128)  # noqa: N806
        ins = np.ascontiguousarray(np.atleast_2d(ins), dtype=np.uint8)
        outs = np.ascontiguousarray(np.atleast_2d(outs), dtype=np.uint8)
        m = U.shape[0]
        if U.ndim != 2 or U.shape[1] != m or ins.shape[1] != m or outs.shape[1] != m:
            raise ValueError("unitary / state dimensions do not match")
        out = np.zeros((ins.shape[0], outs.shape[0]), dtype=np.float64 if probs else np.complex128)
        check(cdll().lwb_perm_amplitudes(self._h, ptr(U), m, ptr(ins), ins.shape[0], ptr(outs), outs.shape[0],
                                         dtype, int(probs), ptr(out)))
        return out

    def perm_amplitudes_batch(self, U, ins, outs, dtype: int = LWB_C128, probs: bool = False) -> np.ndarray:  # noqa: N803
        """lwb_perm_amplitudes_batch: U [B, m, m], ins [B, n_in, m] (or [n_in, m] shared by every task), outs [n_out, m]
        -> [B, n_in, n_out] amplitudes (or probabilities), one launch for the whole batch."""
        U = np.ascontiguousarray(U, dtype=np.complex128)  # noqa: N806
        ins = np.ascontiguousarray(ins, dtype=np.uint8)
        outs = np.ascontiguousarray(np.atleast_2d(outs), dtype=np.uint8)
        if U.ndim != 3 or U.shape[1] != U.shape[2]:
            raise ValueError("expected unitaries of shape (B, m, m)")
        b, m = U.shape[0], U.shape[1]
        shared = ins.ndim == 2
        if ins.shape[-1] != m or outs.shape[1] != m or (not shared and (ins.ndim != 3 or ins.shape[0] != b)):
            raise ValueError("unitary / state dimensions do not match")
        n_in = ins.shape[-2]
        out = np.zeros((b, n_in, outs.shape[0]), dtype=np.float64 if probs else np.complex128)
        check(cdll().lwb_perm_amplitudes_batch(self._h, ptr(U), b, m, ptr(ins), n_in, int(shared), ptr(outs), outs.shape[0],
                                               dtype, int(probs), ptr(out)))
        return out

    def perm_basis_probs(self, U, in_state, r0: int = 0, cnt: int | None = None, dtype: int = LWB_C128,  # noqa: N803
                         out: np.ndarray | None = None) -> np.ndarray:
        U = np.ascontiguousarray(U, dtype=np.complex128)  # noqa: N806
        s = _u8(in_state)
        m = U.shape[0]
        if len(s) != m:
            raise ValueError("unitary / state dimensions do not match")
        cnt = fock_count(m, int(s.sum())) - r0 if cnt is None else cnt
        if out is None:
            out = np.zeros(cnt, dtype=np.float64)
        check(cdll().lwb_perm_basis_probs(self._h, ptr(U), m, ptr(s), r0, cnt, dtype, ptr(out)))
        return out

    def perm_single(self, A, dtype: int = LWB_C128, slice_: int = 0, n_slices: int = 1) -> complex:  # noqa: N803
        A = np.ascontiguousarray(A, dtype=np.complex128)  # noqa: N806
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("expected a square matrix")
        out = np.zeros(2)
        check(cdll().lwb_perm_single(self._h, ptr(A), A.shape[0], dtype, slice_, n_slices, ptr(out)))
        return complex(out[0], out[1])


BACKEND_PERMANENT, BACKEND_SLOS = 0, 1
ORDER_FOCK, ORDER_SLOS = 0, 1


def _slos_methods():
    def _prep(U, in_state):  # noqa: N803
        U = np.ascontiguousarray(U, dtype=np.complex128)  # noqa: N806
        s = _u8(in_state)
        if U.ndim != 2 or U.shape[0] != U.shape[1] or len(s) != U.shape[0]:
            raise ValueError("unitary / state dimensions do not match")
        return U, s, fock_count(U.shape[0], int(s.sum()))

    def slos_amplitudes(self, U, in_state, order=ORDER_SLOS):  # noqa: N803
        """SLOSBackend.calculate (slos.py:82-105) as an array: amplitudes of every output state."""
        U, s, S = _prep(U, in_state)  # noqa: N806
        out = np.zeros(S, dtype=np.complex128)
        check(cdll().lwb_slos_amplitudes(self._h, ptr(U), U.shape[0], ptr(s), order, ptr(out)))
        return out

    def slos_basis_probs(self, U, in_state, order=ORDER_SLOS, out=None):  # noqa: N803
        U, s, S = _prep(U, in_state)  # noqa: N806
        if out is None:
            out = np.zeros(S, dtype=np.float64)
        check(cdll().lwb_slos_basis_probs(self._h, ptr(U), U.shape[0], ptr(s), order, ptr(out)))
        return out

    def slos_basis_probs_dev(self, d_U, m, in_state, d_probs, order=ORDER_SLOS):  # noqa: N803
        s = _u8(in_state)
        check(cdll().lwb_slos_basis_probs_dev(self._h, ptr(d_U), m, ptr(s), order, ptr(d_probs)))

    def slos_amplitudes_dev(self, d_U, m, in_state, d_amps, order=ORDER_SLOS):  # noqa: N803
        s = _u8(in_state)
        check(cdll().lwb_slos_amplitudes_dev(self._h, ptr(d_U), m, ptr(s), order, ptr(d_amps)))

    def pdist(self, U_full, n_modes, in_state, threshold=1e-9, backend=BACKEND_PERMANENT, dtype=LWB_C128):  # noqa: N803
        """full_probability_distribution (permanent.py:116-163 / slos.py:43-80) as parallel arrays in the
        reference's dict insertion order: (keys uint8[K, n_modes], vals float64[K])."""
        from math import comb

        U_full = np.ascontiguousarray(U_full, dtype=np.complex128)  # noqa: N806
        s = _u8(in_state)
        m_tot = U_full.shape[0]
        if U_full.ndim != 2 or U_full.shape[1] != m_tot or len(s) != n_modes or n_modes > m_tot:
            raise ValueError("unitary / state dimensions do not match")
        n = int(s.sum())
        if m_tot == n_modes:
            cap = comb(m_tot + n - 1, n)
        else:
            cap = sum(comb(n_modes + k - 1, k) for k in range(n + 1)) + 1
        keys = np.zeros((cap, n_modes), dtype=np.uint8)
        vals = np.zeros(cap, dtype=np.float64)
        count = C.c_uint64(0)
        check(cdll().lwb_pdist(self._h, ptr(U_full), m_tot, n_modes, ptr(s), float(threshold), backend, dtype,
                               ptr(keys), ptr(vals), cap, C.byref(count)))
        k = int(count.value)
        return keys[:k], vals[:k]

    def sample(self, probs, uniform, threshold=-1.0):
        """numpy's Generator.choice(len(probs), p=probs / probs.sum(), size=len(uniform)) for caller-drawn uniform
        variates (sampler.py:329-349): index drawn by every variate, and the sum of the kept probabilities."""
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        uniform = np.ascontiguousarray(uniform, dtype=np.float64)
        idx = np.zeros(len(uniform), dtype=np.uint64)
        total = C.c_double(0)
        check(cdll().lwb_sample(self._h, ptr(probs), len(probs), float(threshold), ptr(uniform), len(uniform), ptr(idx),
                                C.byref(total)))
        return idx, float(total.value)

    def sample_dev(self, d_probs, S, d_uniform, n_samples, d_index, threshold=-1.0, d_total=None):  # noqa: N803
        check(cdll().lwb_sample_dev(self._h, ptr(d_probs), S, float(threshold), ptr(d_uniform), n_samples, ptr(d_index),
                                    ptr(d_total)))

    def basis_sample(self, U, in_state, uniform, threshold=1e-9, backend=BACKEND_PERMANENT):  # noqa: N803
        """Full-basis distribution of one input state computed AND sampled on the device: returns the drawn
        positions (fock_basis ranks for the permanent backend, SLOS dict positions for slos) and sum(p)."""
        U, s, _ = _prep(U, in_state)  # noqa: N806
        uniform = np.ascontiguousarray(uniform, dtype=np.float64)
        idx = np.zeros(len(uniform), dtype=np.uint64)
        total = C.c_double(0)
        check(cdll().lwb_basis_sample(self._h, backend, ptr(U), U.shape[0], ptr(s), float(threshold), ptr(uniform),
                                      len(uniform), ptr(idx), C.byref(total)))
        return idx, float(total.value)

    def basis_sample_selected(self, U, in_state, uniform, threshold=1e-9, backend=BACKEND_PERMANENT,  # noqa: N803
                              threshold_detect=False, min_detection=0, heralds=None, rules=None):
        """basis_sample restricted to the output states SamplerRunner._sample_N_outputs keeps (sampler.py:275-327):
        heralds = {mode: photons}, rules = [(modes, allowed photon numbers)] on FULL-circuit modes; the selection is
        a device mask over the whole basis.  Returns (drawn positions, probability mass of the kept states)."""
        U, s, _ = _prep(U, in_state)  # noqa: N806
        uniform = np.ascontiguousarray(uniform, dtype=np.float64)
        idx = np.zeros(len(uniform), dtype=np.uint64)
        total = C.c_double(0)
        heralds = dict(heralds or {})
        hm = np.array(list(heralds.keys()), dtype=np.int32)
        hc = np.array(list(heralds.values()), dtype=np.int32)
        rules = list(rules or [])
        rm = np.array([m for modes, _ in rules for m in modes], dtype=np.int32)
        rmo = np.cumsum([0] + [len(modes) for modes, _ in rules]).astype(np.int32)
        rc = np.array([c for _, counts in rules for c in counts], dtype=np.int32)
        rco = np.cumsum([0] + [len(counts) for _, counts in rules]).astype(np.int32)
        check(cdll().lwb_basis_sample_selected(self._h, backend, ptr(U), U.shape[0], ptr(s), float(threshold),
                                               int(bool(threshold_detect)), int(min_detection), ptr(hm), ptr(hc), len(hm),
                                               ptr(rm), ptr(rmo), ptr(rc), ptr(rco), len(rules), ptr(uniform),
                                               len(uniform), ptr(idx), C.byref(total)))
        return idx, float(total.value)

    def dist_zeros(self, n_modes, kmax):
        out = C.c_int(-1)
        check(cdll().lwb_dist_create(self._h, n_modes, kmax, None, C.byref(out)))
        return DeviceDistribution(self, out.value)

    def dist_from_dense(self, n_modes, kmax, dense):
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        if len(dense) != keyspace_size(n_modes, kmax):
            raise ValueError("dense vector does not match the key space")
        out = C.c_int(-1)
        check(cdll().lwb_dist_create(self._h, n_modes, kmax, ptr(dense), C.byref(out)))
        return DeviceDistribution(self, out.value)

    def dists_from_circuit(self, U_full, n_modes, in_states, threshold=1e-9, backend=BACKEND_PERMANENT, dtype=LWB_C128,  # noqa: N803
                           comm=None):
        """lwb_dists_from_circuit: full_probability_distribution of MANY input states of one circuit
        (probability_distribution.py:127-133) into device-resident dense distributions of this context, one call and
        one synchronisation; with a LOCAL communicator the inputs are computed round-robin on its GPUs."""
        U_full = np.ascontiguousarray(U_full, dtype=np.complex128)  # noqa: N806
        ins = np.ascontiguousarray(np.atleast_2d(in_states), dtype=np.uint8)
        m_tot = U_full.shape[0]
        if U_full.ndim != 2 or U_full.shape[1] != m_tot or ins.shape[1] != n_modes or n_modes > m_tot:
            raise ValueError("unitary / state dimensions do not match")
        slots = np.full(len(ins), -1, dtype=np.int32)
        check(cdll().lwb_dists_from_circuit(self._h, comm.handle if comm is not None else None, ptr(U_full), m_tot, n_modes,
                                            ptr(ins), len(ins), float(threshold), backend, dtype, ptr(slots)))
        return [DeviceDistribution(self, int(sl)) for sl in slots]

    def dist_from_circuit(self, U_full, n_modes, in_state, threshold=1e-9, backend=BACKEND_PERMANENT, dtype=LWB_C128):  # noqa: N803
        """full_probability_distribution (permanent.py:116-163 / slos.py:43-80) into a device-resident dense
        distribution over K(n_modes, photons of in_state)."""
        U_full = np.ascontiguousarray(U_full, dtype=np.complex128)  # noqa: N806
        s = _u8(in_state)
        m_tot = U_full.shape[0]
        if U_full.ndim != 2 or U_full.shape[1] != m_tot or len(s) != n_modes or n_modes > m_tot:
            raise ValueError("unitary / state dimensions do not match")
        out = C.c_int(-1)
        check(cdll().lwb_dist_from_circuit(self._h, ptr(U_full), m_tot, n_modes, ptr(s), float(threshold), backend,
                                           dtype, C.byref(out)))
        return DeviceDistribution(self, out.value)

    for f in (slos_amplitudes, slos_basis_probs, slos_basis_probs_dev, slos_amplitudes_dev, pdist, sample, sample_dev,
              basis_sample, basis_sample_selected, dist_zeros, dist_from_dense, dist_from_circuit, dists_from_circuit):
        setattr(Context, f.__name__, f)


_slos_methods()

_default_ctx: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if "LWB_DEVICE" not in os.environ else int(os.environ["LWB_DEVICE"])
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---- device-pointer entry points (torch tensors or raw addresses; async on the handle's stream) ----
def _dev_methods():
    def perm_matrices_dev(self, d_A, B, n, d_out, dtype=LWB_C128):  # noqa: N803
        check(cdll().lwb_perm_matrices_dev(self._h, ptr(d_A), B, n, dtype, ptr(d_out)))

    def perm_amplitudes_dev(self, d_U, m, d_ins, n_in, d_outs, n_out, n_photons, d_out, dtype=LWB_C128, probs=False):  # noqa: N803
        check(cdll().lwb_perm_amplitudes_dev(self._h, ptr(d_U), m, ptr(d_ins), n_in, ptr(d_outs), n_out, n_photons,
                                             dtype, int(probs), ptr(d_out)))

    def perm_basis_probs_dev(self, d_U, m, d_in_state, n_photons, r0, cnt, d_probs, dtype=LWB_C128):  # noqa: N803
        check(cdll().lwb_perm_basis_probs_dev(self._h, ptr(d_U), m, ptr(d_in_state), n_photons, r0, cnt, dtype,
                                              ptr(d_probs)))

    def perm_single_dev(self, d_A, n, d_out2, dtype=LWB_C128, slice_=0, n_slices=1):  # noqa: N803
        check(cdll().lwb_perm_single_dev(self._h, ptr(d_A), n, dtype, slice_, n_slices, ptr(d_out2)))

    def fock_basis_dev(self, m, n, r0, cnt, d_states):
        check(cdll().lwb_fock_basis_dev(self._h, m, n, r0, cnt, ptr(d_states)))

    for f in (perm_matrices_dev, perm_amplitudes_dev, perm_basis_probs_dev, perm_single_dev, fock_basis_dev):
        setattr(Context, f.__name__, f)


_dev_methods()


# ---- multi-GPU communicators (include/lwb200.h "multi-GPU") ---------------------------------------------
def shard_cuts(m: int, n: int, fractions) -> list[int]:
    """lwb_shard_cuts: fock_basis ranks where the cumulative work of the permanent walk reaches the given fractions."""
    fr = np.ascontiguousarray(fractions, dtype=np.float64)
    cuts = np.zeros(len(fr), dtype=np.uint64)
    check(cdll().lwb_shard_cuts(m, n, len(fr), ptr(fr), ptr(cuts)))
    return [int(v) for v in cuts]


class DevicePointer:
    """A library-owned device buffer as a `__cuda_array_interface__` object (torch.as_tensor / cupy.asarray view it
    without copying).  Valid until the communicator reallocates or is destroyed."""

    def __init__(self, address: int, count: int, typestr: str = "<f8"):
        self.address, self.count = address, count
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (address, False), "version": 3,
                                         "strides": None}


class Comm:
    """lwb_comm: the hot path sharded over several GPUs.  `Comm.local(n)` - this process drives n GPUs;
    `Comm.rank(ctx, rank, world, unique_id)` - one process per GPU, NCCL-bootstrapped (unique_id from
    `Comm.unique_id()` on rank 0, broadcast by the host framework)."""

    def __init__(self, handle, ctx=None):
        self._c, self._ctx = handle, ctx
        self._contexts: dict[int, Context] = {}

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(cdll().lwb_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def local(cls, n_gpus: int) -> "Comm":
        h = C.c_void_p(0)
        check(cdll().lwb_comm_init_local(int(n_gpus), C.byref(h)))
        return cls(h)

    @classmethod
    def rank(cls, ctx: Context, rank: int, world: int, unique_id: bytes) -> "Comm":
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of Comm.unique_id()")
        h = C.c_void_p(0)
        check(cdll().lwb_comm_init_rank(ctx.handle, rank, world, C.create_string_buffer(unique_id, 128), C.byref(h)))
        return cls(h, ctx)

    @property
    def handle(self):
        return self._c

    def close(self):
        if self._c:
            for c in self._contexts.values():  # borrowed handles die with the communicator: objects that still hold
                c._h = C.c_void_p(0)           # the Context (device distributions) must see a dead handle, not a dangling one
            self._contexts.clear()
            cdll().lwb_comm_destroy(self._c)
            self._c = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        r, w, k, p = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        check(cdll().lwb_comm_info(self._c, C.byref(r), C.byref(w), C.byref(k), C.byref(p)))
        return {"rank": r.value, "world": w.value, "kind": "rank" if k.value else "local", "peer_stores": bool(p.value)}

    def set_option(self, name: str, value: int):
        check(cdll().lwb_comm_set_option(self._c, name.encode(), int(value)))

    def context(self, index: int = 0) -> Context:
        """The Context of device `index` (LOCAL) / the rank's own context (RANK, index 0); owned by the communicator."""
        if index not in self._contexts:
            h = C.c_void_p(0)
            check(cdll().lwb_comm_handle(self._c, index, C.byref(h)))
            self._contexts[index] = Context(index, _borrowed=h)
        return self._contexts[index]

    def shard(self, m: int, n: int, rank: int) -> tuple[int, int]:
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(cdll().lwb_comm_shard(self._c, m, n, rank, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def perm_basis_probs(self, U, in_state, dtype: int = LWB_C128, gather_host: bool = True, out=None, want_probs=True):  # noqa: N803
        """Whole-basis probabilities sharded over the communicator -> (probs float64[S] | None, sum of probabilities)."""
        U = np.ascontiguousarray(U, dtype=np.complex128)  # noqa: N806
        s = _u8(in_state)
        m = U.shape[0]
        if U.ndim != 2 or U.shape[1] != m or len(s) != m:
            raise ValueError("unitary / state dimensions do not match")
        if want_probs and out is None:
            out = np.zeros(fock_count(m, int(s.sum())), dtype=np.float64)
        total = C.c_double(0)
        check(cdll().lwb_comm_perm_basis_probs(self._c, ptr(U), m, ptr(s), dtype, int(gather_host),
                                               ptr(out) if want_probs else None, C.byref(total)))
        return (out if want_probs else None), float(total.value)

    def perm_basis_probs_dev(self, d_U, m: int, d_in_state, n_photons: int, dtype: int = LWB_C128):  # noqa: N803
        """RANK communicators, device inputs, asynchronous: (DevicePointer of this rank's complete distribution,
        DevicePointer of the all-reduced sum)."""
        p, q = C.c_void_p(0), C.c_void_p(0)
        check(cdll().lwb_comm_perm_basis_probs_dev(self._c, ptr(d_U), m, ptr(d_in_state), n_photons, dtype, C.byref(p),
                                                   C.byref(q)))
        return DevicePointer(int(p.value), fock_count(m, n_photons)), DevicePointer(int(q.value), 1)

    def pdist(self, U_full, n_modes, in_state, threshold=1e-9, dtype=LWB_C128):  # noqa: N803
        """Context.pdist(backend = permanent) with the permanents sharded over the communicator."""
        from math import comb

        U_full = np.ascontiguousarray(U_full, dtype=np.complex128)  # noqa: N806
        s = _u8(in_state)
        m_tot = U_full.shape[0]
        if U_full.ndim != 2 or U_full.shape[1] != m_tot or len(s) != n_modes or n_modes > m_tot:
            raise ValueError("unitary / state dimensions do not match")
        n = int(s.sum())
        cap = comb(m_tot + n - 1, n) if m_tot == n_modes else sum(comb(n_modes + k - 1, k) for k in range(n + 1)) + 1
        keys = np.zeros((cap, n_modes), dtype=np.uint8)
        vals = np.zeros(cap, dtype=np.float64)
        count = C.c_uint64(0)
        check(cdll().lwb_comm_pdist(self._c, ptr(U_full), m_tot, n_modes, ptr(s), float(threshold), dtype, ptr(keys),
                                    ptr(vals), cap, C.byref(count)))
        k = int(count.value)
        return keys[:k], vals[:k]

    def perm_single(self, A, dtype: int = LWB_C128) -> complex:  # noqa: N803
        A = np.ascontiguousarray(A, dtype=np.complex128)  # noqa: N806
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("expected a square matrix")
        out = np.zeros(2)
        check(cdll().lwb_comm_perm_single(self._c, ptr(A), A.shape[0], dtype, ptr(out)))
        return complex(out[0], out[1])

    def perm_single_dev(self, d_A, n: int, d_out2, dtype: int = LWB_C128):  # noqa: N803
        check(cdll().lwb_comm_perm_single_dev(self._c, ptr(d_A), n, dtype, ptr(d_out2)))


def set_option(name: str, value: int):
    check(cdll().lwb_set_option(name.encode(), int(value)))


def multiplicity_plan(n: int) -> list[dict]:
    """Host-only: the walk of the multiplicity-aware permanent for every multiplicity pattern of n photons."""
    out = []
    npat = C.c_int(0)
    t = 0
    while True:
        s_ = (C.c_int * 64)()
        d, T, f, a = C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0)  # noqa: N806
        check(cdll().lwb_multiplicity_plan(n, t, C.byref(npat), s_, C.byref(d), C.byref(T), C.byref(f), C.byref(a)))
        out.append({"s": list(s_[: d.value]), "terms": T.value, "factor": f.value, "abs_coef_sum": a.value})
        t += 1
        if t >= npat.value:
            return out


def set_tuning(n: int, log_l: int = -1):
    check(cdll().lwb_set_tuning(n, log_l))
