"""ctypes binding of the C ABI in include/mbgdml_b200.h.

The shared library is built in-tree (``mbgdml_b200/libmbgdml_b200.so``) by
``python -m mbgdml_b200.build`` / ``__graft_entry__.build()``.  There is no CPU
fallback: if the library is missing, or no B200 is visible, the product path
raises :class:`NativeUnavailable`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

LIB_NAME = "libmbgdml_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

MBG_OK, MBG_ERR_ARG, MBG_ERR_CUDA, MBG_ERR_UNSUPPORTED, MBG_ERR_NOMEM = 0, -1, -2, -3, -4
CRIT_NONE, CRIT_COM_DISTANCE_SUM, CRIT_MAX_ATOM_PAIR_DIST = 0, 1, 2
BOUND_UPPER, BOUND_LOWER, BOUND_RANGE = 0, 1, 2
FLAG_R_ON_DEVICE, FLAG_OUT_ON_DEVICE, FLAG_VIRIAL, FLAG_ACCUMULATE = 1, 2, 4, 8
CONTRACT_AUTO, CONTRACT_FMA, CONTRACT_DMMA, CONTRACT_DMMA_WAVE = 0, 1, 2, 3
CULL_AUTO, CULL_FULL, CULL_PAIRLIST = 0, 1, 2

# every symbol include/mbgdml_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "mbg_last_error", "mbg_abi_version", "mbg_device_count", "mbg_context_create", "mbg_context_destroy",
    "mbg_context_set_stream", "mbg_context_set_contraction", "mbg_context_set_culling", "mbg_context_synchronize", "mbg_context_launch_count", "mbg_context_last_timing",
    "mbg_context_last_launches",
    "mbg_model_create", "mbg_model_destroy", "mbg_predict", "mbg_predict_decomp", "mbg_step_create", "mbg_step_run",
    "mbg_step_run_device", "mbg_step_stats", "mbg_step_destroy", "mbg_accept_mask", "mbg_enumerate",
    "mbg_from_r", "mbg_criteria_eval", "mbg_r_mic", "mbg_d_mic", "mbg_mbe_accumulate", "mbg_measure_fp64_peak", "mbg_measure_fp64_tensor_peak", "mbg_measure_exp_sqrt_rate",
    "mbg_kernel_mat", "mbg_kernel_mat_cols", "mbg_mat52", "mbg_mat52_rows",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class NativeUnavailable(RuntimeError):
    """The CUDA library (or a B200) is not available; there is no CPU fallback."""


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32), ("order", C.c_int32), ("n_train", C.c_int32), ("n_perms", C.c_int32),
        ("sig", C.c_double), ("c", C.c_double), ("std", C.c_double),
        ("R_desc", _dp), ("R_d_desc_alpha", _dp), ("tril_perms_lin", _lp), ("alphas_E", _dp),
        ("criteria_kind", C.c_int32), ("criteria_bound", C.c_int32),
        ("criteria_lo", C.c_double), ("criteria_hi", C.c_double), ("criteria_entity_ids", _ip), ("lattice", _dp),
    ]


class SystemDesc(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32), ("n_entities", C.c_int32), ("entity_offsets", _ip), ("entity_atoms", _ip),
        ("masses", _dp), ("cell", _dp), ("cell_cutoff", C.c_double), ("frame_cells", _dp), ("frame_cutoffs", _dp),
        ("pbc", _ip),
    ]


class Job(C.Structure):
    _fields_ = [
        ("model", C.c_void_p), ("set_offsets", _ip), ("set_entities", _ip), ("combs", _ip), ("n_combs", C.c_int64),
        ("n_alchemy", C.c_int32), ("alchemy_entity", _ip), ("alchemy_factor", _dp),
    ]


class JobStats(C.Structure):
    _fields_ = [("n_candidates", C.c_int64), ("n_enumerated", C.c_int64), ("n_accepted", C.c_int64)]


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen the in-tree library and declare the prototypes."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeUnavailable(
                f"{LIB_PATH} is missing: build it with `python -m mbgdml_b200.build` "
                "(there is no CPU fallback for the prediction path)"
            )
        lib = C.CDLL(LIB_PATH)
        lib.mbg_last_error.restype = C.c_char_p
        lib.mbg_abi_version.restype = C.c_int
        lib.mbg_device_count.argtypes = [C.POINTER(C.c_int)]
        lib.mbg_context_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.mbg_context_destroy.argtypes = [C.c_void_p]
        lib.mbg_context_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.mbg_context_set_contraction.argtypes = [C.c_void_p, C.c_int]
        lib.mbg_context_set_culling.argtypes = [C.c_void_p, C.c_int]
        lib.mbg_context_synchronize.argtypes = [C.c_void_p]
        lib.mbg_context_launch_count.argtypes = [C.c_void_p, _lp]
        lib.mbg_context_last_timing.argtypes = [C.c_void_p, _dp, _dp, _lp]
        lib.mbg_context_last_launches.argtypes = [C.c_void_p, C.c_int32, _ip, _ip, _lp, _ip, _ip, _dp]
        lib.mbg_model_create.argtypes = [C.c_void_p, C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
        lib.mbg_model_destroy.argtypes = [C.c_void_p]
        lib.mbg_predict.argtypes = [C.c_void_p, C.POINTER(SystemDesc), C.c_void_p, C.c_int32, C.POINTER(Job), C.c_int32,
                                    C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(JobStats)]
        lib.mbg_predict_decomp.argtypes = [C.c_void_p, C.POINTER(SystemDesc), C.c_void_p, C.POINTER(Job), C.c_uint32,
                                           C.c_void_p, C.c_void_p, C.POINTER(JobStats)]
        lib.mbg_step_create.argtypes = [C.c_void_p, C.POINTER(SystemDesc), C.c_int32, C.POINTER(Job), C.c_int32, C.c_uint32,
                                        C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        lib.mbg_step_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mbg_step_stats.argtypes = [C.c_void_p, C.POINTER(JobStats)]
        lib.mbg_step_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(JobStats)]
        lib.mbg_step_destroy.argtypes = [C.c_void_p]
        lib.mbg_accept_mask.argtypes = [C.c_void_p, C.POINTER(SystemDesc), C.c_void_p, C.c_int32, C.POINTER(Job), C.c_uint32,
                                        C.c_void_p, _lp]
        lib.mbg_enumerate.argtypes = [C.c_void_p, C.c_int32, _ip, _ip, _ip, _lp]
        lib.mbg_from_r.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, C.c_int64, _dp, _dp]
        lib.mbg_criteria_eval.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _dp, _ip, _dp, C.c_int64, _dp]
        lib.mbg_r_mic.argtypes = [C.c_void_p, _dp, C.c_double, _dp, C.c_int32, _dp, _ip, _ip]
        lib.mbg_d_mic.argtypes = [C.c_void_p, _dp, C.c_double, _dp, C.c_int32, C.c_int32, _dp, _ip, _ip]
        lib.mbg_mbe_accumulate.argtypes = [C.c_void_p, C.c_int64, _lp, _lp, _lp, _ip, C.c_int32, C.c_int32, _ip, _ip,
                                           C.c_int32, _ip, _ip, _dp, _dp, C.c_int64, C.c_double, _dp, _dp, C.c_int64]
        lib.mbg_kernel_mat.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _lp, C.c_double, C.c_int32,
                                       C.c_uint32, C.c_void_p]
        lib.mbg_kernel_mat_cols.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _lp, C.c_double, _ip,
                                            C.c_int32, C.c_uint32, C.c_void_p]
        lib.mbg_mat52.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p]
        lib.mbg_mat52_rows.argtypes = [C.c_void_p, C.c_int32, _dp, C.c_int64, C.c_double, _dp, C.c_int64, _dp]
        lib.mbg_measure_fp64_peak.argtypes = [C.c_void_p, _dp, _dp]
        lib.mbg_measure_fp64_tensor_peak.argtypes = [C.c_void_p, _dp]
        lib.mbg_measure_exp_sqrt_rate.argtypes = [C.c_void_p, _dp]
        for name in EXPORTED_SYMBOLS:
            fn = getattr(lib, name)
            if name != "mbg_last_error":
                fn.restype = C.c_int
        _lib = lib
        return lib


def check(rc):
    """Map a C-ABI return code onto the exception types of the reference's Python surface."""
    if rc == MBG_OK:
        return
    msg = load_library().mbg_last_error().decode("utf-8", "replace")
    if rc == MBG_ERR_ARG:
        raise ValueError(msg)
    if rc == MBG_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == MBG_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


class Context:
    """One native context = one GPU + one stream + scratch buffers."""

    def __init__(self, device=0):
        self.lib = load_library()
        n = C.c_int(0)
        rc = self.lib.mbg_device_count(C.byref(n))
        if rc != MBG_OK or n.value == 0:
            raise NativeUnavailable("no CUDA device visible: the mbgdml_b200 prediction path needs a B200 "
                                    "(there is no CPU fallback)")
        self.device = int(device)
        h = C.c_void_p()
        check(self.lib.mbg_context_create(self.device, C.byref(h)))
        self.handle = h
        self._models = {}
        self._stream = 0  # 0 = the context's own stream

    def close(self):
        if getattr(self, "handle", None):
            for m in list(self._models.values()):
                self.lib.mbg_model_destroy(m)
            self._models.clear()
            self.lib.mbg_context_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """Run on an existing CUDA stream handle (0 / None: the context's own stream; 1: cudaStreamLegacy)."""
        check(self.lib.mbg_context_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))
        self._stream = int(cuda_stream or 0)

    def swap_stream(self, cuda_stream):
        """set_stream, returning the previous handle (so that a caller can restore it)."""
        prev = self._stream
        self.set_stream(cuda_stream)
        return prev

    def set_contraction(self, kind):
        """'auto' | 'fma' | 'dmma' | 'dmma_wave': which kernel evaluates the Matern contraction."""
        code = {"auto": CONTRACT_AUTO, "fma": CONTRACT_FMA, "dmma": CONTRACT_DMMA, "dmma_wave": CONTRACT_DMMA_WAVE}[kind] if isinstance(kind, str) else kind
        check(self.lib.mbg_context_set_contraction(self.handle, int(code)))

    def set_culling(self, kind):
        """'auto' | 'full' | 'pairlist': how the candidates of homogeneous 2-/3-body jobs are generated (same accepted
        fragments either way; the pair list scales with the accepted count instead of C(n, order))."""
        code = {"auto": CULL_AUTO, "full": CULL_FULL, "pairlist": CULL_PAIRLIST}[kind] if isinstance(kind, str) else kind
        check(self.lib.mbg_context_set_culling(self.handle, int(code)))

    def synchronize(self):
        check(self.lib.mbg_context_synchronize(self.handle))

    def launch_count(self):
        n = C.c_int64(0)
        check(self.lib.mbg_context_launch_count(self.handle, C.byref(n)))
        return n.value

    def last_timing(self):
        a, b, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        check(self.lib.mbg_context_last_timing(self.handle, C.byref(a), C.byref(b), C.byref(n)))
        return {"ms_contract": a.value, "ms_other": b.value, "n_contract_launches": n.value}

    def last_launches(self, max_records=4096):
        """Per-launch records of the contraction kernel in the last call (list of dicts)."""
        n = C.c_int32(0)
        na = np.zeros(max_records, dtype=np.int32)
        nf = np.zeros(max_records, dtype=np.int64)
        nt = np.zeros(max_records, dtype=np.int32)
        npm = np.zeros(max_records, dtype=np.int32)
        ms = np.zeros(max_records)
        check(self.lib.mbg_context_last_launches(self.handle, max_records, C.byref(n), ptr(na, _ip), ptr(nf, _lp),
                                                 ptr(nt, _ip), ptr(npm, _ip), ptr(ms, _dp)))
        k = min(n.value, max_records)
        return [dict(n_frag_atoms=int(na[i]), n_fragments=int(nf[i]), n_train_padded=int(nt[i]), n_perms=int(npm[i]),
                     ms=float(ms[i])) for i in range(k)]

    def measure_fp64_peak(self):
        t, f = C.c_double(0), C.c_double(0)
        check(self.lib.mbg_measure_fp64_peak(self.handle, C.byref(t), C.byref(f)))
        return t.value, f.value

    def measure_exp_sqrt_rate(self):
        t = C.c_double(0)
        check(self.lib.mbg_measure_exp_sqrt_rate(self.handle, C.byref(t)))
        return t.value

    def measure_fp64_tensor_peak(self):
        t = C.c_double(0)
        check(self.lib.mbg_measure_fp64_tensor_peak(self.handle, C.byref(t)))
        return t.value

    # -- models ---------------------------------------------------------------------
    def model_handle(self, key, build_desc):
        """Device model cached by ``key``; ``build_desc()`` returns (ModelDesc, keepalive)."""
        h = self._models.get(key)
        if h is None:
            desc, _keep = build_desc()
            h = C.c_void_p()
            check(self.lib.mbg_model_create(self.handle, C.byref(desc), C.byref(h)))
            self._models[key] = h
        return h

    def drop_model(self, key):
        h = self._models.pop(key, None)
        if h is not None:
            self.lib.mbg_model_destroy(h)

    # -- small operators ------------------------------------------------------------
    def enumerate(self, sets):
        order = len(sets)
        offs = np.zeros(order + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(s) for s in sets])
        vals = as_i32(np.concatenate([np.asarray(s, dtype=np.int64).ravel() for s in sets])
                      if offs[-1] else np.zeros(0))
        n = C.c_int64(0)
        check(self.lib.mbg_enumerate(self.handle, order, ptr(offs, _ip), ptr(vals, _ip), _ip(), C.byref(n)))
        combs = np.zeros((n.value, order), dtype=np.int32)
        if n.value:
            check(self.lib.mbg_enumerate(self.handle, order, ptr(offs, _ip), ptr(vals, _ip), ptr(combs, _ip),
                                         C.byref(n)))
        return combs

    def from_r(self, R, lattice=None):
        """_from_r (_gdml/desc.py:203-233) for R (n_R, n_atoms, 3): descriptors (n_R, D) and Jacobians (n_R, D, 3)."""
        R = as_f64(R)
        n_R, n_atoms = R.shape[0], R.shape[1]
        D = n_atoms * (n_atoms - 1) // 2
        x, J = np.empty((n_R, D)), np.empty((n_R, D, 3))
        lat = as_f64(lattice).reshape(9) if lattice is not None else None
        check(self.lib.mbg_from_r(self.handle, n_atoms, ptr(lat, _dp), ptr(R, _dp), n_R, ptr(x, _dp), ptr(J, _dp)))
        return x, J

    def criteria_eval(self, kind, masses, entity_ids, R):
        R = as_f64(R)
        n_R, n_atoms = R.shape[0], R.shape[1]
        out = np.zeros(n_R)
        m = as_f64(masses) if masses is not None else None
        e = as_i32(entity_ids) if entity_ids is not None else None
        check(self.lib.mbg_criteria_eval(self.handle, kind, n_atoms, ptr(m, _dp), ptr(e, _ip), ptr(R, _dp), n_R,
                                         ptr(out, _dp)))
        return out

    def mbe_accumulate(self, ref_rows, pair_offsets, lower_idx, slot_entity, atom_entity, atom_rank, lower_off,
                       lower_atoms, E_lower, Deriv_lower, sign, E, Deriv):
        """In-place E[ref_rows] / Deriv[ref_rows] += sign * (sum of the paired lower-order rows); see the header."""
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        ref_rows, pair_offsets, lower_idx = i64(ref_rows), i64(pair_offsets), i64(lower_idx)
        slot_entity = as_i32(slot_entity).reshape(len(lower_idx), -1) if len(lower_idx) else np.zeros((0, len(lower_off) - 1), np.int32)
        atom_entity, atom_rank, lower_off, lower_atoms = map(as_i32, (atom_entity, atom_rank, lower_off, lower_atoms))
        E_lower, Deriv_lower = as_f64(E_lower), as_f64(Deriv_lower)
        assert E.dtype == np.float64 and Deriv.dtype == np.float64 and E.flags.c_contiguous and Deriv.flags.c_contiguous
        n_slots = len(lower_off) - 1
        n_atoms_lower = Deriv_lower.shape[1] if Deriv_lower.ndim == 3 else int(lower_off[-1])
        check(self.lib.mbg_mbe_accumulate(
            self.handle, len(ref_rows), ptr(ref_rows, _lp), ptr(pair_offsets, _lp), ptr(lower_idx, _lp),
            ptr(slot_entity, _ip), n_slots, Deriv.shape[1], ptr(atom_entity, _ip), ptr(atom_rank, _ip), n_atoms_lower,
            ptr(lower_off, _ip), ptr(lower_atoms, _ip), ptr(E_lower, _dp), ptr(Deriv_lower, _dp), len(E_lower),
            float(sign), ptr(E, _dp), ptr(Deriv, _dp), len(E)))

    # -- training-side kernels ---------------------------------------------
    def kernel_mat(self, R_desc, R_d_desc, tril_perms_lin, sig, use_E_cstr=False, out=None):
        """Force-field kernel matrix (train.py:1105-1337).  ``out``: optional device pointer (int) of a
        [rows][rows] float64 buffer; otherwise a host array is returned."""
        R_desc, R_d_desc = as_f64(R_desc), as_f64(R_d_desc)
        tpl = np.ascontiguousarray(tril_perms_lin, dtype=np.int64)
        T, D = R_desc.shape
        n_atoms = int((1 + np.sqrt(8 * D + 1)) / 2)
        if n_atoms * (n_atoms - 1) // 2 != D or R_d_desc.shape != (T, D, 3) or tpl.size % D:
            raise ValueError("R_desc (T, D), R_d_desc (T, D, 3) and tril_perms_lin (P*D,) have inconsistent shapes")
        P = tpl.size // D
        rows = T * 3 * n_atoms + (T if use_E_cstr else 0)
        if out is None:
            K = np.empty((rows, rows))
            dst, flags = K.ctypes.data, 0
        else:
            K, dst, flags = None, int(out), FLAG_OUT_ON_DEVICE
        check(self.lib.mbg_kernel_mat(self.handle, n_atoms, T, P, ptr(R_desc, _dp), ptr(R_d_desc, _dp), ptr(tpl, _lp),
                                      float(sig), int(bool(use_E_cstr)), flags, C.c_void_p(dst)))
        return K

    def kernel_mat_cols(self, R_desc, R_d_desc, tril_perms_lin, sig, col_blocks, out=None):
        """The column blocks ``col_blocks`` (training points) of the force-field kernel matrix against all training
        points: [T 3N][len(col_blocks) 3N] (mbg_kernel_mat_cols).  ``out``: optional device pointer."""
        R_desc, R_d_desc = as_f64(R_desc), as_f64(R_d_desc)
        tpl = np.ascontiguousarray(tril_perms_lin, dtype=np.int64)
        blocks = np.ascontiguousarray(col_blocks, dtype=np.int32)
        T, D = R_desc.shape
        n_atoms = int((1 + np.sqrt(8 * D + 1)) / 2)
        if n_atoms * (n_atoms - 1) // 2 != D or R_d_desc.shape != (T, D, 3) or tpl.size % D:
            raise ValueError("R_desc (T, D), R_d_desc (T, D, 3) and tril_perms_lin (P*D,) have inconsistent shapes")
        P = tpl.size // D
        if out is None:
            K = np.empty((T * 3 * n_atoms, blocks.size * 3 * n_atoms))
            dst, flags = K.ctypes.data, 0
        else:
            K, dst, flags = None, int(out), FLAG_OUT_ON_DEVICE
        check(self.lib.mbg_kernel_mat_cols(self.handle, n_atoms, T, P, ptr(R_desc, _dp), ptr(R_d_desc, _dp), ptr(tpl, _lp),
                                           float(sig), ptr(blocks, _ip), int(blocks.size), flags,
                                           C.c_void_p(dst)))
        return K

    def mat52(self, model_handle, R, n_rows):
        """Covariances of structures R (n_R, n_atoms, 3) vs. the T*P permuted training rows of a device model."""
        R = as_f64(R)
        out = np.empty((R.shape[0], n_rows))
        check(self.lib.mbg_mat52(self.handle, model_handle, C.c_void_p(R.ctypes.data), R.shape[0], 0,
                                 C.c_void_p(out.ctypes.data)))
        return out

    def mat52_rows(self, r_desc, sig, rows):
        r_desc, rows = as_f64(r_desc), as_f64(rows)
        q = r_desc.reshape(-1, rows.shape[1])
        out = np.empty((q.shape[0], rows.shape[0]))
        check(self.lib.mbg_mat52_rows(self.handle, rows.shape[1], ptr(q, _dp), q.shape[0], float(sig), ptr(rows, _dp),
                                      rows.shape[0], ptr(out, _dp)))
        return out

    def d_mic(self, cell, cutoff, d, check_cutoff=True, pbc=None):
        d = as_f64(d)
        cell = as_f64(cell).reshape(9)
        out = np.zeros_like(d)
        acc = C.c_int32(0)
        check(self.lib.mbg_d_mic(self.handle, ptr(cell, _dp), float(cutoff), ptr(d, _dp), d.shape[0], int(bool(check_cutoff)),
                                 ptr(out, _dp), C.byref(acc), ptr(pbc, _ip)))
        return out, bool(acc.value)

    def r_mic(self, cell, cutoff, r, pbc=None):
        r = as_f64(r)
        cell = as_f64(cell).reshape(9)
        out = np.zeros_like(r)
        acc = C.c_int32(0)
        check(self.lib.mbg_r_mic(self.handle, ptr(cell, _dp), float(cutoff), ptr(r, _dp), r.shape[0], ptr(out, _dp),
                                 C.byref(acc), ptr(pbc, _ip)))
        return out, bool(acc.value)


class Step:
    """A prepared evaluation (mbg_step_create): tables on the device, the whole pipeline in one CUDA graph."""

    def __init__(self, ctx, system_desc, n_frames, n_atoms, jobs, n_jobs, want_virial, shard=(0, 1)):
        self.ctx, self.n_frames, self.n_atoms, self.n_jobs, self.want_virial = ctx, n_frames, n_atoms, n_jobs, want_virial
        h = C.c_void_p()
        check(ctx.lib.mbg_step_create(ctx.handle, C.byref(system_desc), n_frames, jobs, n_jobs,
                                      FLAG_VIRIAL if want_virial else 0, int(shard[0]), int(shard[1]), C.byref(h)))
        self.handle = h

    def run_device(self, R_ptr, E_ptr, F_ptr, V_ptr=None, cells=None, cutoffs=None):
        """Device pointers in and out; enqueued on the context's stream, not synchronised."""
        cells = as_f64(cells) if cells is not None else None
        cutoffs = as_f64(cutoffs) if cutoffs is not None else None
        check(self.ctx.lib.mbg_step_run_device(
            self.handle, R_ptr, cells.ctypes.data if cells is not None else None,
            cutoffs.ctypes.data if cutoffs is not None else None, E_ptr, F_ptr, V_ptr))

    def stats(self):
        st = (JobStats * self.n_jobs)()
        check(self.ctx.lib.mbg_step_stats(self.handle, st))
        return [(s_.n_candidates, s_.n_enumerated, s_.n_accepted) for s_ in st]

    def run(self, R, cells=None, cutoffs=None):
        R = as_f64(R)
        assert R.shape == (self.n_frames, self.n_atoms, 3)
        E = np.empty(self.n_frames)
        F = np.empty((self.n_frames, self.n_atoms, 3))
        V = np.empty((self.n_frames, 3, 3)) if self.want_virial else None
        stats = (JobStats * self.n_jobs)()
        cells = as_f64(cells) if cells is not None else None
        cutoffs = as_f64(cutoffs) if cutoffs is not None else None
        check(self.ctx.lib.mbg_step_run(
            self.handle, R.ctypes.data, cells.ctypes.data if cells is not None else None,
            cutoffs.ctypes.data if cutoffs is not None else None, E.ctypes.data, F.ctypes.data,
            V.ctypes.data if V is not None else None, stats))
        return E, F, V, [(s_.n_candidates, s_.n_enumerated, s_.n_accepted) for s_ in stats]

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.lib.mbg_step_destroy(self.handle)
        self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_contexts = {}


def get_context(device=None):
    """Process-wide context of a device (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
        try:
            import torch

            if torch.cuda.is_available() and torch.cuda.is_initialized():
                device = torch.cuda.current_device()
        except Exception:  # pragma: no cover
            pass
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx
