"""ctypes wrapper around the CPU oracle (oracle/degk_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libdegk_oracle.so"
_SRCS = ("degk_oracle.cpp", "oracle_stiff.inc", "oracle_kvaerno.inc", "oracle_sde.inc", "oracle_tables.inc")

MODELS = {"lorenz": 0, "henon_heiles": 1, "rober": 2, "decay": 3, "linear15": 4, "gbm": 5,
          "lorenz_additive": 6, "scalar_sde": 7, "osc_t": 8, "gbm_nd": 9, "quad_decay": 10, "rober_dae": 11, "ball": 12, "lin_dae": 13}
ALGS = {"tsit5": 0, "vern7": 1, "vern9": 2, "rosenbrock23": 3, "rodas4": 4, "rodas5p": 5,
        "em": 6, "siea": 7, "kvaerno3": 8, "kvaerno5": 9}
RETCODES = {0: "Default", 1: "Success", 2: "DtLessThanMin", 3: "Unstable", 4: "MaxIters",
            5: "Singular", 6: "Terminated"}
# discrete-callback specs shared with tests/cases.py (which lowers them to CUDA-C for the device)
COND_KINDS = {"t_eq": 0, "u_lt": 1, "u_gt": 2, "t_ge": 3}
AFFECT_KINDS = {"u_add": 0, "u_set": 1, "u_scale": 2, "terminate": 3, "p_set": 4}
CC_COND_KINDS = {"u_minus": 0, "t_minus": 1}      # continuous conditions: u[i] - v, t - v

_lib = None


def build(force=False):
    """Compile the oracle if the .so is missing or older than its sources."""
    stale = force or not _LIB_PATH.exists() or any(
        (_HERE / f).stat().st_mtime > _LIB_PATH.stat().st_mtime for f in _SRCS)
    if stale:
        subprocess.check_call(["make", "-s", "-C", str(_HERE)])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.degk_oracle_solve.restype = ctypes.c_int
        _lib.degk_oracle_solve.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int]
        _lib.degk_oracle_solve_events.restype = ctypes.c_int
        _lib.degk_oracle_solve_events.argtypes = _lib.degk_oracle_solve.argtypes + [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.degk_oracle_num_threads.restype = ctypes.c_int
    return _lib


def model_info(model):
    n, np_, m, d = (ctypes.c_int() for _ in range(4))
    lib().degk_oracle_model_info(MODELS[model], ctypes.byref(n), ctypes.byref(np_),
                                 ctypes.byref(m), ctypes.byref(d))
    return n.value, np_.value, m.value, bool(d.value)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def solve(model, alg, u0, p, tspan, *, dt, adaptive=False, abstol=1e-6, reltol=1e-3,
          saveat=None, save_everystep=True, length=None, seed=0, dtype=np.float32,
          fma_stages=False, nthreads=0, tstops=None, callbacks=(), jac_mode=0, continuous_callbacks=()):
    """Solve a batch; returns dict(ts=(N,len), us=(N,len,n), naccept, nreject, retcode).

    u0: (N,n) or (n,) broadcast; p: (N,np) or (np,) broadcast; tspan: (2,) or (N,2).
    `length` = number of output rows (the host-side `len` of lowerlevel_solve.jl); required
    unless saveat is given (-> len(saveat)) or endpoints-only (-> 2).
    tstops: times the steppers must hit; callbacks: sequence of
    ((cond_kind, cond_idx, cond_val), (affect_kind, affect_idx, affect_val)) discrete callbacks
    (kinds: COND_KINDS / AFFECT_KINDS), applied in order after every step.
    jac_mode (stiff solvers): 0 analytic jac/tgrad, 1 finite differences, 2 forward-mode duals.
    """
    dtype = np.dtype(dtype)
    n, npar, _, _ = model_info(model)
    u0 = np.ascontiguousarray(u0, dtype=dtype)
    p = np.ascontiguousarray(p if p is not None else np.zeros(max(npar, 1)), dtype=dtype)
    tspan = np.ascontiguousarray(tspan, dtype=dtype)
    N = max(u0.shape[0] if u0.ndim == 2 else 1, p.shape[0] if p.ndim == 2 else 1,
            tspan.shape[0] if tspan.ndim == 2 else 1)
    u0s = n if u0.ndim == 2 else 0
    ps = p.shape[1] if p.ndim == 2 else 0
    tss = 2 if tspan.ndim == 2 else 0
    if saveat is not None:
        saveat = np.ascontiguousarray(saveat, dtype=dtype)
        length = len(saveat)
    elif length is None:
        if not save_everystep:
            length = 2
        else:
            raise ValueError("length required for save_everystep without saveat")
    t0 = tspan.reshape(-1, 2)[:, 0]
    ts = np.empty((N, length), dtype=dtype)
    ts[:] = t0[:, None] if tspan.ndim == 2 else t0[0]
    us = np.zeros((N, length, n), dtype=dtype)
    na = np.zeros(N, np.int32)
    nr = np.zeros(N, np.int32)
    rc = np.zeros(N, np.int32)
    args = [0 if dtype == np.float32 else 1, MODELS[model], ALGS[alg], int(adaptive), N,
            _ptr(u0), u0s, _ptr(p), ps, _ptr(tspan), tss,
            float(dtype.type(dt)), float(dtype.type(abstol)), float(dtype.type(reltol)),
            _ptr(saveat), 0 if saveat is None else len(saveat), int(save_everystep), int(seed),
            _ptr(us), _ptr(ts), length, _ptr(na), _ptr(nr), _ptr(rc), int(fma_stages), int(nthreads)]
    if tstops is not None or len(callbacks) or jac_mode or len(continuous_callbacks):
        # tstops reach the integrator already converted to the time type (adapt(backend, tstops))
        tst = np.ascontiguousarray([] if tstops is None else np.asarray(tstops, dtype=dtype), dtype=np.float64)
        cb_i = np.zeros((max(len(callbacks), 1), 4), np.int32)
        cb_v = np.zeros((max(len(callbacks), 1), 2), np.float64)
        for c, ((ck, ci, cv), (ak, ai, av)) in enumerate(callbacks):
            cb_i[c] = (COND_KINDS[ck], ci, AFFECT_KINDS[ak], ai)
            cb_v[c] = (float(dtype.type(cv)), float(dtype.type(av)))
        # continuous callbacks: dict(condition=(kind, idx, val), affect=(kind, idx, val) | None,
        #   affect_neg=(kind, idx, val) | None | "same", rootfind="left"|"right"|"none", abstol, repeat_nudge, dtrelax)
        cc_i = np.zeros((max(len(continuous_callbacks), 1), 8), np.int32)
        cc_v = np.zeros((max(len(continuous_callbacks), 1), 6), np.float64)
        for c, cc in enumerate(continuous_callbacks):
            ck, ci, cv = cc["condition"]
            aff = cc.get("affect")
            neg = cc.get("affect_neg", "same")
            neg = aff if neg == "same" else neg
            cc_i[c] = (CC_COND_KINDS[ck], ci, -1 if aff is None else AFFECT_KINDS[aff[0]], 0 if aff is None else aff[1],
                       -1 if neg is None else AFFECT_KINDS[neg[0]], 0 if neg is None else neg[1],
                       {"left": 0, "right": 1, "none": 2}[cc.get("rootfind", "left")], 0)
            cc_v[c] = (float(dtype.type(cv)), 0.0 if aff is None else float(dtype.type(aff[2])),
                       0.0 if neg is None else float(dtype.type(neg[2])),
                       float(dtype.type(cc.get("abstol", 10 * np.finfo(np.float32).eps))),
                       float(dtype.type(cc.get("repeat_nudge", 0.01))), float(dtype.type(cc.get("dtrelax", 1))))
        r = lib().degk_oracle_solve_events(*args, _ptr(tst), len(tst), _ptr(cb_i), _ptr(cb_v), len(callbacks), int(jac_mode),
                                           _ptr(cc_i), _ptr(cc_v), len(continuous_callbacks))
    else:
        r = lib().degk_oracle_solve(*args)
    if r != 0:
        raise RuntimeError(f"oracle error {r}")
    return dict(ts=ts, us=us, naccept=na, nreject=nr, retcode=rc)


def num_threads():
    return lib().degk_oracle_num_threads()
