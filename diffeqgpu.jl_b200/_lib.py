"""ctypes binding of libdegk.so (include/degk.h).

The product path has no fallback: if the shared library is missing this module raises, and
if no CUDA device is present `Context()` raises.  Nothing under oracle/ is imported here.
"""
import ctypes as C
import os
import subprocess
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libdegk.so"

OK, ERR_INVALID, ERR_CUDA, ERR_NVRTC, ERR_UNSUPPORTED, ERR_NOMEM = range(6)
F32, F64 = 0, 1
FP_STRICT, FP_FAST = 0, 1
LAYOUT_REF, LAYOUT_SOA = 0, 1
SCHED_STATIC, SCHED_QUEUE, SCHED_AUTO = 0, 1, 2
NOISE_NONE, NOISE_DIAGONAL, NOISE_GENERAL = 0, 1, 2
ENGINE_AUTO, ENGINE_V1, ENGINE_LOCKSTEP = 0, 1, 2
ENGINES = {"auto": ENGINE_AUTO, "v1": ENGINE_V1, "lockstep": ENGINE_LOCKSTEP}
RETCODES = {0: "Default", 1: "Success", 2: "DtLessThanMin", 3: "Unstable", 4: "MaxIters",
            5: "Singular", 6: "Terminated", 7: "InitialFailure"}


class DegkError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libdegk error {status}: {msg}")
        self.status = status


class ModelDesc(C.Structure):
    _fields_ = [("builtin", C.c_char_p), ("rhs_src", C.c_char_p), ("jac_src", C.c_char_p),
                ("tgrad_src", C.c_char_p), ("noise_src", C.c_char_p),
                ("n_state", C.c_int32), ("n_param", C.c_int32), ("n_noise", C.c_int32),
                ("noise_kind", C.c_int32), ("dtype", C.c_int32), ("alg", C.c_int32),
                ("fp_mode", C.c_int32), ("force_jit", C.c_int32),
                ("events", C.c_int32), ("n_callbacks", C.c_int32),
                ("cb_condition_src", C.POINTER(C.c_char_p)), ("cb_affect_src", C.POINTER(C.c_char_p)),
                ("jac_mode", C.c_int32), ("reserved", C.c_int32), ("mass_src", C.c_char_p),
                ("n_ccallbacks", C.c_int32), ("reserved3", C.c_int32),
                ("cc_condition_src", C.POINTER(C.c_char_p)), ("cc_affect_src", C.POINTER(C.c_char_p)),
                ("cc_affect_neg_src", C.POINTER(C.c_char_p)), ("cc_rootfind", C.POINTER(C.c_int32)),
                ("cc_abstol", C.POINTER(C.c_double)), ("cc_repeat_nudge", C.POINTER(C.c_double)),
                ("cc_dtrelax", C.POINTER(C.c_double))]


class ProgramInfo(C.Structure):
    _fields_ = [("n_state", C.c_int32), ("n_param", C.c_int32), ("n_noise", C.c_int32),
                ("noise_kind", C.c_int32), ("dtype", C.c_int32), ("alg", C.c_int32),
                ("fp_mode", C.c_int32), ("is_jit", C.c_int32),
                ("regs_fixed", C.c_int32), ("regs_adaptive", C.c_int32),
                ("local_bytes_fixed", C.c_int32), ("local_bytes_adaptive", C.c_int32),
                ("max_blocks_per_sm", C.c_int32),
                ("regs_adaptive2", C.c_int32), ("local_bytes_adaptive2", C.c_int32),
                ("slots_per_thread2", C.c_int32), ("max_blocks_per_sm2", C.c_int32),
                ("jit_seconds", C.c_double)]


class SolveArgs(C.Structure):
    _fields_ = [("n_traj", C.c_int64), ("traj_offset", C.c_int64),
                ("u0", C.c_void_p), ("u0_stride", C.c_int64),
                ("p", C.c_void_p), ("p_stride", C.c_int64),
                ("tspan", C.c_void_p), ("tspan_stride", C.c_int64),
                ("dt", C.c_double), ("adaptive", C.c_int32),
                ("abstol", C.c_double), ("reltol", C.c_double),
                ("saveat", C.c_void_p), ("n_saveat", C.c_int32), ("save_everystep", C.c_int32),
                ("n_rows", C.c_int64), ("us", C.c_void_p), ("ts", C.c_void_p),
                ("out_layout", C.c_int32), ("schedule", C.c_int32),
                ("retcode", C.c_void_p), ("naccept", C.c_void_p), ("nreject", C.c_void_p),
                ("seed", C.c_uint64), ("reduce", C.c_void_p), ("totals", C.c_void_p),
                ("max_iters", C.c_int64), ("engine", C.c_int32), ("dae_init", C.c_int32),
                ("tstops", C.c_void_p), ("n_tstops", C.c_int32), ("reserved2", C.c_int32),
                ("nsaved", C.c_void_p), ("saveat_stride", C.c_int64), ("order", C.c_void_p)]


# every symbol include/degk.h declares (checked by tests/test_abi.py)
API_SYMBOLS = ["degk_version", "degk_ctx_create", "degk_ctx_destroy", "degk_last_error",
               "degk_program_build", "degk_program_get_info", "degk_program_destroy",
               "degk_builtin_count", "degk_builtin_name", "degk_output_rows", "degk_solve",
               "degk_solve_host", "degk_debug_philox", "degk_jit_compile_check"]

_lib = None
_lock = threading.Lock()


def build_library(force=False):
    """Compile libdegk.so in-tree (nvcc, sm_100a).  Used by __graft_entry__.build()."""
    csrc = _PKG / "csrc"
    if force:
        subprocess.check_call(["make", "-s", "-C", str(csrc), "clean"])
    subprocess.check_call(["make", "-s", "-j", str(os.cpu_count() or 4), "-C", str(csrc)])
    return LIB_PATH


def lib():
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise DegkError(-1, f"{LIB_PATH} is missing: build it with "
                                    f"`make -C {_PKG / 'csrc'}` (there is no CPU fallback)")
            L = C.CDLL(str(LIB_PATH))
            L.degk_version.restype = C.c_int
            L.degk_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
            L.degk_ctx_destroy.argtypes = [C.c_void_p]
            L.degk_ctx_destroy.restype = None
            L.degk_last_error.argtypes = [C.c_void_p]
            L.degk_last_error.restype = C.c_char_p
            L.degk_program_build.argtypes = [C.c_void_p, C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
            L.degk_program_get_info.argtypes = [C.c_void_p, C.POINTER(ProgramInfo)]
            L.degk_program_destroy.argtypes = [C.c_void_p]
            L.degk_program_destroy.restype = None
            L.degk_builtin_count.restype = C.c_int
            L.degk_builtin_name.argtypes = [C.c_int]
            L.degk_builtin_name.restype = C.c_char_p
            L.degk_output_rows.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                           C.c_int, C.c_int]
            L.degk_output_rows.restype = C.c_int64
            L.degk_solve.argtypes = [C.c_void_p, C.POINTER(SolveArgs), C.c_void_p]
            L.degk_solve_host.argtypes = [C.c_void_p, C.POINTER(SolveArgs), C.c_int64]
            L.degk_debug_philox.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_int64, C.c_void_p]
            L.degk_jit_compile_check.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_int64),
                                                 C.c_char_p, C.c_int64]
            _lib = L
    return _lib


def _b(s):
    return None if s is None else s.encode()


def make_desc(*, builtin=None, rhs_src=None, jac_src=None, tgrad_src=None, noise_src=None,
              n_state=0, n_param=0, n_noise=0, noise_kind=NOISE_NONE, dtype=F32, alg=0,
              fp_mode=FP_STRICT, force_jit=False, events=False, callbacks=(), jac_mode=0, mass_src=None, ccallbacks=()):
    """callbacks: sequence of (condition_src, affect_src) CUDA-C bodies (degk.h, degk_model_desc)."""
    d = ModelDesc(_b(builtin), _b(rhs_src), _b(jac_src), _b(tgrad_src), _b(noise_src),
                  n_state, n_param, n_noise, noise_kind, dtype, alg, fp_mode, int(force_jit))
    d.events = int(bool(events) or len(callbacks) > 0)
    d.n_callbacks = len(callbacks)
    d.jac_mode = int(jac_mode)
    d.mass_src = _b(mass_src)
    # ccallbacks: sequence of (condition_src, affect_src | None, affect_neg_src | None, rootfind, abstol, repeat_nudge, dtrelax)
    d.n_ccallbacks = len(ccallbacks)
    if ccallbacks:
        n = len(ccallbacks)
        arrs = ((C.c_char_p * n)(*[_b(c[0]) for c in ccallbacks]), (C.c_char_p * n)(*[_b(c[1]) for c in ccallbacks]),
                (C.c_char_p * n)(*[_b(c[2]) for c in ccallbacks]), (C.c_int32 * n)(*[int(c[3]) for c in ccallbacks]),
                (C.c_double * n)(*[float(c[4]) for c in ccallbacks]), (C.c_double * n)(*[float(c[5]) for c in ccallbacks]),
                (C.c_double * n)(*[float(c[6]) for c in ccallbacks]))
        (d.cc_condition_src, d.cc_affect_src, d.cc_affect_neg_src, d.cc_rootfind, d.cc_abstol, d.cc_repeat_nudge,
         d.cc_dtrelax) = arrs
        d.events = 1
        d._keepalive_cc = arrs
    if callbacks:
        conds = (C.c_char_p * len(callbacks))(*[_b(c[0]) for c in callbacks])
        affs = (C.c_char_p * len(callbacks))(*[_b(c[1]) for c in callbacks])
        d.cb_condition_src, d.cb_affect_src = conds, affs
        d._keepalive = (conds, affs)
    return d


def jit_compile_check(desc):
    """NVRTC-compile a description without a GPU; returns (status, cubin_bytes, log)."""
    n = C.c_int64(0)
    buf = C.create_string_buffer(16384)
    st = lib().degk_jit_compile_check(C.byref(desc), C.byref(n), buf, len(buf))
    return st, n.value, buf.value.decode(errors="replace")


class Context:
    """degk_ctx: one per (process, device)."""

    def __init__(self, device=-1):
        h = C.c_void_p()
        st = lib().degk_ctx_create(device, C.byref(h))
        if st != OK:
            raise DegkError(st, lib().degk_last_error(None).decode())
        self._h = h
        self._programs = {}

    def last_error(self):
        return lib().degk_last_error(self._h).decode()

    def check(self, st):
        if st != OK:
            raise DegkError(st, self.last_error())

    def program(self, desc, key=None):
        if key is not None and key in self._programs:
            return self._programs[key]
        h = C.c_void_p()
        self.check(lib().degk_program_build(self._h, C.byref(desc), C.byref(h)))
        prog = Program(self, h)
        if key is not None:
            self._programs[key] = prog
        return prog

    def debug_philox(self, c0, c1, k0, k1, n):
        import numpy as np
        out = np.zeros(4 * n, np.uint32)
        self.check(lib().degk_debug_philox(self._h, c0, c1, k0, k1, n, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(n, 4)

    def close(self):
        if self._h:
            for p in self._programs.values():
                p.close()
            self._programs.clear()
            lib().degk_ctx_destroy(self._h)
            self._h = None


class Program:
    def __init__(self, ctx, h):
        self.ctx = ctx
        self._h = h
        info = ProgramInfo()
        lib().degk_program_get_info(h, C.byref(info))
        self.info = info

    def solve(self, args, stream=None):
        self.ctx.check(lib().degk_solve(self._h, C.byref(args), C.c_void_p(stream or 0)))

    def solve_host(self, args, chunk_traj=0):
        self.ctx.check(lib().degk_solve_host(self._h, C.byref(args), chunk_traj))

    def close(self):
        if self._h:
            lib().degk_program_destroy(self._h)
            self._h = None


_ctxs = {}


def context(device=None):
    """Process-wide context cache keyed by device index."""
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    if device not in _ctxs:
        _ctxs[device] = Context(device)
    return _ctxs[device]
