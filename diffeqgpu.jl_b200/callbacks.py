"""Discrete callbacks and CallbackSet for the kernel path.

Mirror of reference src/ensemblegpukernel/callbacks.jl:1-36 (GPUDiscreteCallback) and the
`callback = ...`, `tstops = ...` keywords of vectorized_solve / vectorized_asolve
(lowerlevel_solve.jl:53-59, 253-260).  A Julia callback is a pair of closures; here the pair
arrives lowered to CUDA-C function BODIES (the `ext/` binding produces them from the Julia
side the same way it lowers the RHS):

    condition(u, t, integrator)  ->  body of `bool condition(u, p, t)`, e.g. "return t == (T)2.4;"
    affect!(integrator)          ->  body that may assign u[i], p[i] and call terminate(),
                                     e.g. "u[0] = u[0] + (T)10;"

ContinuousCallback's condition body returns the root function's value; the kernels restate the ITP root finder
and the event handling of integrator_utils.jl:186-479.
"""


class DiscreteCallback:
    def __init__(self, condition, affect, *, save_positions=(False, False), initialize=None, finalize=None):
        if tuple(save_positions) != (False, False):
            # callbacks.jl:12-14
            raise ValueError("Callback `save_positions` are incompatible with kernel-based GPU ODE solvers due "
                             "requiring static sizing. Please ensure `save_positions = (false,false)` is set in "
                             "all callback definitions used with such solvers.")
        if not isinstance(condition, str) or not isinstance(affect, str):
            raise TypeError("condition and affect must be CUDA-C function bodies (str)")
        self.condition, self.affect = condition, affect

    @classmethod
    def from_python(cls, condition, affect, n_state, n_param=0, **kw):
        """condition(u, t, p) -> comparison, affect(u, p, t) -> new u | (new u, new p) | 'terminate': host
        functions traced once and lowered to the bodies above (lowering.py)"""
        from . import lowering
        return cls(lowering.lower_condition(condition, n_state, n_param),
                   lowering.lower_affect(affect, n_state, n_param), **kw)


GPUDiscreteCallback = DiscreteCallback


class ContinuousCallback:
    """ContinuousCallback(condition, affect!; affect_neg! = affect!, rootfind = LeftRootFind, abstol, repeat_nudge,
    dtrelax, save_positions = (false, false))  (callbacks.jl:38-124).  `condition` is a body RETURNING the root
    function's value, e.g. "return u[0];"; `affect` / `affect_neg` are bodies like a DiscreteCallback's affect
    (None = `nothing`, "same" = affect_neg! defaults to affect!)."""

    ROOTFIND = {"left": 0, "right": 1, "none": 2}

    def __init__(self, condition, affect, *, affect_neg="same", rootfind="left", save_positions=(False, False),
                 abstol=10 * 1.1920928955078125e-07, reltol=0, repeat_nudge=0.01, dtrelax=1, interp_points=10,
                 initialize=None, finalize=None):
        if tuple(save_positions) != (False, False):      # callbacks.jl:62-64
            raise ValueError("Callback `save_positions` are incompatible with kernel-based GPU ODE solvers due "
                             "requiring static sizing. Please ensure `save_positions = (false,false)` is set in "
                             "all callback definitions used with such solvers.")
        if not isinstance(condition, str) or not (affect is None or isinstance(affect, str)):
            raise TypeError("condition and affect must be CUDA-C function bodies (str)")
        if rootfind not in self.ROOTFIND:
            raise ValueError("rootfind must be 'left', 'right' or 'none'")
        self.condition, self.affect = condition, affect
        self.affect_neg = affect if affect_neg == "same" else affect_neg
        self.rootfind, self.abstol, self.repeat_nudge, self.dtrelax = rootfind, float(abstol), float(repeat_nudge), float(dtrelax)

    def key(self):
        return (self.condition, self.affect, self.affect_neg, self.ROOTFIND[self.rootfind], self.abstol, self.repeat_nudge,
                self.dtrelax)

    @classmethod
    def from_python(cls, condition, affect, n_state, n_param=0, *, affect_neg="same", **kw):
        """condition(u, t, p) -> value of the root function, affect / affect_neg as DiscreteCallback.from_python
        (None = `nothing`)"""
        from . import lowering
        low = lambda a: None if a is None else lowering.lower_affect(a, n_state, n_param)   # noqa: E731
        return cls(lowering.lower_condition(condition, n_state, n_param, continuous=True), low(affect),
                   affect_neg="same" if isinstance(affect_neg, str) and affect_neg == "same" else low(affect_neg), **kw)


GPUContinuousCallback = ContinuousCallback


class CallbackSet:
    """CallbackSet(cb1, cb2, ...): discrete callbacks run in order after every step."""

    def __init__(self, *callbacks):
        flat, cont = [], []
        for c in callbacks:
            if c is None:
                continue
            if isinstance(c, CallbackSet):
                flat.extend(c.discrete_callbacks)
                cont.extend(c.continuous_callbacks)
            elif isinstance(c, DiscreteCallback):
                flat.append(c)
            elif isinstance(c, ContinuousCallback):
                cont.append(c)
            else:
                raise TypeError(f"unsupported callback {c!r}")
        self.discrete_callbacks = tuple(flat)
        self.continuous_callbacks = tuple(cont)

    def key(self):
        return tuple((c.condition, c.affect) for c in self.discrete_callbacks)

    def ckey(self):
        return tuple(c.key() for c in self.continuous_callbacks)

    def __len__(self):
        return len(self.discrete_callbacks) + len(self.continuous_callbacks)


def as_callback_set(cb):
    if cb is None:
        return CallbackSet()
    return cb if isinstance(cb, CallbackSet) else CallbackSet(cb)
