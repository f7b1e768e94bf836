"""Discrete callbacks and CallbackSet for the kernel path.

Mirror of reference src/ensemblegpukernel/callbacks.jl:1-36 (GPUDiscreteCallback) and the
`callback = ...`, `tstops = ...` keywords of vectorized_solve / vectorized_asolve
(lowerlevel_solve.jl:53-59, 253-260).  A Julia callback is a pair of closures; here the pair
arrives lowered to CUDA-C function BODIES (the `ext/` binding produces them from the Julia
side the same way it lowers the RHS):

    condition(u, t, integrator)  ->  body of `bool condition(u, p, t)`, e.g. "return t == (T)2.4;"
    affect!(integrator)          ->  body that may assign u[i], p[i] and call terminate(),
                                     e.g. "u[0] = u[0] + (T)10;"

ContinuousCallback (root finding, integrator_utils.jl:331-479) is not lowered.
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


GPUDiscreteCallback = DiscreteCallback


class ContinuousCallback:
    def __init__(self, *a, **k):
        raise NotImplementedError("ContinuousCallback (ITP root finding, integrator_utils.jl:331-479) is not "
                                  "lowered to the C ABI; see DESIGN.md §7")


class CallbackSet:
    """CallbackSet(cb1, cb2, ...): discrete callbacks run in order after every step."""

    def __init__(self, *callbacks):
        flat = []
        for c in callbacks:
            if c is None:
                continue
            if isinstance(c, CallbackSet):
                flat.extend(c.discrete_callbacks)
            elif isinstance(c, DiscreteCallback):
                flat.append(c)
            else:
                raise TypeError(f"unsupported callback {c!r}")
        self.discrete_callbacks = tuple(flat)

    def key(self):
        return tuple((c.condition, c.affect) for c in self.discrete_callbacks)

    def __len__(self):
        return len(self.discrete_callbacks)


def as_callback_set(cb):
    if cb is None:
        return CallbackSet()
    return cb if isinstance(cb, CallbackSet) else CallbackSet(cb)
