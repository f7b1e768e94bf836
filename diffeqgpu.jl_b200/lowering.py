"""RHS lowering front-end: a host function f(u, p, t) -> CUDA C++ bodies for the NVRTC path.

Reference: SURVEY §8(f) row 4.  In the reference a Julia closure is compiled for the device by
GPUCompiler, and symbolic models arrive through ModelingToolkit / Symbolics
(docs/src/tutorials/modelingtoolkit.md: `mtkcompile` + `ODEProblem{false}(sys, op, tspan)`, with
`jac = true` building the analytic Jacobian symbolically).  A closure cannot cross a C ABI, so the
equivalent here is what `Symbolics.build_function(...; target = CTarget())` does: the function is
*traced* once with symbolic arguments, the expression graph is simplified by common-subexpression
elimination only (no re-association beyond what the tracer's canonical form implies), and printed
as a body over `u[i]`, `p[i]`, `t` in the kernel's scalar type `T` -- float, double, the packed
pair type or a forward-mode dual, which is why every literal is written `(T)c` and only functions
that exist for all four are emitted (sqrt, sin, cos, exp, log; tan / sinh / cosh / tanh are rewritten).

    f = dg.ODEFunction.from_python(lambda u, p, t: [p[0] * (u[1] - u[0]), ...], n_state=3, n_param=3, jac=True)

`jac=True` adds the symbolic Jacobian and time gradient (`ODEFunction(f; jac, tgrad)`); without it
the stiff solvers differentiate the body themselves (duals / finite differences).
numpy ufuncs (`np.sin(u[0])`) work on the traced values as well as the `sin`, `cos`, ... of this module.
"""
import numpy as np

try:
    import sympy
    from sympy.printing.c import C99CodePrinter
except ImportError as e:      # pragma: no cover - sympy ships with torch
    raise ImportError("diffeqgpu_b200.lowering needs sympy") from e

__all__ = ["lower_function", "lower_noise", "lower_condition", "lower_affect", "sin", "cos", "tan", "exp", "log", "sqrt",
           "sinh", "cosh", "tanh", "LoweringError"]


class LoweringError(ValueError):
    """the traced function uses something that has no counterpart in the device scalar types"""


class Traced:
    """a traced scalar: Python arithmetic and numpy ufuncs build a sympy expression"""
    __slots__ = ("e",)

    def __init__(self, e):
        self.e = sympy.sympify(e)

    @staticmethod
    def _x(o):
        if isinstance(o, Traced):
            return o.e
        if isinstance(o, (bool, np.bool_)):
            raise LoweringError("booleans cannot enter arithmetic of a traced function")
        if isinstance(o, (int, np.integer)):
            return sympy.Integer(int(o))
        if isinstance(o, (float, np.floating)):
            return sympy.Float(float(o), 17)
        if isinstance(o, sympy.Expr):
            return o
        raise LoweringError(f"cannot trace a value of type {type(o).__name__}")

    def _bin(self, o, fn):
        if isinstance(o, np.ndarray):       # let numpy broadcast element by element (object arrays of traced values)
            return NotImplemented
        return Traced(fn(self.e, self._x(o)))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __rpow__(self, o): return self._bin(o, lambda a, b: b ** a)
    def __neg__(self): return Traced(-self.e)
    def __pos__(self): return self

    def __lt__(self, o): return TracedBool(sympy.StrictLessThan(self.e, self._x(o)))
    def __le__(self, o): return TracedBool(sympy.LessThan(self.e, self._x(o)))
    def __gt__(self, o): return TracedBool(sympy.StrictGreaterThan(self.e, self._x(o)))
    def __ge__(self, o): return TracedBool(sympy.GreaterThan(self.e, self._x(o)))
    def __eq__(self, o): return TracedBool(sympy.Eq(self.e, self._x(o), evaluate=False))
    def __ne__(self, o): return TracedBool(sympy.Ne(self.e, self._x(o), evaluate=False))
    __hash__ = None

    def __bool__(self):
        raise LoweringError("data-dependent Python control flow cannot be traced (`if u[0] > 0:`); "
                            "write the branch as arithmetic")

    def __abs__(self):
        raise LoweringError("abs() has no counterpart for the packed / dual device types")

    # numpy ufuncs on object scalars call these
    def sin(self): return Traced(sympy.sin(self.e))
    def cos(self): return Traced(sympy.cos(self.e))
    def tan(self): return Traced(sympy.sin(self.e) / sympy.cos(self.e))
    def exp(self): return Traced(sympy.exp(self.e))
    def log(self): return Traced(sympy.log(self.e))
    def sqrt(self): return Traced(sympy.sqrt(self.e))
    def sinh(self): return Traced((sympy.exp(self.e) - sympy.exp(-self.e)) / 2)
    def cosh(self): return Traced((sympy.exp(self.e) + sympy.exp(-self.e)) / 2)
    def tanh(self): return Traced(1 - 2 / (sympy.exp(2 * self.e) + 1))

    def __repr__(self):
        return f"Traced({self.e})"


class TracedBool:
    """a traced comparison (only meaningful as the value of a discrete-callback condition)"""
    __slots__ = ("e",)

    def __init__(self, e):
        self.e = e

    def __and__(self, o): return TracedBool(sympy.And(self.e, _rel(o)))
    def __or__(self, o): return TracedBool(sympy.Or(self.e, _rel(o)))
    def __invert__(self): return TracedBool(sympy.Not(self.e))

    def __bool__(self):
        raise LoweringError("data-dependent Python control flow cannot be traced; combine comparisons with & | ~")


def _rel(o):
    if isinstance(o, TracedBool):
        return o.e
    if isinstance(o, (bool, np.bool_)):
        return sympy.true if o else sympy.false
    raise LoweringError("expected a traced comparison")


def _unary(name):
    def fn(x):
        if isinstance(x, Traced):
            return getattr(x, name)()
        return getattr(np, name)(x)
    fn.__name__ = name
    return fn


sin, cos, tan, exp, log, sqrt = (_unary(n) for n in ("sin", "cos", "tan", "exp", "log", "sqrt"))
sinh, cosh, tanh = (_unary(n) for n in ("sinh", "cosh", "tanh"))


class _Printer(C99CodePrinter):
    """C body in the kernel scalar type T: literals are cast, integer powers are products"""

    def _print_Float(self, e):
        return f"(T){float(e)!r}"

    def _print_Integer(self, e):
        return f"(T){int(e)}"

    def _print_Rational(self, e):
        return f"(T)({int(e.p)}.0 / {int(e.q)}.0)"

    def _print_NumberSymbol(self, e):
        return f"(T){float(e.evalf(17))!r}"

    _print_Pi = _print_Exp1 = _print_NumberSymbol

    def _print_Pow(self, e):
        b, x = e.base, e.exp
        if x.is_Float and float(2 * x) == int(float(2 * x)):      # u ** 2.0, u ** 0.5 written with float literals
            x = sympy.Rational(int(float(2 * x)), 2)
        bs = self.parenthesize(b, 1000, strict=True)        # atoms bare, everything else parenthesised
        if x.is_Integer and 1 <= abs(int(x)) <= 16:
            prod = " * ".join([bs] * abs(int(x)))
            return f"({prod})" if int(x) > 0 else f"((T)1 / ({prod}))"
        if x == sympy.Rational(1, 2):
            return f"sqrt({self._print(b)})"
        if x == sympy.Rational(-1, 2):
            return f"((T)1 / sqrt({self._print(b)}))"
        if x.is_Rational and x.q == 2 and abs(x.p) <= 15:      # x^(k/2) = sqrt(x)^k
            s = f"sqrt({self._print(b)})"
            prod = " * ".join([s] * abs(int(x.p)))
            return f"({prod})" if x.p > 0 else f"((T)1 / ({prod}))"
        # general real power through exp/log (both exist for every device scalar type)
        return f"exp({self._print(x)} * log({self._print(b)}))"

    def _print_Function(self, e):
        name = e.func.__name__
        if name in ("sin", "cos", "exp", "log"):
            return f"{name}({self._print(e.args[0])})"
        raise LoweringError(f"function `{name}` has no counterpart for the device scalar types "
                            f"(available: sqrt, sin, cos, tan, exp, log, sinh, cosh, tanh)")

    _print_sin = _print_cos = _print_exp = _print_log = _print_Function

    def _print_Abs(self, e):
        raise LoweringError("abs() has no counterpart for the packed / dual device types")

    def _print_Relational(self, e):
        return f"({self._print(e.lhs)} {e.rel_op} {self._print(e.rhs)})"

    def _print_And(self, e):
        return "(" + " && ".join(self._print(a) for a in e.args) + ")"

    def _print_Or(self, e):
        return "(" + " || ".join(self._print(a) for a in e.args) + ")"

    def _print_Not(self, e):
        return f"(!{self._print(e.args[0])})"

    def _print_BooleanTrue(self, e):
        return "true"

    def _print_BooleanFalse(self, e):
        return "false"


_printer = _Printer()


def _symbols(n_state, n_param):
    u = [sympy.Symbol(f"u[{i}]", real=True) for i in range(n_state)]
    p = [sympy.Symbol(f"p[{i}]", real=True) for i in range(n_param)]
    return u, p, sympy.Symbol("t", real=True)


def _trace(f, n_state, n_param, what, shape=None):
    """call f on traced arguments -> (array of sympy expressions, (u, p, t) symbols)"""
    u, p, t = _symbols(n_state, n_param)
    out = f([Traced(s) for s in u], [Traced(s) for s in p], Traced(t))
    arr = np.empty(np.shape(out), dtype=object)
    for idx, v in np.ndenumerate(np.asarray(out, dtype=object)):
        arr[idx] = Traced._x(v)
    if shape is not None and arr.shape != shape:
        raise LoweringError(f"{what} returned shape {arr.shape}, expected {shape}")
    return arr, (u, p, t)


def _emit(assignments, indent="    "):
    """assignments: [(lhs string, sympy expr)] -> body with shared subexpressions hoisted"""
    exprs = [e for _, e in assignments]
    if not exprs:
        return ""
    repl, red = sympy.cse(exprs, symbols=sympy.numbered_symbols("x_"), order="none")
    # cse also hoists bare negations / single operations: put those back
    keep, back = [], {}
    for s, e in repl:
        e = e.xreplace(back)
        if sympy.count_ops(e) <= 1:
            back[s] = e
        else:
            keep.append((s, e))
    repl, red = keep, [e.xreplace(back) for e in red]
    lines = [f"{indent}const T {_printer.doprint(s)} = {_printer.doprint(e)};" for s, e in repl]
    lines += [f"{indent}{lhs} = {_printer.doprint(e)};" for (lhs, _), e in zip(assignments, red)]
    return "\n".join(lines) + "\n"


def lower_function(f, n_state, n_param, jac=False):
    """trace f(u, p, t) -> dict(rhs=, jac=, tgrad=) of CUDA C++ bodies (jac / tgrad None unless `jac`)"""
    du, (u, p, t) = _trace(f, n_state, n_param, "f(u, p, t)", (n_state,))
    out = dict(rhs=_emit([(f"du[{i}]", du[i]) for i in range(n_state)]), jac=None, tgrad=None)
    if jac:
        # only the structural non-zeros: the kernel zero-initialises J and dT
        ja = [(f"J[{i}][{j}]", sympy.diff(du[i], u[j])) for i in range(n_state) for j in range(n_state)]
        out["jac"] = _emit([(l, e) for l, e in ja if e != 0])
        tg = [(f"dT[{i}]", sympy.diff(du[i], t)) for i in range(n_state)]
        # an autonomous right-hand side has no time gradient: leaving the body out (instead of an empty one) tells the
        # library so, and the fast build drops the dT terms of the Rosenbrock stages (TGRAD_ZERO, degk_jit.cpp)
        out["tgrad"] = _emit([(l, e) for l, e in tg if e != 0]) or None
    return out


def lower_noise(g, n_state, n_param, noise="diagonal", n_noise=0):
    """trace g(u, p, t): diagonal noise -> `g[i]`, general noise (n x m matrix) -> `G[i][j]`"""
    if noise == "diagonal":
        gv, _ = _trace(g, n_state, n_param, "g(u, p, t)", (n_state,))
        return _emit([(f"g[{i}]", gv[i]) for i in range(n_state)])
    gm, _ = _trace(g, n_state, n_param, "g(u, p, t)", (n_state, n_noise))
    return _emit([(f"G[{i}][{j}]", gm[i, j]) for i in range(n_state) for j in range(n_noise) if gm[i, j] != 0])


def lower_condition(condition, n_state, n_param, continuous=False):
    """condition(u, t, p) of a callback (argument order of the reference's `condition(u, t, integrator)`):
    a traced comparison for a DiscreteCallback, a traced value for a ContinuousCallback -> `return ...;`"""
    u, p, t = _symbols(n_state, n_param)
    out = condition([Traced(s) for s in u], Traced(t), [Traced(s) for s in p])
    if continuous:
        return f"    return {_printer.doprint(Traced._x(out))};\n"
    if isinstance(out, (bool, np.bool_)):
        return f"    return {'true' if out else 'false'};\n"
    if not isinstance(out, TracedBool):
        raise LoweringError("a discrete condition must return a comparison (u[0] > 3, t == 2.4, ...)")
    return f"    return {_printer.doprint(out.e)};\n"


def lower_affect(affect, n_state, n_param):
    """affect(u, p, t) -> new u (sequence), or (new u, new p), or the string 'terminate':
    the out-of-place form of `affect!(integrator)`; every new value is computed from the old state first"""
    u, p, t = _symbols(n_state, n_param)
    out = affect([Traced(s) for s in u], [Traced(s) for s in p], Traced(t))
    if isinstance(out, str):
        if out != "terminate":
            raise LoweringError("an affect returns the new u, (new u, new p) or 'terminate'")
        return "    terminate();\n"
    new_p = None
    if isinstance(out, tuple) and len(out) == 2 and np.ndim(out[0]) == 1:
        out, new_p = out
    nu = [Traced._x(v) for v in out]
    if len(nu) != n_state:
        raise LoweringError(f"affect returned {len(nu)} state values, expected {n_state}")
    asg = [(f"const T nu_{i}", nu[i]) for i in range(n_state) if nu[i] != u[i]]
    fin = [f"    u[{i}] = nu_{i};" for i in range(n_state) if nu[i] != u[i]]
    if new_p is not None:
        np_ = [Traced._x(v) for v in new_p]
        if len(np_) != n_param:
            raise LoweringError(f"affect returned {len(np_)} parameters, expected {n_param}")
        asg += [(f"const T np_{i}", np_[i]) for i in range(n_param) if np_[i] != p[i]]
        fin += [f"    p[{i}] = np_{i};" for i in range(n_param) if np_[i] != p[i]]
    return _emit(asg) + "\n".join(fin) + ("\n" if fin else "")
