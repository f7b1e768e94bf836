// degk_dual.cuh -- Jacobian and time gradient of a model without analytic `jac` / `tgrad`.
//
// Reference nlsolve/type.jl:129-157 (build_J_W, build_tgrad): `f.jac` if the function has one,
// else ForwardDiff.jacobian (alg autodiff = true, the default) else finite_diff_jac
// (alg_utils.jl:17-27: dx = sqrt(eps(T)), column i = (f(x + dx e_i) - f(x)) / dx); the time
// gradient likewise (ForwardDiff.derivative, or (f(t + dt) - f(t)) / dt with dt = sqrt(eps(T))).
//
// Forward mode needs no lowering work: the model bodies are templates over the scalar type T,
// so instantiating them with Dual<T, NP> differentiates them.  The arithmetic follows
// ForwardDiff's dual.jl (ForwardDiff is not vendored in the reference tree; formulas restated
// from the published package):
//   x*y : value vx*vy, partials muladd(vy, px, vx*py)
//   x/y : value vx/vy, partials muladd(inv(vy), px, (-(vx/(vy*vy)))*py)
//   f(x): value f(vx), partials f'(vx)*px   (DiffRules: sin->cos, cos->-sin, exp->exp, log->inv,
//         sqrt->inv(2 sqrt))
// Parameters and time enter as duals with zero partials.
#pragma once
#include "degk_common.cuh"

namespace degk {

enum { JAC_ANALYTIC = 0, JAC_FINITE_DIFF = 1, JAC_FORWARD_AD = 2 };

template <class T, int NP>
struct Dual {
    T v;
    T d[NP];
    DEGK_DEV Dual() {}
    DEGK_DEV Dual(T x) : v(x) { DEGK_UNROLL for (int i = 0; i < NP; ++i) d[i] = (T)0; }
    template <class U> DEGK_DEV Dual(U x) : v((T)x) { DEGK_UNROLL for (int i = 0; i < NP; ++i) d[i] = (T)0; }

    friend DEGK_DEV Dual operator+(const Dual& a, const Dual& b) {
        Dual r; r.v = a.v + b.v; DEGK_UNROLL for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] + b.d[i]; return r;
    }
    friend DEGK_DEV Dual operator-(const Dual& a, const Dual& b) {
        Dual r; r.v = a.v - b.v; DEGK_UNROLL for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] - b.d[i]; return r;
    }
    friend DEGK_DEV Dual operator-(const Dual& a) {
        Dual r; r.v = -a.v; DEGK_UNROLL for (int i = 0; i < NP; ++i) r.d[i] = -a.d[i]; return r;
    }
    friend DEGK_DEV Dual operator+(const Dual& a) { return a; }
    friend DEGK_DEV Dual operator*(const Dual& a, const Dual& b) {
        Dual r; r.v = a.v * b.v;
        DEGK_UNROLL for (int i = 0; i < NP; ++i) r.d[i] = fma_(b.v, a.d[i], a.v * b.d[i]);
        return r;
    }
    friend DEGK_DEV Dual operator/(const Dual& a, const Dual& b) {
        Dual r; r.v = a.v / b.v;
        const T ib = (T)1 / b.v, c = -(a.v / (b.v * b.v));
        DEGK_UNROLL for (int i = 0; i < NP; ++i) r.d[i] = fma_(ib, a.d[i], c * b.d[i]);
        return r;
    }
    DEGK_DEV Dual& operator+=(const Dual& b) { *this = *this + b; return *this; }
    DEGK_DEV Dual& operator-=(const Dual& b) { *this = *this - b; return *this; }
    DEGK_DEV Dual& operator*=(const Dual& b) { *this = *this * b; return *this; }
    DEGK_DEV Dual& operator/=(const Dual& b) { *this = *this / b; return *this; }

    static DEGK_DEV Dual chain(T val, T deriv, const Dual& x) {
        Dual r; r.v = val; DEGK_UNROLL for (int i = 0; i < NP; ++i) r.d[i] = deriv * x.d[i]; return r;
    }
    friend DEGK_DEV Dual sin(const Dual& x) { return chain(::sin(x.v), ::cos(x.v), x); }
    friend DEGK_DEV Dual cos(const Dual& x) { return chain(::cos(x.v), -::sin(x.v), x); }
    friend DEGK_DEV Dual exp(const Dual& x) { const T e = ::exp(x.v); return chain(e, e, x); }
    friend DEGK_DEV Dual log(const Dual& x) { return chain(::log(x.v), (T)1 / x.v, x); }
    friend DEGK_DEV Dual sqrt(const Dual& x) { const T s = ::sqrt(x.v); return chain(s, (T)1 / ((T)2 * s), x); }
    friend DEGK_DEV bool operator<(const Dual& a, const Dual& b) { return a.v < b.v; }
    friend DEGK_DEV bool operator>(const Dual& a, const Dual& b) { return a.v > b.v; }
    friend DEGK_DEV bool operator<=(const Dual& a, const Dual& b) { return a.v <= b.v; }
    friend DEGK_DEV bool operator>=(const Dual& a, const Dual& b) { return a.v >= b.v; }
};

// which mode a model asks for when it has no analytic jac/tgrad (JIT models carry JAC_MODE)
template <class...> struct void_t_ { typedef void type; };
template <class M, class = void> struct jac_mode_of { static constexpr int value = JAC_FORWARD_AD; };
template <class M> struct jac_mode_of<M, typename void_t_<decltype(M::JAC_MODE)>::type> { static constexpr int value = M::JAC_MODE; };

// constant mass matrix (reference: ODEFunction(f; mass_matrix = M), src/utils.jl:42-57): models that
// have one define HAS_MASS = true and mass(Mm); everything else is the identity (UniformScaling)
template <class M, class = void> struct has_mass_of { static constexpr bool value = false; };
template <class M> struct has_mass_of<M, typename void_t_<decltype(M::HAS_MASS)>::type> { static constexpr bool value = M::HAS_MASS; };
// models whose time gradient is identically zero (autonomous right-hand side with a user Jacobian) say
// TGRAD_ZERO = true: the fast build then leaves the `dt * d_i * dT` terms of the Rosenbrock stages out (adding an exact
// zero changes nothing but the sign of a zero); the strict build keeps every term the reference has
template <class M, class = void> struct tgrad_zero_of { static constexpr bool value = false; };
template <class M> struct tgrad_zero_of<M, typename void_t_<decltype(M::TGRAD_ZERO)>::type> { static constexpr bool value = M::TGRAD_ZERO && !DEGK_STRICT; };
// StaticArrays SMatrix * SVector: row sums as a left fold of the products
template <class T, int N>
DEGK_DEV void mass_mul(const T (&Mm)[N][N], const T (&v)[N], T (&out)[N]) {
    DEGK_UNROLL for (int i = 0; i < N; ++i) {
        T s = Mm[i][0] * v[0];
        DEGK_UNROLL for (int j = 1; j < N; ++j) s = s + Mm[i][j] * v[j];
        out[i] = s;
    }
}

template <class T> DEGK_DEV T sqrt_eps_();
template <> DEGK_DEV float sqrt_eps_<float>() { return 3.4526698300124393e-4f; }      // sqrt(eps(Float32))
template <> DEGK_DEV double sqrt_eps_<double>() { return 1.4901161193847656e-8; }     // sqrt(eps(Float64))

template <class T, class Model>
DEGK_DEV void eval_jac(T (&J)[Model::N][Model::N], const T (&u)[Model::N], const T* p, T t) {
    constexpr int N = Model::N;
    if constexpr (Model::HAS_JAC) {
        Model::template jac<T>(J, u, p, t);
    } else if constexpr (jac_mode_of<Model>::value == JAC_FINITE_DIFF) {
        // finite_diff_jac, alg_utils.jl:17-27
        const T dx = sqrt_eps_<T>();
        T f0[N];
        Model::template f<T>(f0, u, p, t);
        DEGK_UNROLL for (int i = 0; i < N; ++i) {
            T x[N], f1[N];
            DEGK_UNROLL for (int c = 0; c < N; ++c) x[c] = u[c];
            x[i] = x[i] + dx;
            Model::template f<T>(f1, x, p, t);
            DEGK_UNROLL for (int r = 0; r < N; ++r) J[r][i] = (f1[r] - f0[r]) / dx;
        }
    } else {
        // ForwardDiff.jacobian(u -> f(u, p, t), u): one dual per state with a unit partial
        typedef Dual<T, N> D;
        D ud[N], du[N], pd[Model::NP > 0 ? Model::NP : 1];
        DEGK_UNROLL for (int i = 0; i < N; ++i) { ud[i] = D(u[i]); ud[i].d[i] = (T)1; }
        DEGK_UNROLL for (int i = 0; i < Model::NP; ++i) pd[i] = D(p[i]);
        Model::template f<D>(du, ud, pd, D(t));
        DEGK_UNROLL for (int r = 0; r < N; ++r) DEGK_UNROLL for (int c = 0; c < N; ++c) J[r][c] = du[r].d[c];
    }
}

template <class T, class Model>
DEGK_DEV void eval_tgrad(T (&dT)[Model::N], const T (&u)[Model::N], const T* p, T t) {
    constexpr int N = Model::N;
    if constexpr (Model::HAS_TGRAD) {
        Model::template tgrad<T>(dT, u, p, t);
    } else if constexpr (jac_mode_of<Model>::value == JAC_FINITE_DIFF) {
        // build_tgrad, nlsolve/type.jl:149-153
        const T dt = sqrt_eps_<T>();
        T f0[N], f1[N];
        Model::template f<T>(f1, u, p, t + dt);
        Model::template f<T>(f0, u, p, t);
        DEGK_UNROLL for (int r = 0; r < N; ++r) dT[r] = (f1[r] - f0[r]) / dt;
    } else {
        // ForwardDiff.derivative(t -> f(u, p, t), t)
        typedef Dual<T, 1> D;
        D ud[N], du[N], pd[Model::NP > 0 ? Model::NP : 1];
        DEGK_UNROLL for (int i = 0; i < N; ++i) ud[i] = D(u[i]);
        DEGK_UNROLL for (int i = 0; i < Model::NP; ++i) pd[i] = D(p[i]);
        D td(t); td.d[0] = (T)1;
        Model::template f<D>(du, ud, pd, td);
        DEGK_UNROLL for (int r = 0; r < N; ++r) dT[r] = du[r].d[0];
    }
}

}  // namespace degk
