// degk_rosenbrock.cuh -- stiff steppers with a register-resident factorisation of W.
//
//   Rosenbrock23  reference perform_step/gpu_rosenbrock23_perform_step.jl:1-70, 74-201
//   Rodas4        reference perform_step/gpu_rodas4_perform_step.jl:1-116, 118-277
//   Rodas5P       reference perform_step/gpu_rodas5P_perform_step.jl:1-153, 155-351
//   LinSolve      reference linalg/linsolve.jl:8-57 (n<=3 adjugate), :88-108 + linalg/lu.jl:114-159
//
// The reference calls linear_solve(W, b) 2-8 times per attempt with the same W and relies on
// LLVM to CSE the cofactors / refactorise.  Here W is factored ONCE per attempt into registers
// (cofactors + det for n<=3, partial-pivot LU for n>=4) and each stage only does the
// substitution.  The values are identical to recomputing (same expressions), so the strict
// build stays bit-equal to the oracle.
// A constant mass matrix (Model::HAS_MASS) enters W and the stage right-hand sides; models without one
// keep the identity forms.
#pragma once
#include "degk_common.cuh"
#include "degk_pack.cuh"
#include "degk_dual.cuh"
#include "gen_rodas_consts.cuh"

namespace degk {

// ---------------------------------------------------------------------------------------
// n >= 4 uses loops that are fully unrolled (register-resident LU) up to n = 8; larger systems keep
// rolled loops (local memory) so that compile time stays bounded.
#define DEGK_UNROLL_LU _Pragma("unroll (N <= 8 ? N : 1)")

// NEGINV (fast build, n = 2 or 3 only; ignored otherwise): factor() stores -inverse(A) = cofactors * (-1 / det), so
// that each solve_neg() is the bare matrix-vector product -- for a stepper that solves 6-8 times per factorisation
// (Rodas4 / Rodas5P) this trades N*N multiplies per factorisation for N per solve.
#ifndef DEGK_ROS_NEGINV
#define DEGK_ROS_NEGINV 1
#endif
template <class T, int N, bool NEGINV_ = false>
struct LinSolve {
    static constexpr bool NEGINV = NEGINV_ && !DEGK_STRICT && DEGK_ROS_NEGINV && (N == 2 || N == 3);
    T lu[N][N];          // n>=4: L (unit, below diag) and U;  n<=3: cofactor matrix (NEGINV: -inverse)
    T dinv[N];           // n>=4: 1/U_jj ; n<=3: dinv[0] = det (strict) or 1/det (fast)
    T ndinv;             // n<=3: -dinv[0] (solve_neg)
    int piv[N];          // n>=4: row interchanged with row k at elimination step k

    DEGK_DEV bool factor(const T (&A)[N][N]) {
        if constexpr (N == 1) {
            dinv[0] = (T)1 / A[0][0];
            ndinv = -dinv[0];
            return true;
        } else if constexpr (N == 2) {
            const T d = A[0][0] * A[1][1] - A[0][1] * A[1][0];
#if DEGK_STRICT
            dinv[0] = d;
#else
            dinv[0] = (T)1 / d;
#endif
            ndinv = -dinv[0];
            lu[0][0] = A[1][1]; lu[0][1] = A[0][1]; lu[1][0] = A[1][0]; lu[1][1] = A[0][0];
            if constexpr (NEGINV) {
                const T nd = (T)-1 / d;
                lu[0][0] = lu[0][0] * nd; lu[0][1] = lu[0][1] * nd; lu[1][0] = lu[1][0] * nd; lu[1][1] = lu[1][1] * nd;
            }
            return true;
        } else if constexpr (N == 3) {
            const T a11 = A[0][0], a12 = A[0][1], a13 = A[0][2];
            const T a21 = A[1][0], a22 = A[1][1], a23 = A[1][2];
            const T a31 = A[2][0], a32 = A[2][1], a33 = A[2][2];
            lu[0][0] = a22 * a33 - a23 * a32; lu[0][1] = a13 * a32 - a12 * a33; lu[0][2] = a12 * a23 - a13 * a22;
            lu[1][0] = a23 * a31 - a21 * a33; lu[1][1] = a11 * a33 - a13 * a31; lu[1][2] = a13 * a21 - a11 * a23;
            lu[2][0] = a21 * a32 - a22 * a31; lu[2][1] = a12 * a31 - a11 * a32; lu[2][2] = a11 * a22 - a12 * a21;
            // det = x0 . (x1 x x2) over columns (StaticArrays): cross terms equal the first
            // cofactor column above (products commute), so reuse them.
            const T d = (a11 * lu[0][0] + a21 * lu[0][1]) + a31 * lu[0][2];
#if DEGK_STRICT
            dinv[0] = d;
#else
            dinv[0] = (T)1 / d;
#endif
            ndinv = -dinv[0];
            if constexpr (NEGINV) {
                const T nd = (T)-1 / d;
                DEGK_UNROLL for (int i = 0; i < 3; ++i)
                    DEGK_UNROLL for (int j = 0; j < 3; ++j) lu[i][j] = lu[i][j] * nd;
            }
            return true;
        } else {
            DEGK_UNROLL_LU for (int i = 0; i < N; ++i)
                DEGK_UNROLL_LU for (int j = 0; j < N; ++j) lu[i][j] = A[i][j];
            bool ok = true;
            DEGK_UNROLL_LU for (int k = 0; k < N; ++k) {
                int kp = k;
                T amax = abs_(lu[k][k]);
                DEGK_UNROLL_LU for (int i = k + 1; i < N; ++i) {
                    const T v = abs_(lu[i][k]);
                    if (v > amax) { kp = i; amax = v; }
                }
                piv[k] = kp;
                DEGK_UNROLL_LU for (int i = k + 1; i < N; ++i) {
                    if (kp == i) {
                        DEGK_UNROLL_LU for (int j = 0; j < N; ++j) { const T s = lu[k][j]; lu[k][j] = lu[i][j]; lu[i][j] = s; }
                    }
                }
                const T inv = (T)1 / lu[k][k];
                const bool fin = finite_(inv);
                DEGK_UNROLL_LU for (int i = k + 1; i < N; ++i) {
                    const T l = fin ? lu[i][k] * inv : (T)0;
                    lu[i][k] = l;
                    DEGK_UNROLL_LU for (int j = k + 1; j < N; ++j) lu[i][j] = lu[i][j] - l * lu[k][j];
                }
            }
            DEGK_UNROLL_LU for (int j = 0; j < N; ++j) {
                if (lu[j][j] == (T)0) ok = false;
                dinv[j] = (T)1 / lu[j][j];
            }
            return ok;
        }
    }

    // x = A \ (-b), bit for bit what solve() returns for the negated right-hand side (products, sums, quotients are
    // sign-symmetric under round-to-nearest), without negating b: for the closed forms the sign rides on the
    // determinant factor -- one negation per solve instead of one per component, and for packed pairs (whose negation
    // is two integer XORs the FMA pipe cannot fold) none at all in the stage arithmetic.
    DEGK_DEV void solve_neg(const T (&b)[N], T (&x)[N]) const {
        if constexpr (NEGINV && N == 2) {
            x[0] = lu[0][0] * b[0] - lu[0][1] * b[1];
            x[1] = lu[1][1] * b[1] - lu[1][0] * b[0];
        } else if constexpr (NEGINV && N == 3) {
            DEGK_UNROLL for (int i = 0; i < 3; ++i) x[i] = (lu[i][0] * b[0] + lu[i][1] * b[1]) + lu[i][2] * b[2];
        } else if constexpr (N == 1) {
            x[0] = ndinv * b[0];
        } else if constexpr (N == 2) {
            const T nd = ndinv;
#if DEGK_STRICT
            x[0] = (lu[0][0] * b[0] - lu[0][1] * b[1]) / nd;
            x[1] = (lu[1][1] * b[1] - lu[1][0] * b[0]) / nd;
#else
            x[0] = (lu[0][0] * b[0] - lu[0][1] * b[1]) * nd;
            x[1] = (lu[1][1] * b[1] - lu[1][0] * b[0]) * nd;
#endif
        } else if constexpr (N == 3) {
            const T nd = ndinv;
            DEGK_UNROLL for (int i = 0; i < 3; ++i) {
                const T s = (lu[i][0] * b[0] + lu[i][1] * b[1]) + lu[i][2] * b[2];
#if DEGK_STRICT
                x[i] = s / nd;
#else
                x[i] = s * nd;
#endif
            }
        } else {
            solve(b, x);
            DEGK_UNROLL_LU for (int i = 0; i < N; ++i) x[i] = -x[i];
        }
    }

    DEGK_DEV void solve(const T (&b)[N], T (&x)[N]) const {
        static_assert(!NEGINV, "a NEGINV factorisation only serves solve_neg()");
        if constexpr (N == 1) {
            x[0] = dinv[0] * b[0];
        } else if constexpr (N == 2) {
#if DEGK_STRICT
            x[0] = (lu[0][0] * b[0] - lu[0][1] * b[1]) / dinv[0];
            x[1] = (lu[1][1] * b[1] - lu[1][0] * b[0]) / dinv[0];
#else
            x[0] = (lu[0][0] * b[0] - lu[0][1] * b[1]) * dinv[0];
            x[1] = (lu[1][1] * b[1] - lu[1][0] * b[0]) * dinv[0];
#endif
        } else if constexpr (N == 3) {
            DEGK_UNROLL for (int i = 0; i < 3; ++i) {
                const T s = (lu[i][0] * b[0] + lu[i][1] * b[1]) + lu[i][2] * b[2];
#if DEGK_STRICT
                x[i] = s / dinv[0];
#else
                x[i] = s * dinv[0];
#endif
            }
        } else {
            T y[N];
            DEGK_UNROLL_LU for (int i = 0; i < N; ++i) y[i] = b[i];
            DEGK_UNROLL_LU for (int k = 0; k < N; ++k) {
                DEGK_UNROLL_LU for (int i = k + 1; i < N; ++i) {
                    if (piv[k] == i) { const T s = y[k]; y[k] = y[i]; y[i] = s; }
                }
            }
            DEGK_UNROLL_LU for (int j = 0; j < N; ++j)
                DEGK_UNROLL_LU for (int i = j + 1; i < N; ++i) y[i] = y[i] - lu[i][j] * y[j];
            DEGK_UNROLL_LU for (int j = N - 1; j >= 0; --j) {
                y[j] = dinv[j] * y[j];
                DEGK_UNROLL_LU for (int i = j - 1; i >= 0; --i) y[i] = y[i] - lu[i][j] * y[j];
            }
            DEGK_UNROLL_LU for (int i = 0; i < N; ++i) x[i] = y[i];
        }
    }
};

// ---------------------------------------------------------------------------------------
template <class T, class Model>
struct Rosenbrock23 {
    static constexpr int N = Model::N;
    static constexpr int ORDER = 2;
    static constexpr bool FSAL = false;
    static constexpr bool ALWAYS_SOLVED = false;  // attempt() returns false when W is singular
    struct Keep { T k1[N], k2[N]; };
    static DEGK_DEV T dtmin() { return (T)1.0e-14f; }     // convert(T, 1.0f-14), SURVEY Q13
    static DEGK_DEV T land()  { return (T)1.0e-14f; }
    static DEGK_DEV void init(Keep&, const T (&)[N], const T*, T) {}
    static DEGK_DEV void accepted(Keep&) {}
    static DEGK_DEV void on_accept(Keep&) {}
    static DEGK_DEV void init_sel(Keep&, const T (&)[N], const T*, T, unsigned) {}
    static DEGK_DEV void accepted_sel(Keep&, unsigned) {}
    static DEGK_DEV void accepted_if(Keep&, const bool*) {}

    template <bool WANT_ERR>
    static DEGK_DEV bool attempt(Keep& K, const T (&uprev)[N], const T* p, T t, T h,
                                 T (&unew)[N], T (&err)[N]) {
        const T two = (T)2;
        const T d = (T)1 / (two + sqrt_(two));            // stiff/types.jl:47-48
        const T gam = h * d;
        const T dto2 = h / (T)2, dto6 = h / (T)6;
        constexpr bool TG = !tgrad_zero_of<Model>::value;    // false: dT == 0 by declaration, its terms are left out
        T J[N][N], W[N][N], dT[N];
        eval_jac<T, Model>(J, uprev, p, t);      // analytic, ForwardDiff-style duals or finite differences
        if constexpr (TG) eval_tgrad<T, Model>(dT, uprev, p, t);
        constexpr bool MASS = has_mass_of<Model>::value;     // W = mass_matrix - gamma*J (:45, :120)
        const T ngam = -gam;
        T Mm[MASS ? N : 1][MASS ? N : 1];
        if constexpr (MASS) {
            Model::template mass<T>(Mm);
            DEGK_UNROLL for (int i = 0; i < N; ++i)
                DEGK_UNROLL for (int j = 0; j < N; ++j) W[i][j] = Mm[i][j] - gam * J[i][j];
        } else {
            DEGK_UNROLL for (int i = 0; i < N; ++i)
                DEGK_UNROLL for (int j = 0; j < N; ++j) {
                    const T v = ngam * J[i][j];              // -(gam * J): the sign rides on the scalar
                    W[i][j] = (i == j) ? v + (T)1 : v;
                }
        }
        LinSolve<T, N> F;
        if (!F.factor(W)) return false;
        T F0[N], F1[N], rhs[N], tmp[N];
        Model::template f<T>(F0, uprev, p, t);
        DEGK_UNROLL for (int c = 0; c < N; ++c) rhs[c] = TG ? F0[c] + gam * dT[c] : F0[c];
        F.solve(rhs, K.k1);
        DEGK_UNROLL for (int c = 0; c < N; ++c) tmp[c] = uprev[c] + dto2 * K.k1[c];
        Model::template f<T>(F1, tmp, p, t + dto2);
        if constexpr (MASS) {                                // F1 - mass_matrix * k1 (:57, :132)
            T mk[N];
            mass_mul<T, N>(Mm, K.k1, mk);
            DEGK_UNROLL for (int c = 0; c < N; ++c) rhs[c] = F1[c] - mk[c];
        } else {
            DEGK_UNROLL for (int c = 0; c < N; ++c) rhs[c] = F1[c] - K.k1[c];
        }
        F.solve(rhs, K.k2);
        DEGK_UNROLL for (int c = 0; c < N; ++c) K.k2[c] = K.k2[c] + K.k1[c];
        DEGK_UNROLL for (int c = 0; c < N; ++c) unew[c] = uprev[c] + h * K.k2[c];
        if (WANT_ERR) {
            const T e32 = (T)6 + sqrt_((T)2);
            T F2[N], k3[N];
            Model::template f<T>(F2, unew, p, t + h);
            if constexpr (MASS) {
                // F2 - mass_matrix * (e32*k2 + 2*k1) + e32*F1 + 2*F0 + dt*dT  (:144-148)
                T v[N], mv[N];
                DEGK_UNROLL for (int c = 0; c < N; ++c) v[c] = e32 * K.k2[c] + two * K.k1[c];
                mass_mul<T, N>(Mm, v, mv);
                DEGK_UNROLL for (int c = 0; c < N; ++c)
                    { const T r_ = ((F2[c] - mv[c]) + e32 * F1[c]) + two * F0[c]; rhs[c] = TG ? r_ + h * dT[c] : r_; }
            } else {
                DEGK_UNROLL for (int c = 0; c < N; ++c)
                    { const T r_ = (F2[c] - e32 * (K.k2[c] - F1[c])) - two * (K.k1[c] - F0[c]); rhs[c] = TG ? r_ + h * dT[c] : r_; }
            }
            F.solve(rhs, k3);
            DEGK_UNROLL for (int c = 0; c < N; ++c) err[c] = dto6 * ((K.k1[c] - two * K.k2[c]) + k3[c]);
        }
        return true;
    }

    // nonstiff/interpolants.jl:389-399
    static DEGK_DEV void interp(const Keep& K, T theta, T h, const T (&uprev)[N],
                                const T (&unew)[N], const T* p, T tprev, T (&out)[N]) {
        const T two = (T)2;
        const T d = (T)1 / (two + sqrt_(two));
        const T den = fma_((T)-2, d, (T)1);
        const T c1 = theta * ((T)1 - theta) / den;
        const T c2 = theta * fma_((T)-2, d, theta) / den;
        DEGK_UNROLL for (int c = 0; c < N; ++c)
            out[c] = fma_(h, fma_(c2, K.k2[c], c1 * K.k1[c]), uprev[c]);
    }
};

// ---------------------------------------------------------------------------------------
// Rodas4 (6 stages) and Rodas5P (8 stages) share the structure; R5 selects the tableau.
template <class T, class Model, bool R5>
struct Rodas {
    static constexpr int N = Model::N;
    static constexpr int ORDER = R5 ? 5 : 4;
    static constexpr bool FSAL = false;
    static constexpr bool ALWAYS_SOLVED = false;  // attempt() returns false when W is singular
    static constexpr int NS = R5 ? 8 : 6;
    struct Keep { T ks[NS][N]; T kk[3][N]; };
    static DEGK_DEV T dtmin() { return (T)1.0e-14f; }
    static DEGK_DEV T land()  { return (T)1.0e-14f; }
    static DEGK_DEV void init(Keep&, const T (&)[N], const T*, T) {}
    static DEGK_DEV void accepted(Keep&) {}
    static DEGK_DEV void init_sel(Keep&, const T (&)[N], const T*, T, unsigned) {}
    static DEGK_DEV void accepted_sel(Keep&, unsigned) {}
    static DEGK_DEV void accepted_if(Keep&, const bool*) {}

#define R4C(x) ((T)rodas4c::x)
#define R5C(x) ((T)rodas5pc::x)
#define RC(x) (R5 ? R5C(x) : R4C(x))

    template <bool WANT_ERR>
    static DEGK_DEV bool attempt(Keep& K, const T (&uprev)[N], const T* p, T t, T h,
                                 T (&unew)[N], T (&err)[N]) {
        constexpr bool TG = !tgrad_zero_of<Model>::value;    // false: dT == 0 by declaration, its terms are left out
        T J[N][N], dT[TG ? N : 1];
        eval_jac<T, Model>(J, uprev, p, t);      // analytic, ForwardDiff-style duals or finite differences
        if constexpr (TG) eval_tgrad<T, Model>(dT, uprev, p, t);
#define DTD_(x, dtd, c) (TG ? (x) + (dtd) * dT[TG ? c : 0] : (x))
        // The stage right-hand sides hold sum_j (C_ij / dt) k_j.  Strict: dtC_ij = C_ij / dt first, then the products,
        // as the reference writes it.  Fast: the sum with the C_ij as immediates, one multiply by 1/dt per component --
        // seven packed multiplies fewer per attempt, and 84 FFMA2 read an immediate instead of a third register pair
        // (2.08 instead of 3.06 issue cycles each, profiles/r2_pipe_probe2.jsonl).
#ifndef DEGK_ROS_HINV
#define DEGK_ROS_HINV 1
#endif
#if DEGK_STRICT || !DEGK_ROS_HINV
#define CH_(v) ((v) / h)
#define CSC_(v) (v)
#else
        const T hinv = (T)1 / h;
#define CH_(v) (v)
#define CSC_(v) (hinv * (v))
#endif
#if DEGK_STRICT || !DEGK_ROS_HINV
        const T dtgamma = h * RC(gamma);
        const T invdg = (T)1 / dtgamma;
#else
        const T invdg = hinv * (T)(1.0 / (R5 ? (double)rodas5pc::gamma : (double)rodas4c::gamma));
#endif
        // W = J - mass_matrix * inv(dtgamma) (gpu_rodas5P_perform_step.jl:81-82, 238-239); every
        // `mass_matrix * (dtC.. * k..)` below goes through MASSV (identity: nothing happens)
        constexpr bool MASS = has_mass_of<Model>::value;
        T Mm[MASS ? N : 1][MASS ? N : 1];
        if constexpr (MASS) {
            Model::template mass<T>(Mm);
            DEGK_UNROLL for (int i = 0; i < N; ++i)
                DEGK_UNROLL for (int j = 0; j < N; ++j) J[i][j] = J[i][j] - Mm[i][j] * invdg;
        } else {
            DEGK_UNROLL for (int i = 0; i < N; ++i) J[i][i] = J[i][i] - invdg;
        }
#define MASSV(v) do { if constexpr (MASS) { T mv_[N]; mass_mul<T, N>(Mm, v, mv_); DEGK_UNROLL for (int c_ = 0; c_ < N; ++c_) v[c_] = mv_[c_]; } } while (0)
        T cs[N];
        LinSolve<T, N, true> F;
        if (!F.factor(J)) return false;
        T (&k)[NS][N] = K.ks;
        T du[N], lt[N], uu[N];
        Model::template f<T>(du, uprev, p, t);
        // Step 1
        { const T dtd1 = h * RC(d1);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = DTD_(du[c], dtd1, c); }
        F.solve_neg(lt, k[0]);
        DEGK_UNROLL for (int c = 0; c < N; ++c) uu[c] = uprev[c] + RC(a21) * k[0][c];
        Model::template f<T>(du, uu, p, t + RC(c2) * h);
        // Step 2
        { const T dtd2 = h * RC(d2), C21 = CH_(RC(C21));
          DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_(C21 * k[0][c]);
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = DTD_(du[c], dtd2, c) + cs[c]; }
        F.solve_neg(lt, k[1]);
        DEGK_UNROLL for (int c = 0; c < N; ++c) uu[c] = (uprev[c] + RC(a31) * k[0][c]) + RC(a32) * k[1][c];
        Model::template f<T>(du, uu, p, t + RC(c3) * h);
        // Step 3
        { const T dtd3 = h * RC(d3), C31 = CH_(RC(C31)), C32 = CH_(RC(C32));
          DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_((C31 * k[0][c] + C32 * k[1][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = DTD_(du[c], dtd3, c) + cs[c]; }
        F.solve_neg(lt, k[2]);
        DEGK_UNROLL for (int c = 0; c < N; ++c)
            uu[c] = ((uprev[c] + RC(a41) * k[0][c]) + RC(a42) * k[1][c]) + RC(a43) * k[2][c];
        Model::template f<T>(du, uu, p, t + RC(c4) * h);
        // Step 4
        { const T dtd4 = h * RC(d4), C41 = CH_(RC(C41)), C42 = CH_(RC(C42)), C43 = CH_(RC(C43));
          DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_(((C41 * k[0][c] + C42 * k[1][c]) + C43 * k[2][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = DTD_(du[c], dtd4, c) + cs[c]; }
        F.solve_neg(lt, k[3]);
        DEGK_UNROLL for (int c = 0; c < N; ++c)
            uu[c] = (((uprev[c] + RC(a51) * k[0][c]) + RC(a52) * k[1][c]) + RC(a53) * k[2][c]) + RC(a54) * k[3][c];
        const T C51 = CH_(RC(C51)), C52 = CH_(RC(C52)), C53 = CH_(RC(C53)), C54 = CH_(RC(C54));
        if (R5) {
            Model::template f<T>(du, uu, p, t + R5C(c5) * h);
            // Step 5: summands in the order k2,k4,k1,k3 (gpu_rodas5P_perform_step.jl:271)
            { const T dtd5 = h * R5C(d5);
              DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_((((C52 * k[1][c] + C54 * k[3][c]) + C51 * k[0][c]) + C53 * k[2][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = DTD_(du[c], dtd5, c) + cs[c]; }
            F.solve_neg(lt, k[4]);
            DEGK_UNROLL for (int c = 0; c < N; ++c)
                uu[c] = ((((uprev[c] + R5C(a61) * k[0][c]) + R5C(a62) * k[1][c]) + R5C(a63) * k[2][c]) +
                         R5C(a64) * k[3][c]) + R5C(a65) * k[4][c];
            Model::template f<T>(du, uu, p, t + h);
            // Step 6
            { const T C61 = CH_(R5C(C61)), C62 = CH_(R5C(C62)), C63 = CH_(R5C(C63)), C64 = CH_(R5C(C64)), C65 = CH_(R5C(C65));
              DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_(((((C61 * k[0][c] + C62 * k[1][c]) + C63 * k[2][c]) + C64 * k[3][c]) + C65 * k[4][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = du[c] + cs[c]; }
            F.solve_neg(lt, k[5]);
            DEGK_UNROLL for (int c = 0; c < N; ++c) uu[c] = uu[c] + k[5][c];
            Model::template f<T>(du, uu, p, t + h);
            // Step 7
            { const T C71 = CH_(R5C(C71)), C72 = CH_(R5C(C72)), C73 = CH_(R5C(C73)), C74 = CH_(R5C(C74)),
                      C75 = CH_(R5C(C75)), C76 = CH_(R5C(C76));
              DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_((((((C71 * k[0][c] + C72 * k[1][c]) + C73 * k[2][c]) + C74 * k[3][c]) +
                                      C75 * k[4][c]) + C76 * k[5][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = du[c] + cs[c]; }
            F.solve_neg(lt, k[NS > 6 ? 6 : 0]);
            DEGK_UNROLL for (int c = 0; c < N; ++c) uu[c] = uu[c] + k[NS > 6 ? 6 : 0][c];
            Model::template f<T>(du, uu, p, t + h);
            // Step 8
            { const T C81 = CH_(R5C(C81)), C82 = CH_(R5C(C82)), C83 = CH_(R5C(C83)), C84 = CH_(R5C(C84)),
                      C85 = CH_(R5C(C85)), C86 = CH_(R5C(C86)), C87 = CH_(R5C(C87));
              DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_(((((((C81 * k[0][c] + C82 * k[1][c]) + C83 * k[2][c]) + C84 * k[3][c]) +
                                       C85 * k[4][c]) + C86 * k[5][c]) + C87 * k[NS > 6 ? 6 : 0][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = du[c] + cs[c]; }
            F.solve_neg(lt, k[NS - 1]);
            DEGK_UNROLL for (int c = 0; c < N; ++c) unew[c] = uu[c] + k[NS - 1][c];
        } else {
            Model::template f<T>(du, uu, p, t + h);
            // Step 5: summands in the order k2,k4,k1,k3 (gpu_rodas4_perform_step.jl:213)
            DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_((((C52 * k[1][c] + C54 * k[3][c]) + C51 * k[0][c]) + C53 * k[2][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = du[c] + cs[c];
            F.solve_neg(lt, k[4]);
            DEGK_UNROLL for (int c = 0; c < N; ++c) uu[c] = uu[c] + k[4][c];
            Model::template f<T>(du, uu, p, t + h);
            // Step 6: summands in the order k1,k2,k5,k4,k3 (:219)
            { const T C61 = CH_(R4C(C61)), C62 = CH_(R4C(C62)), C63 = CH_(R4C(C63)), C64 = CH_(R4C(C64)), C65 = CH_(R4C(C65));
              DEGK_UNROLL for (int c = 0; c < N; ++c) cs[c] = CSC_(((((C61 * k[0][c] + C62 * k[1][c]) + C65 * k[4][c]) + C64 * k[3][c]) + C63 * k[2][c]));
          MASSV(cs);
          DEGK_UNROLL for (int c = 0; c < N; ++c) lt[c] = du[c] + cs[c]; }
            F.solve_neg(lt, k[5]);
            DEGK_UNROLL for (int c = 0; c < N; ++c) unew[c] = uu[c] + k[5][c];
        }
#undef CH_
#undef CSC_
#undef DTD_
        if (WANT_ERR) {
            DEGK_UNROLL for (int c = 0; c < N; ++c) err[c] = k[NS - 1][c];
        }
#undef MASSV
        return true;
    }

    // interpolation vectors, formed on accept only (gpu_rodas4:245-247, gpu_rodas5P:313-323)
    static DEGK_DEV void on_accept(Keep& K) {
        T (&k)[NS][N] = K.ks;
        DEGK_UNROLL for (int c = 0; c < N; ++c) {
            if (R5) {
                K.kk[0][c] = ((((((R5C(h21) * k[0][c] + R5C(h22) * k[1][c]) + R5C(h23) * k[2][c]) + R5C(h24) * k[3][c]) +
                                R5C(h25) * k[4][c]) + R5C(h26) * k[5][c]) + R5C(h27) * k[NS > 6 ? 6 : 0][c]) + R5C(h28) * k[NS - 1][c];
                K.kk[1][c] = ((((((R5C(h31) * k[0][c] + R5C(h32) * k[1][c]) + R5C(h33) * k[2][c]) + R5C(h34) * k[3][c]) +
                                R5C(h35) * k[4][c]) + R5C(h36) * k[5][c]) + R5C(h37) * k[NS > 6 ? 6 : 0][c]) + R5C(h38) * k[NS - 1][c];
                K.kk[2][c] = ((((((R5C(h41) * k[0][c] + R5C(h42) * k[1][c]) + R5C(h43) * k[2][c]) + R5C(h44) * k[3][c]) +
                                R5C(h45) * k[4][c]) + R5C(h46) * k[5][c]) + R5C(h47) * k[NS > 6 ? 6 : 0][c]) + R5C(h48) * k[NS - 1][c];
            } else {
                K.kk[0][c] = (((R4C(h21) * k[0][c] + R4C(h22) * k[1][c]) + R4C(h23) * k[2][c]) + R4C(h24) * k[3][c]) + R4C(h25) * k[4][c];
                K.kk[1][c] = (((R4C(h31) * k[0][c] + R4C(h32) * k[1][c]) + R4C(h33) * k[2][c]) + R4C(h34) * k[3][c]) + R4C(h35) * k[4][c];
                K.kk[2][c] = (T)0;
            }
        }
    }

    // stiff/interpolants.jl:1-11 (Rodas4), :13-23 (Rodas5P)
    static DEGK_DEV void interp(const Keep& K, T theta, T h, const T (&uprev)[N],
                                const T (&unew)[N], const T* p, T tprev, T (&out)[N]) {
        const T th1 = (T)1 - theta;
        DEGK_UNROLL for (int c = 0; c < N; ++c) {
            const T inner = R5 ? fma_(theta, fma_(theta, K.kk[2][c], K.kk[1][c]), K.kk[0][c])
                               : fma_(theta, K.kk[1][c], K.kk[0][c]);
            out[c] = fma_(theta, fma_(th1, inner, unew[c]), th1 * uprev[c]);
        }
    }
#undef R4C
#undef R5C
#undef RC
};

}  // namespace degk
