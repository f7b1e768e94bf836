// =====================================================================================
// degk_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++ restatement of the per-trajectory algorithm of SciML/DiffEqGPU.jl's
// `EnsembleGPUKernel` path, used to *check* the CUDA engine (tests/, __graft_entry__.smoke(),
// bench.py's cpu_baseline / --impl reference leg).  Nothing in the product path
// (diffeqgpu.jl_b200/) includes, links, or calls this file.
//
// PARITY STATUS: "parity unpinned" for per-step arithmetic.  The reference has no golden
// vectors and cannot run here (no Julia).  What *is* pinned: the reference
// tests' own assertions (solution vs a high-accuracy truth within 5e-4..1e-2, output
// shapes, ts grids), reproduced in tests/ with scipy DOP853/Radau (rtol 1e-12) standing in
// for OrdinaryDiffEq.  Rounding-level behaviour that lives in un-vendored Julia deps
// (Float32 `^`, StaticArrays `det`/`sum`, MuladdMacro association) is restated from the
// published upstream algorithms and exposed as policy switches below.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src/ensemblegpukernel/).
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp).
// -ffp-contract=off is essential: the reference's stage arithmetic has no @muladd, so Julia
// emits separate fmul/fadd; only the interpolants are fused (explicit fma).
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <limits>
#include <vector>
#include <algorithm>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int MAXN = 16;   // max state dimension handled by the oracle
constexpr int MAXS = 27;   // max stage index (Vern9 extras go to 26)

// ---- return codes (shared numbering with include/degk.h) ----
enum { RC_DEFAULT = 0, RC_SUCCESS = 1, RC_DT_LESS_THAN_MIN = 2, RC_UNSTABLE = 3, RC_MAXITERS = 4,
       RC_SINGULAR = 5 };

struct Term { int j; double a; };
struct Row  { const Term* t; int n; double c; };
struct Poly { int stage; const double* r; int n; };
struct ErkTable {
    const char* name; int order; int stages; int fsal;
    const Row* rows; const Term* b; int nb; const Term* bt; int nbt;
    const Row* extra; int nextra; const Poly* interp; int ninterp;
    const int* kept; int nkept;
    double dtmin_lit; int dtmin_via_f32; double land_lit; int land_via_f32;
};
#include "oracle_tables.inc"

// ---- arithmetic policies ----------------------------------------------------------
struct Policy {
    int fma_stages;   // 0 = reference (separate mul/add); 1 = emulate a contracted GPU build
};

template <class T> inline T fmaT(T a, T b, T c);
template <> inline float  fmaT<float>(float a, float b, float c)   { return std::fmaf(a, b, c); }
template <> inline double fmaT<double>(double a, double b, double c) { return std::fma(a, b, c); }

// Julia Base `^` (base/math.jl): Float32 goes through Float64 exp2/log2; Float64 is its own
// <1ulp pow -- restated here with libm pow.  Call sites: gpu_tsit5_perform_step.jl:127-128.
inline float  jl_pow(float x, float y) {
    if (x == 1.0f) return 1.0f;
    return (float)std::exp2(std::log2((double)x) * (double)y);
}
inline double jl_pow(double x, double y) { return std::pow(x, y); }

// Julia max/min propagate NaN (C fmax/fmin do not).
template <class T> inline T jl_max(T a, T b) { return (a != a) ? a : (b != b) ? b : (a > b ? a : b); }
template <class T> inline T jl_min(T a, T b) { return (a != a) ? a : (b != b) ? b : (a < b ? a : b); }

// convert(T, 1.0f-14) vs T(1.0e-14)  
template <class T> inline T thresh(double lit, int via_f32) {
    return via_f32 ? (T)(float)lit : (T)lit;
}

// =====================================================================================
// Models (restated from the reference's tests / SURVEY §8d)
// =====================================================================================
enum ModelId { M_LORENZ = 0, M_HENON_HEILES = 1, M_ROBER = 2, M_DECAY = 3, M_LINEAR15 = 4,
               M_GBM = 5, M_LORENZ_ADDITIVE = 6, M_SCALAR_SDE = 7, M_OSC_T = 8, M_GBM_ND = 9,
               M_QUAD_DECAY = 10,
               M_ROBER_DAE = 11,
               M_BALL = 12,
               M_LIN_DAE = 13 };      // 2-state index-1 DAE, M = diag(1, 0): stiff_ode/gpu_ode_modelingtoolkit_dae.jl:22-44         // bouncing ball x'' = -g: test/gpu_kernel_de/gpu_ode_continuous_callbacks.jl:6-10    // Robertson DAE + mass matrix diag(1,1,0): stiff_ode/gpu_ode_mass_matrix.jl:5-31   // du = -p u^2: test/gpu_kernel_de/finite_diff.jl:6-9, forward_diff.jl

struct ModelInfo { int n, np, m; bool has_jac; bool diag_noise; };

inline ModelInfo model_info(int id) {
    switch (id) {
    case M_LORENZ:          return {3, 3, 0, true, true};
    case M_HENON_HEILES:    return {4, 0, 0, false, true};
    case M_ROBER:           return {3, 3, 0, true, true};
    case M_DECAY:           return {1, 1, 0, true, true};
    case M_LINEAR15:        return {15, 0, 0, true, true};
    case M_GBM:             return {3, 2, 3, false, true};   // du = p1 u dt + p2 u dW (diag)
    case M_LORENZ_ADDITIVE: return {3, 3, 3, false, true};
    case M_SCALAR_SDE:      return {1, 2, 1, false, true};
    case M_OSC_T:           return {2, 1, 0, true, true};    // non-autonomous forced oscillator
    case M_GBM_ND:          return {2, 2, 4, false, false};  // 2x4 non-diagonal noise
    case M_QUAD_DECAY:      return {1, 1, 0, true, true};
    case M_ROBER_DAE:       return {3, 3, 0, true, true};
    case M_BALL:            return {2, 1, 0, true, true};
    case M_LIN_DAE:         return {2, 2, 0, true, true};
    }
    return {0, 0, 0, false, true};
}

// Forward-mode dual numbers for the `ForwardDiff.jacobian` / `ForwardDiff.derivative` branch of
// nlsolve/type.jl:129-157.  ForwardDiff is not vendored; arithmetic restated from its dual.jl:
//   x*y: (vx*vy, muladd(vy, px, vx*py))     x/y: (vx/vy, muladd(inv(vy), px, (-(vx/(vy*vy)))*py))
//   f(x): (f(vx), f'(vx)*px) with DiffRules derivatives.  p and t enter with zero partials.
template <class T, int NP>
struct ODual {
    T v; T d[NP];
    ODual() {}
    ODual(T x) : v(x) { for (int i = 0; i < NP; ++i) d[i] = (T)0; }
    template <class U> ODual(U x) : v((T)x) { for (int i = 0; i < NP; ++i) d[i] = (T)0; }
    friend ODual operator+(const ODual& a, const ODual& b) { ODual r; r.v = a.v + b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
    friend ODual operator-(const ODual& a, const ODual& b) { ODual r; r.v = a.v - b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
    friend ODual operator-(const ODual& a) { ODual r; r.v = -a.v; for (int i = 0; i < NP; ++i) r.d[i] = -a.d[i]; return r; }
    friend ODual operator*(const ODual& a, const ODual& b) {
        ODual r; r.v = a.v * b.v;
        for (int i = 0; i < NP; ++i) r.d[i] = fmaT<T>(b.v, a.d[i], a.v * b.d[i]);
        return r;
    }
    friend ODual operator/(const ODual& a, const ODual& b) {
        ODual r; r.v = a.v / b.v;
        const T ib = (T)1 / b.v, c = -(a.v / (b.v * b.v));
        for (int i = 0; i < NP; ++i) r.d[i] = fmaT<T>(ib, a.d[i], c * b.d[i]);
        return r;
    }
    static ODual chain(T val, T deriv, const ODual& x) { ODual r; r.v = val; for (int i = 0; i < NP; ++i) r.d[i] = deriv * x.d[i]; return r; }
    friend ODual sin(const ODual& x) { return chain(std::sin(x.v), std::cos(x.v), x); }
    friend ODual cos(const ODual& x) { return chain(std::cos(x.v), -std::sin(x.v), x); }
    friend ODual exp(const ODual& x) { const T e = std::exp(x.v); return chain(e, e, x); }
};

// f(u,p,t).  lorenz: test/gpu_kernel_de/gpu_ode_regression.jl:4-12.
// rober (ODE form): test/gpu_kernel_de/stiff_ode/gpu_ode_mass_matrix.jl:5-13 with the third
// row replaced by k2*y2^2.  decay: stiff_ode/gpu_ode_regression.jl:5-8.
template <class T>
inline void model_f(int id, T* du, const T* u, const T* p, T t) {
    switch (id) {
    case M_LORENZ:
    case M_LORENZ_ADDITIVE:
        du[0] = p[0] * (u[1] - u[0]);
        du[1] = u[0] * (p[1] - u[2]) - u[1];
        du[2] = u[0] * u[1] - p[2] * u[2];
        break;
    case M_HENON_HEILES:   // u = (x, y, px, py)
        du[0] = u[2];
        du[1] = u[3];
        du[2] = -u[0] - (T)2 * u[0] * u[1];
        du[3] = -u[1] - (u[0] * u[0] - u[1] * u[1]);
        break;
    case M_ROBER:
        du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];
        du[1] = p[0] * u[0] - p[1] * (u[1] * u[1]) - p[2] * u[1] * u[2];
        du[2] = p[1] * (u[1] * u[1]);
        break;
    case M_DECAY:
        du[0] = -p[0] * u[0];
        break;
    case M_LINEAR15:       // stiff_ode/gpu_ode_regression.jl:24-26  f_large = 1.01 u
        for (int i = 0; i < 15; ++i) du[i] = (T)1.01 * u[i];
        break;
    case M_GBM:
        for (int i = 0; i < 3; ++i) du[i] = p[0] * u[i];
        break;
    case M_SCALAR_SDE:
        du[0] = p[0] * u[0];
        break;
    case M_OSC_T:          // x'' = -x + p cos(t)
        du[0] = u[1];
        { using std::cos; du[1] = -u[0] + p[0] * cos(t); }
        break;
    case M_GBM_ND:
        du[0] = p[0] * u[0];
        du[1] = p[0] * u[1];
        break;
    case M_QUAD_DECAY:
        du[0] = -p[0] * u[0] * u[0];
        break;
    case M_ROBER_DAE:
        du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];
        du[1] = p[0] * u[0] - p[1] * (u[1] * u[1]) - p[2] * u[1] * u[2];
        du[2] = u[0] + u[1] + u[2] - (T)1;
        break;
    case M_BALL:
        du[0] = u[1];
        du[1] = -p[0];
        break;
    case M_LIN_DAE:        // dae_f with the literals -0.04f0, 1.0f4 as parameters (sweepable)
        du[0] = -p[0] * u[0] + p[1] * u[1];
        du[1] = u[0] + u[1] - (T)1;
        break;
    }
}

// constant mass matrix of the model; returns false for the identity (UniformScaling)
template <class T>
inline bool model_mass(int id, T (*Mm)[MAXN]) {
    if (id != M_ROBER_DAE && id != M_LIN_DAE) return false;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Mm[i][j] = (T)0;
    Mm[0][0] = (T)1;
    if (id == M_ROBER_DAE) Mm[1][1] = (T)1;
    return true;
}

// analytic Jacobian J[i][j] = d f_i / d u_j  (nlsolve/type.jl:129-132, `f.jac` branch)
template <class T>
inline void model_jac(int id, T (*J)[MAXN], const T* u, const T* p, T t) {
    const int n = model_info(id).n;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) J[i][j] = (T)0;
    switch (id) {
    case M_LORENZ:
        J[0][0] = -p[0];        J[0][1] = p[0];
        J[1][0] = p[1] - u[2];  J[1][1] = (T)-1;  J[1][2] = -u[0];
        J[2][0] = u[1];         J[2][1] = u[0];   J[2][2] = -p[2];
        break;
    case M_ROBER:
        J[0][0] = -p[0];  J[0][1] = p[2] * u[2];                              J[0][2] = p[2] * u[1];
        J[1][0] = p[0];   J[1][1] = (T)-2 * p[1] * u[1] - p[2] * u[2];        J[1][2] = -(p[2] * u[1]);
        J[2][1] = (T)2 * p[1] * u[1];
        break;
    case M_DECAY:          // stiff_ode/gpu_ode_regression.jl:10-12: jac is the literal [-1.0f0]
        J[0][0] = (T)-1;
        break;
    case M_LINEAR15:
        for (int i = 0; i < 15; ++i) J[i][i] = (T)1.01;
        break;
    case M_OSC_T:
        J[0][1] = (T)1; J[1][0] = (T)-1;
        break;
    case M_QUAD_DECAY:
        J[0][0] = (T)-2 * p[0] * u[0];
        break;
    case M_BALL:
        J[0][1] = (T)1;
        break;
    case M_LIN_DAE:        // dae_jac, :32-37
        J[0][0] = -p[0];  J[0][1] = p[1];
        J[1][0] = (T)1;   J[1][1] = (T)1;
        break;
    case M_ROBER_DAE:      // the test passes no jac (ForwardDiff); rows 1-2 as rober_jac (:14-22), row 3 of the constraint
        J[0][0] = p[0] * (T)-1;  J[0][1] = u[2] * p[2];                                J[0][2] = p[2] * u[1];
        J[1][0] = p[0];          J[1][1] = u[1] * p[1] * (T)-2 + u[2] * p[2] * (T)-1;  J[1][2] = p[2] * u[1] * (T)-1;
        J[2][0] = (T)1;          J[2][1] = (T)1;                                       J[2][2] = (T)1;
        break;
    default: break;
    }
}

// analytic time gradient (nlsolve/type.jl:142-146, `f.tgrad` branch)
template <class T>
inline void model_tgrad(int id, T* dT, const T* u, const T* p, T t) {
    const int n = model_info(id).n;
    for (int i = 0; i < n; ++i) dT[i] = (T)0;
    if (id == M_OSC_T) dT[1] = -(p[0] * std::sin(t));
}

// diagonal noise g(u,p,t) (n values) -- test/gpu_kernel_de/gpu_sde_regression.jl:8-11,53-55
template <class T>
inline void model_g(int id, T* g, const T* u, const T* p, T t) {
    switch (id) {
    case M_GBM:             for (int i = 0; i < 3; ++i) g[i] = p[1] * u[i]; break;
    case M_LORENZ_ADDITIVE: g[0] = g[1] = g[2] = (T)3; break;
    case M_SCALAR_SDE:      g[0] = p[1] * u[0]; break;
    default: break;
    }
}
// non-diagonal noise G (n x m)
template <class T>
inline void model_G(int id, T (*G)[MAXN], const T* u, const T* p, T t) {
    if (id == M_GBM_ND) {   // gpu_sde_regression.jl:86-110 pattern: 2x4 matrix, rows scale with u
        G[0][0] = p[1] * u[0]; G[0][1] = (T)0.5 * p[1] * u[0]; G[0][2] = (T)0; G[0][3] = (T)0.25 * p[1] * u[0];
        G[1][0] = (T)0; G[1][1] = p[1] * u[1]; G[1][2] = (T)0.5 * p[1] * u[1]; G[1][3] = (T)0.25 * p[1] * u[1];
    }
}

// =====================================================================================
// Philox4x32-10 (Salmon et al. 2011), key = seed, counter = (block_lo, block_hi, traj_lo, traj_hi) -- see
// normals_for_step.  Normals by Box-Muller on (u32 -> (0,1]) pairs.
// The reference's RNG is backend-dependent and degenerate on its CPU backend, so
// this stream is OUR definition; oracle and kernel must agree bit-for-bit on the u32s.
// =====================================================================================
inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// u32 -> uniform in (0,1]: (x + 1) * 2^-32 for double; for float use the top 24 bits:
// ((x >> 8) + 1) * 2^-24 so that the value is exactly representable and never 0.
template <class T> inline T u01(uint32_t x);
template <> inline float  u01<float>(uint32_t x)  { return (float)((x >> 8) + 1u) * 5.9604644775390625e-8f; }
template <> inline double u01<double>(uint32_t x) { return ((double)x + 1.0) * 2.3283064365386963e-10; }

// two normals from two u32 (Box-Muller): r = sqrt(-2 ln u1), (r cos 2πu2, r sin 2πu2)
template <class T>
inline void box_muller(uint32_t a, uint32_t b, T& z0, T& z1) {
    T u1 = u01<T>(a), u2 = u01<T>(b);
    T r = std::sqrt((T)-2 * std::log(u1));
    T th = (T)6.283185307179586476925286766559 * u2;
    z0 = r * std::cos(th);
    z1 = r * std::sin(th);
}

// normals for (traj, step): the normals of a trajectory form one sequence over the whole run -- step j (0-based) with
// m noise terms uses q = j m ... j m + m - 1, and normal q is element q & 3 of Philox block q >> 2 (elements (0, 1) and
// (2, 3) are the two Box-Muller pairs of the block's four words), so no word of a block is thrown away
template <class T>
inline void normals_for_step(uint64_t seed, uint64_t traj, uint32_t step, int m, T* z) {
    // seed and trajectory index live in different Philox words (key = seed, counter = (block_lo, block_hi, traj)):
    // two seeds never share a stream, whatever the trajectory indices are
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint64_t have = ~(uint64_t)0;
    T zz[4] = {0, 0, 0, 0};
    for (int c = 0; c < m; ++c) {
        const uint64_t q = (uint64_t)step * (uint64_t)m + (uint64_t)c, blk = q >> 2;
        if (blk != have) {
            uint32_t r[4];
            philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)traj, (uint32_t)(traj >> 32), k0, k1, r);
            box_muller<T>(r[0], r[1], zz[0], zz[1]);
            box_muller<T>(r[2], r[3], zz[2], zz[3]);
            have = blk;
        }
        z[c] = zz[q & 3];
    }
}

// =====================================================================================
// Output views in the REFERENCE layout: us is (len x N) column-major of SVector{n,T}
// => element (k, i, c) at ((i*len + k)*n + c);  ts (k,i) at (i*len + k).
// lowerlevel_solve.jl:81-83, 311-324.
// =====================================================================================
template <class T>
struct Out {
    T* us; T* ts; int64_t len; int n;
    inline void put_u(int64_t k, const T* u) const {
        if (k < 0 || k >= len) return;  // the reference would write out of bounds (@inbounds); we drop
        for (int c = 0; c < n; ++c) us[k * n + c] = u[c];
    }
    inline void put_t(int64_t k, T t) const { if (k >= 0 && k < len) ts[k] = t; }
};

// controller constants -- integrators/integrator_utils.jl:1-11
template <class T>
struct Controller {
    T beta1, beta2, qmax, qmin, gamma, qoldinit;
    explicit Controller(int order) {
        beta1 = (T)(7.0 / (10.0 * order));
        beta2 = (T)(2.0 / (5.0 * order));
        qmax = (T)10.0; qmin = (T)(1.0 / 5.0); gamma = (T)(9.0 / 10.0); qoldinit = (T)1.0e-4;
    }
};

// scaled error norm: tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol), then
// ODE_DEFAULT_NORM = sqrt(sum(abs2, tmp) / n)   (gpu_tsit5_perform_step.jl:121-122;
// DiffEqBase ODE_DEFAULT_NORM for SArray; StaticArrays sum = left fold)
template <class T>
inline T error_norm(int n, const T* e, const T* uprev, const T* u, T abstol, T reltol) {
    T acc = (T)0;
    for (int c = 0; c < n; ++c) {
        T sc = abstol + jl_max(std::fabs(uprev[c]), std::fabs(u[c])) * reltol;
        T v = e[c] / sc;
        T sq = v * v;
        acc = (c == 0) ? sq : acc + sq;
    }
    return std::sqrt(acc / (T)n);
}

// PI step-size controller shared by every adaptive stepper (e.g. gpu_tsit5_perform_step.jl:124-137).
// Returns true when the attempt is accepted; updates dt (on reject) or dtnew/qold (on accept).
template <class T>
inline bool pi_controller(const Controller<T>& C, T EEst, T& dt, T& dtnew, T& qold, T t, T tf) {
    T q, q11 = (T)0;
    if (EEst == (T)0) {
        q = (T)1 / C.qmax;
    } else {
        q11 = jl_pow(EEst, C.beta1);
        q = q11 / jl_pow(qold, C.beta2);
    }
    if (EEst > (T)1) {
        dt = dt / jl_min((T)1 / C.qmin, q11 / C.gamma);
        return false;
    }
    q = jl_max((T)1 / C.qmax, jl_min((T)1 / C.qmin, q / C.gamma));
    qold = jl_max(EEst, C.qoldinit);
    dtnew = dt / q;
    dtnew = jl_min(std::fabs(dtnew), std::fabs(tf - t - dt));
    return true;
}

// =====================================================================================
// Explicit RK integrator (Tsit5 / Vern7 / Vern9), table driven.
// State mirrors GPUTsit5Integrator / GPUATsit5Integrator (integrators/nonstiff/types.jl:3-96).
// =====================================================================================
template <class T>
struct ErkInteg {
    const ErkTable* tab; int model; int n; Policy pol;
    T u[MAXN], uprev[MAXN], p[MAXN];
    T k[MAXS][MAXN];           // k[s] = stage s (1-based); kept after the step for the interpolant
    T t, tprev, dt, dtnew, tf, qold, abstol, reltol;
    bool u_modified;
    int naccept, nreject, nf, retcode;
    // tstops (gpu_tsit5_perform_step.jl:18-27 fixed, :158-164 adaptive; same code in the Vern steppers)
    const T* tstops = nullptr; int n_tstops = 0, tstops_idx = 0;

    inline void rhs(T* du, const T* uu, T tt) { model_f<T>(model, du, uu, p, tt); ++nf; }

    // `tstops[idx] - integ.t - integ.dt - T(100) * eps(T) < T(0)` with integ.t still the old time
    inline bool tstop_hit(T told, T h) const {
        return n_tstops > 0 && tstops_idx < n_tstops &&
               (tstops[tstops_idx] - told - h - (T)100 * std::numeric_limits<T>::epsilon() < (T)0);
    }

    // uprev + dt*(sum a_j k_j)  -- left fold in the written order; stage 2 is (dt*a21)*k1
    // (gpu_tsit5_perform_step.jl:36-47, gpu_vern7_perform_step.jl:118-146).
    inline void combine(T* out, const Term* row, int nterm, T h, bool stage2_form) {
        for (int c = 0; c < n; ++c) {
            if (stage2_form) {
                T a = h * (T)row[0].a;
                out[c] = pol.fma_stages ? fmaT<T>(a, k[row[0].j][c], uprev[c])
                                        : uprev[c] + a * k[row[0].j][c];
            } else {
                T s = (T)row[0].a * k[row[0].j][c];
                for (int m = 1; m < nterm; ++m) {
                    s = pol.fma_stages ? fmaT<T>((T)row[m].a, k[row[m].j][c], s)
                                       : s + (T)row[m].a * k[row[m].j][c];
                }
                out[c] = pol.fma_stages ? fmaT<T>(h, s, uprev[c]) : uprev[c] + h * s;
            }
        }
    }

    // All stages for step size h starting at (uprev, t0).  k[1] must already be valid for
    // FSAL methods; non-FSAL methods recompute it (gpu_vern7_perform_step.jl:117).
    inline void stages(T h, T t0, T* unew) {
        const int S = tab->stages;
        T tmp[MAXN];
        if (!tab->fsal) rhs(k[1], uprev, t0);
        if (tab->fsal) {
            // Tsit5: stages 2..6 are RHS args; stage-7 row gives u, k7 = f(u)
            for (int s = 2; s <= S - 1; ++s) {
                combine(tmp, tab->rows[s].t, tab->rows[s].n, h, s == 2);
                T ts = (tab->rows[s].c == 1.0) ? t0 + h : t0 + (T)tab->rows[s].c * h;
                rhs(k[s], tmp, ts);
            }
            combine(unew, tab->rows[S].t, tab->rows[S].n, h, false);
            rhs(k[S], unew, t0 + h);
        } else {
            // Verner: last two stages (g_{S-1}, g_S) are formed first, then both RHS, then u from b
            for (int s = 2; s <= S - 2; ++s) {
                combine(tmp, tab->rows[s].t, tab->rows[s].n, h, s == 2);
                rhs(k[s], tmp, t0 + (T)tab->rows[s].c * h);
            }
            T g1[MAXN], g2[MAXN];
            combine(g1, tab->rows[S - 1].t, tab->rows[S - 1].n, h, false);
            combine(g2, tab->rows[S].t, tab->rows[S].n, h, false);
            rhs(k[S - 1], g1, t0 + h);
            rhs(k[S], g2, t0 + h);
            combine(unew, tab->b, tab->nb, h, false);
        }
    }

    // Horner with muladd (@evalpoly): coefficients c[0..m-1], ascending powers
    static inline T evalpoly(T x, const T* c, int m) {
        T r = c[m - 1];
        for (int i = m - 2; i >= 0; --i) r = fmaT<T>(x, r, c[i]);
        return r;
    }

    // Dense output at Θ with step dt from y0 = uprev.  @muladd => fused chain, last product
    // pulled out first (MuladdMacro) == left-to-right fma accumulation.
    //   Tsit5: integrators/nonstiff/interpolants.jl:372-387 (+ SimpleDiffEq.bθs)
    //   Vern7: :28-190   Vern9: :192-370
    inline void interpolant(T theta, T h, T* out) {
        const ErkTable& tb = *tab;
        // extra stages (each an RHS evaluation) -- recomputed on EVERY call like the reference
        for (int e = 1; e <= tb.nextra; ++e) {
            const Row& r = tb.extra[e];
            T arg[MAXN];
            for (int c = 0; c < n; ++c) {
                T s = (T)r.t[0].a * k[r.t[0].j][c];
                for (int m = 1; m < r.n; ++m) s = fmaT<T>((T)r.t[m].a, k[r.t[m].j][c], s);
                arg[c] = fmaT<T>(h, s, uprev[c]);
            }
            // tprev + c*dt under @muladd -> muladd(c, dt, tprev)
            rhs(k[tb.stages + e], arg, fmaT<T>((T)r.c, h, tprev));
        }
        if (std::strcmp(tb.name, "vern7") == 0) {
            // (1-Θ)-factored form, interpolants.jl:28-93 (q-polynomials) and :95-190
            T q[16];
            for (int i = 0; i < tb.ninterp; ++i) {
                const Poly& P = tb.interp[i];
                T cf[8];
                if (P.stage == 1) {
                    // r011..r017 : s15=r017, s14=r016+s15, ..., s10=r012+s11 ; evalpoly(s10..s15)
                    T s[6];
                    s[5] = (T)P.r[6];
                    for (int m = 4; m >= 0; --m) s[m] = (T)P.r[m + 1] + s[m + 1];
                    for (int m = 0; m < 6; ++m) cf[m] = s[m];
                } else {
                    // ri2..ri7: s5=ri7, s4=ri6+s5, s3=ri5+s4, s2=ri4+s3, s1=ri3+s2 ;
                    // evalpoly(ri2+s1, s1, s2, s3, s4, s5)
                    T s[6];
                    s[5] = (T)P.r[5];
                    for (int m = 4; m >= 1; --m) s[m] = (T)P.r[m] + s[m + 1];
                    cf[0] = (T)P.r[0] + s[1];
                    for (int m = 1; m < 6; ++m) cf[m] = s[m];
                }
                q[i] = evalpoly(theta, cf, 6);
            }
            for (int c = 0; c < n; ++c) {
                // step_sum = dt * (b1*k1 + b4*k4 + ... + b9*k9)
                T ss = (T)tb.b[0].a * k[tb.b[0].j][c];
                for (int m = 1; m < tb.nb; ++m) ss = fmaT<T>((T)tb.b[m].a, k[tb.b[m].j][c], ss);
                T step_sum = h * ss;
                // correction = dt * (k1*q1 + k4*q4 + ... + k16*q16)
                T cs = k[tb.interp[0].stage][c] * q[0];
                for (int i = 1; i < tb.ninterp; ++i) cs = fmaT<T>(k[tb.interp[i].stage][c], q[i], cs);
                T corr = h * cs;
                // y0 + Θ*step_sum + Θ*(Θ-1)*correction
                //   -> muladd(Θ*(Θ-1), correction, muladd(Θ, step_sum, y0))
                T inner = fmaT<T>(theta, step_sum, uprev[c]);
                out[c] = fmaT<T>(theta * (theta - (T)1), corr, inner);
            }
            return;
        }
        // poly_b form: y0 + dt * Σ b_i(Θ) k_i
        T bth[20];
        for (int i = 0; i < tb.ninterp; ++i) {
            const Poly& P = tb.interp[i];
            T cf[12];
            for (int m = 0; m < P.n; ++m) cf[m] = (T)P.r[m];
            bth[i] = evalpoly(theta, cf, P.n);
        }
        for (int c = 0; c < n; ++c) {
            T s = bth[0] * k[slot_of(tb.interp[0].stage)][c];
            for (int i = 1; i < tb.ninterp; ++i) s = fmaT<T>(bth[i], k[slot_of(tb.interp[i].stage)][c], s);
            out[c] = fmaT<T>(h, s, uprev[c]);
        }
    }
    inline int slot_of(int stage) const { return stage; }  // we keep every stage in place

    // ---- fixed-dt step: gpu_tsit5_perform_step.jl:1-64, gpu_vern7:1-74, gpu_vern9:1-124 ----
    inline void step_fixed() {
        for (int c = 0; c < n; ++c) uprev[c] = u[c];
        T told = t;
        tprev = told;
        T h = dt;                                 // local dt: integ.dt keeps the nominal step
        if (tstop_hit(told, dt)) {
            t = tstops[tstops_idx];
            h = t - tprev;
            ++tstops_idx;
        } else {
            t = t + dt;                           // integ.t += dt (before the stages)
        }
        if (tab->fsal) {
            if (u_modified) { rhs(k[1], uprev, told); u_modified = false; }
            else for (int c = 0; c < n; ++c) k[1][c] = k[tab->stages][c];
        }
        T unew[MAXN];
        stages(h, told, unew);
        for (int c = 0; c < n; ++c) u[c] = unew[c];
        ++naccept;
    }

    // ---- adaptive step: gpu_tsit5_perform_step.jl:68-175, gpu_vern7:78-210, gpu_vern9:128-319
    // returns false on failure (dt < dtmin)
    inline bool step_adaptive(const Controller<T>& C) {
        T h = dtnew;
        T tcur = t;
        for (int c = 0; c < n; ++c) uprev[c] = u[c];
        if (tab->fsal) {
            if (u_modified) { rhs(k[1], uprev, tcur); u_modified = false; }
            else for (int c = 0; c < n; ++c) k[1][c] = k[tab->stages][c];
        } else if (u_modified) {
            // Vern7/9 evaluate k1 here too and then again inside the loop
            rhs(k[1], uprev, tcur); u_modified = false;
        }
        const T dtmin = thresh<T>(tab->dtmin_lit, tab->dtmin_via_f32);
        const T land = thresh<T>(tab->land_lit, tab->land_via_f32);
        T unew[MAXN], e[MAXN];
        for (;;) {
            if (h < dtmin) { retcode = RC_DT_LESS_THAN_MIN; return false; }
            stages(h, tcur, unew);
            for (int c = 0; c < n; ++c) {
                T s = (T)tab->bt[0].a * k[tab->bt[0].j][c];
                for (int m = 1; m < tab->nbt; ++m) {
                    s = pol.fma_stages ? fmaT<T>((T)tab->bt[m].a, k[tab->bt[m].j][c], s)
                                       : s + (T)tab->bt[m].a * k[tab->bt[m].j][c];
                }
                e[c] = h * s;
            }
            T EEst = error_norm<T>(n, e, uprev, unew, abstol, reltol);
            T dtn = dtnew;
            if (!pi_controller<T>(C, EEst, h, dtn, qold, tcur, tf)) { ++nreject; continue; }
            dtnew = dtn;
            dt = h;
            tprev = tcur;
            for (int c = 0; c < n; ++c) u[c] = unew[c];
            // Deviation (documented in DESIGN.md): when the remaining span tf - t - dt is positive but
            // below ulp(t), t + dt == t and the reference loops forever (dtnew is re-clamped to that
            // remainder on every step).  A step that does not advance t is taken to land on tf.
            if ((tf - tcur - h) < land) t = tf;
            else if (tstop_hit(tcur, h)) {
                // integ.t = tstop; integ.u = integ(integ.t)  (dense output of the step just taken)
                t = tstops[tstops_idx];
                T v[MAXN];
                interpolant((t - tprev) / dt, dt, v);
                for (int c = 0; c < n; ++c) u[c] = v[c];
                ++tstops_idx;
            } else { t = tcur + h; if (t == tcur && (tf - tcur - h) <= h) t = tf; }
            ++naccept;
            return true;
        }
    }
};

// saveat loop -- integrators/integrator_utils.jl:34-47
template <class T, class Integ>
inline void savevalues_saveat(Integ& I, const Out<T>& out, const T* saveat, int nsave, int& cur_t) {
    while (cur_t <= nsave && saveat[cur_t - 1] <= I.t) {
        T savet = saveat[cur_t - 1];
        T theta = (savet - I.tprev) / I.dt;
        T v[MAXN];
        I.interpolant(theta, I.dt, v);
        out.put_u(cur_t - 1, v);
        out.put_t(cur_t - 1, savet);
        ++cur_t;
    }
}

struct SolveArgs {
    int model, alg, adaptive, save_everystep, nsave, fma_stages;
    int64_t n_traj, len, u0_stride, p_stride, tspan_stride;
    double dt, abstol, reltol;
    uint64_t seed;
    int64_t max_iters;
    // events: tstops and discrete callbacks given as small specs that the tests
    // also lower to CUDA-C source for the device (tests/cases.py)
    int n_tstops = 0; const double* tstops = nullptr;
    int n_cb = 0; const int32_t* cb_i = nullptr; const double* cb_v = nullptr;
    int jac_mode = 0;   // stiff steppers: 0 analytic jac/tgrad, 1 finite differences, 2 forward-mode duals
    // continuous callbacks (GPUContinuousCallback, callbacks.jl:38-124):
    // cc_i[8*c + ..] = condition kind, condition index, affect kind (-1 = nothing), affect index,
    //                  affect_neg kind (-1 = nothing), affect_neg index, rootfind (0 Left, 1 Right, 2 None), 0
    // cc_v[6*c + ..] = condition value, affect value, affect_neg value, abstol, repeat_nudge, dtrelax
    int n_cc = 0; const int32_t* cc_i = nullptr; const double* cc_v = nullptr;
};

// continuous condition kinds: 0 u[i] - v | 1 t - v
template <class T>
inline T cc_condition(const SolveArgs& a, int c, const T* u, T t) {
    const int kind = a.cc_i[8 * c], idx = a.cc_i[8 * c + 1];
    const T v = (T)a.cc_v[6 * c];
    return kind == 0 ? u[idx] - v : t - v;
}
template <class T>
inline void cc_affect(const SolveArgs& a, int c, bool neg, T* u, T* p, bool& terminated) {
    const int kind = a.cc_i[8 * c + (neg ? 4 : 2)], idx = a.cc_i[8 * c + (neg ? 5 : 3)];
    const T v = (T)a.cc_v[6 * c + (neg ? 2 : 1)];
    switch (kind) {
    case 0: u[idx] = u[idx] + v; break;
    case 1: u[idx] = v; break;
    case 2: u[idx] = u[idx] * v; break;
    case 3: terminated = true; break;
    default: p[idx] = v; break;
    }
}
template <class T> inline T sign_(T x) { return x > (T)0 ? (T)1 : (x < (T)0 ? (T)-1 : x); }   // Base.sign
template <class T> inline T eps_of(T x) {      // Base.eps(x::AbstractFloat)
    if (!std::isfinite((double)x)) return std::numeric_limits<T>::quiet_NaN();
    const T ax = std::fabs(x);
    if (ax >= std::numeric_limits<T>::min()) return std::ldexp(std::numeric_limits<T>::epsilon(), std::ilogb(ax));
    return std::numeric_limits<T>::denorm_min();
}

// gpu_find_root: hand-written ITP, integrator_utils.jl:326-381 (scaled_k1 = 0.2, k2 = 2, n0 = 10)
template <class T, class F>
inline T itp_root(F&& fz, T left, T right, int rootfind) {
    T fl = fz(left), fr = fz(right);
    const T span0 = right - left;
    const T k1 = (T)0.2 / span0;
    T eps_s = span0 * (T)512;
    for (int it = 0; it < 100; ++it) {
        const T span = right - left;
        const T mid = (left + right) / (T)2;
        const T r = eps_s - span / (T)2;
        const T x_f = left + span * fl / (fl - fr);
        const T delta = jl_max(k1 * span * span, eps_of<T>(x_f));
        const T diff = mid - x_f;
        const T xt = (delta <= std::fabs(diff)) ? x_f + std::copysign(delta, diff) : mid;
        const T xp = (std::fabs(xt - mid) <= r) ? xt : mid - std::copysign(r, diff);
        const T yp = fz(xp);
        const T yps = yp * sign_(fr);
        if (yps > (T)0) { right = xp; fr = yp; }
        else if (yps < (T)0) { left = xp; fl = yp; }
        else { left = xp; right = xp; break; }
        eps_s = eps_s / (T)2;
        if (std::nextafter(left, std::numeric_limits<T>::infinity()) >= right) break;
    }
    return rootfind == 0 ? left : right;
}


// condition kinds: 0 t == v | 1 u[i] < v | 2 u[i] > v | 3 t >= v
// affect kinds:    0 u[i] += v | 1 u[i] = v | 2 u[i] *= v | 3 terminate!(integrator) | 4 p[i] = v
template <class T>
inline bool cb_condition(const SolveArgs& a, int c, const T* u, const T* p, T t) {
    (void)p;
    const int kind = a.cb_i[4 * c], idx = a.cb_i[4 * c + 1];
    const T v = (T)a.cb_v[2 * c];
    switch (kind) {
    case 0: return t == v;
    case 1: return u[idx] < v;
    case 2: return u[idx] > v;
    default: return t >= v;
    }
}
template <class T>
inline void cb_affect(const SolveArgs& a, int c, T* u, T* p, T t, bool& terminated) {
    (void)t;
    const int kind = a.cb_i[4 * c + 2], idx = a.cb_i[4 * c + 3];
    const T v = (T)a.cb_v[2 * c + 1];
    switch (kind) {
    case 0: u[idx] = u[idx] + v; break;
    case 1: u[idx] = v; break;
    case 2: u[idx] = u[idx] * v; break;
    case 3: terminated = true; break;
    default: p[idx] = v; break;
    }
}

enum { RC_TERMINATED = 6 };   // ReturnCode.Terminated (terminate! in an affect, integrator_utils.jl:52-66)
enum AlgId { A_TSIT5 = 0, A_VERN7 = 1, A_VERN9 = 2, A_ROS23 = 3, A_RODAS4 = 4, A_RODAS5P = 5,
             A_EM = 6, A_SIEA = 7, A_KVAERNO3 = 8, A_KVAERNO5 = 9 };

// ---- drivers: kernels.jl:1-72 (fixed) and :74-152 (adaptive) ----
template <class T, class Integ>
void drive(Integ& I, const SolveArgs& a, int order, T t0, T tf, const T* u0, const Out<T>& out,
           const T* saveat) {
    const bool has_saveat = saveat != nullptr;
    int cur_t = 0;
    int64_t step_idx = 1;
    if (has_saveat) {
        cur_t = 1;
        if (t0 == saveat[0]) { cur_t = 2; out.put_u(0, u0); }
    } else {
        out.put_t(0, t0);
        out.put_u(0, u0);
    }
    step_idx += 1;
    int64_t iters = 0;
    bool terminated = false;
    // savevalues! (integrator_utils.jl:13-50); adaptive integrators carry save_everystep = false (Q3)
    auto savevalues = [&]() {
        if (!has_saveat && a.save_everystep && !a.adaptive) {
            out.put_u(step_idx - 1, I.u);
            out.put_t(step_idx - 1, I.t);
            ++step_idx;
        } else if (has_saveat) {
            savevalues_saveat<T>(I, out, saveat, a.nsave, cur_t);
        }
    };
    // handle_callbacks! -> apply_discrete_callback! (integrator_utils.jl:69-96, 271-330): each
    // callback whose condition holds saves first, then sets u_modified and runs its affect
    // continuous callbacks: find_callback_time (integrator_utils.jl:383-442), the earliest event over the
    // set (DiffEqBase.find_first_continuous_callback), apply_callback! (:232-269) -- which always runs
    // continuous_callbacks[1]'s affects (in-tree quirk, :296-301)
    int event_last_time = 0;
    const T last_event_error = (T)0;             // never updated by the in-tree handle_callbacks!
    auto get_condition = [&](int c, T abst) -> T {        // :460-479
        if (abst == I.t) return cc_condition<T>(a, c, I.u, abst);
        if (abst == I.tprev) return cc_condition<T>(a, c, I.uprev, abst);
        T v[MAXN];
        I.interpolant((abst - I.tprev) / I.dt, I.dt, v);
        return cc_condition<T>(a, c, v, abst);
    };
    auto continuous = [&]() -> bool {
        bool occurred = false;
        T tmin = I.t, up = (T)0;
        int idx = 0;
        for (int c = 0; c < a.n_cc; ++c) {
            const T cc_abstol = (T)a.cc_v[6 * c + 3], nudge = (T)a.cc_v[6 * c + 4];
            const int rootfind = a.cc_i[8 * c + 6];
            T bottom_t = I.tprev;
            T bottom_condition = cc_condition<T>(a, c, I.uprev, I.tprev);
            if (event_last_time == c + 1 && std::fabs(bottom_condition - last_event_error) <= cc_abstol) {
                bottom_t = I.tprev + I.dt * nudge;        // nudge off the previous root
                bottom_condition = get_condition(c, bottom_t);
            }
            const T bottom_sign = sign_(bottom_condition);
            const T top_t = I.t;
            const T top_sign = sign_(get_condition(c, top_t));
            const bool ev = ((bottom_sign < (T)0 && a.cc_i[8 * c + 2] >= 0) || (bottom_sign > (T)0 && a.cc_i[8 * c + 4] >= 0)) &&
                            bottom_sign * top_sign <= (T)0;
            if (!ev) continue;
            T cbt;
            if (rootfind == 2 || top_sign == (T)0) cbt = top_t;
            else cbt = itp_root<T>([&](T x) { return get_condition(c, x); }, bottom_t, top_t, rootfind);
            if (cbt < tmin || !occurred) { tmin = cbt; up = bottom_sign; occurred = true; idx = c + 1; }
        }
        if (!occurred) { event_last_time = 0; return false; }
        event_last_time = idx;
        if (tmin != I.t) {                                // change_t_via_interpolation!, :186-206
            T v[MAXN];
            I.interpolant((tmin - I.tprev) / I.dt, I.dt, v);
            for (int c = 0; c < I.n; ++c) I.u[c] = v[c];
            step_idx -= (int64_t)std::nearbyint((double)((I.t - tmin) / I.dt));
            I.t = tmin;
        }
        if (a.adaptive && I.dtnew < (T)1.0e-12) {          // :240-249
            const T remaining = std::fabs(I.tf - I.t);
            I.dtnew = jl_min((T)a.cc_v[5] * I.dt, remaining);
        }
        savevalues();
        I.u_modified = true;
        if (up < (T)0) { if (a.cc_i[2] < 0) I.u_modified = false; else cc_affect<T>(a, 0, false, I.u, I.p, terminated); }
        else if (up > (T)0) { if (a.cc_i[4] < 0) I.u_modified = false; else cc_affect<T>(a, 0, true, I.u, I.p, terminated); }
        return true;
    };
    auto callbacks = [&]() -> bool {
        bool saved_in_cb = a.n_cc > 0 ? continuous() : false;
        if (a.n_cb > 0) saved_in_cb = false;          // handle_callbacks! takes saved_in_cb of the discrete pass (:318-326)
        for (int c = 0; c < a.n_cb; ++c) {
            if (cb_condition<T>(a, c, I.u, I.p, I.t)) {
                savevalues();
                saved_in_cb = true;
                I.u_modified = true;
                cb_affect<T>(a, c, I.u, I.p, I.t, terminated);
            }
        }
        return saved_in_cb;
    };
    if (a.adaptive) {
        Controller<T> C(order);
        while (I.t < tf && !terminated) {
            if (!I.step_adaptive(C)) return;
            if (!callbacks()) savevalues();
            if (++iters >= a.max_iters) { I.retcode = RC_MAXITERS; return; }
        }
    } else {
        while (I.t < tf && !terminated) {
            I.step_fixed();
            if (!callbacks()) savevalues();
            // (no step cap for fixed dt: the run makes exactly (tf - t0) / dt steps, like the reference)
        }
    }
    if (I.t > tf && !has_saveat) {          // kernels.jl:53-57 / :133-137
        T theta = (tf - I.tprev) / I.dt;
        T v[MAXN];
        I.interpolant(theta, I.dt, v);
        out.put_u(out.len - 1, v);
        out.put_t(out.len - 1, tf);
    }
    // kernels.jl:59-62 / :139-142.  (Adaptive integrators are built with save_everystep=false,
    // nonstiff/types.jl:378, so with save_everystep=true only row 1 is ever written -- Q3.)
    if (!has_saveat && !a.save_everystep) {
        out.put_u(1, I.u);
        out.put_t(1, I.t);
    }
    bool finite = true;
    for (int c = 0; c < I.n; ++c) if (!(I.u[c] == I.u[c]) || std::isinf((double)I.u[c])) finite = false;
    I.retcode = terminated ? (int)RC_TERMINATED : (finite ? (int)RC_SUCCESS : (int)RC_UNSTABLE);
}

#include "oracle_stiff.inc"
#include "oracle_kvaerno.inc"
#include "oracle_sde.inc"

template <class T>
int solve_T(const SolveArgs& a, const T* u0, const T* p, const T* tspan, const T* saveat,
            T* us, T* ts, int32_t* naccept, int32_t* nreject, int32_t* retcode, int nthreads) {
    const ModelInfo mi = model_info(a.model);
    if (mi.n == 0 || mi.n > MAXN) return -1;
    const ErkTable* tab = nullptr;
    int order = 0;
    switch (a.alg) {
    case A_TSIT5: tab = &tsit5_table; order = 5; break;
    case A_VERN7: tab = &vern7_table; order = 7; break;
    case A_VERN9: tab = &vern9_table; order = 9; break;
    case A_ROS23: order = 2; break;
    case A_RODAS4: order = 4; break;
    case A_RODAS5P: order = 5; break;
    case A_KVAERNO3: order = 3; break;
    case A_KVAERNO5: order = 5; break;
    case A_EM: case A_SIEA: break;
    default: return -2;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    std::vector<T> tstops_T(a.n_tstops);
    for (int i = 0; i < a.n_tstops; ++i) tstops_T[i] = (T)a.tstops[i];
    if ((a.n_tstops > 0 || a.n_cb > 0 || a.n_cc > 0) && (a.alg == A_EM || a.alg == A_SIEA)) return -3;   // events: ODE steppers
    if ((a.model == M_ROBER_DAE || a.model == M_LIN_DAE) && (a.alg < A_ROS23 || a.alg == A_EM || a.alg == A_SIEA)) return -5;   // mass matrices: implicit steppers
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < a.n_traj; ++i) {
        const T* ui = u0 + i * a.u0_stride;
        const T* pi = p ? p + i * a.p_stride : nullptr;
        const T t0 = tspan[i * a.tspan_stride], tf = tspan[i * a.tspan_stride + 1];
        Out<T> out{us + i * a.len * mi.n, ts + i * a.len, a.len, mi.n};
        int na = 0, nr = 0, rc = 0;
        if (tab) {
            ErkInteg<T> I;
            I.tab = tab; I.model = a.model; I.n = mi.n; I.pol = Policy{a.fma_stages};
            for (int c = 0; c < mi.n; ++c) { I.u[c] = ui[c]; I.uprev[c] = ui[c]; }
            for (int c = 0; c < mi.np; ++c) I.p[c] = pi[c];
            for (int s = 0; s < MAXS; ++s) for (int c = 0; c < mi.n; ++c) I.k[s][c] = ui[c];
            I.t = t0; I.tprev = t0; I.dt = (T)a.dt; I.dtnew = (T)a.dt; I.tf = tf;
            I.qold = (T)1.0e-4; I.abstol = (T)a.abstol; I.reltol = (T)a.reltol;
            I.u_modified = true; I.naccept = I.nreject = I.nf = 0; I.retcode = RC_DEFAULT;
            I.tstops = tstops_T.empty() ? nullptr : tstops_T.data(); I.n_tstops = (int)tstops_T.size(); I.tstops_idx = 0;
            drive<T>(I, a, order, t0, tf, ui, out, saveat);
            na = I.naccept; nr = I.nreject; rc = I.retcode;
        } else if (a.alg == A_ROS23 || a.alg == A_RODAS4 || a.alg == A_RODAS5P) {
            RosInteg<T> I;
            I.alg = a.alg; I.model = a.model; I.n = mi.n; I.pol = Policy{a.fma_stages};
            for (int c = 0; c < mi.n; ++c) { I.u[c] = ui[c]; I.uprev[c] = ui[c]; }
            for (int c = 0; c < mi.np; ++c) I.p[c] = pi[c];
            for (int s = 0; s < 3; ++s) for (int c = 0; c < mi.n; ++c) I.kk[s][c] = ui[c];
            I.t = t0; I.tprev = t0; I.dt = (T)a.dt; I.dtnew = (T)a.dt; I.tf = tf;
            I.qold = (T)1.0e-4; I.abstol = (T)a.abstol; I.reltol = (T)a.reltol;
            I.u_modified = true; I.naccept = I.nreject = I.nf = 0; I.retcode = RC_DEFAULT;
            T two = (T)2;
            I.d = (T)1 / (two + std::sqrt(two));    // stiff/types.jl:47-48
            I.jac_mode = a.jac_mode;
            I.tstops = tstops_T.empty() ? nullptr : tstops_T.data(); I.n_tstops = (int)tstops_T.size(); I.tstops_idx = 0;
            drive<T>(I, a, order, t0, tf, ui, out, saveat);
            na = I.naccept; nr = I.nreject; rc = I.retcode;
        } else if (a.alg == A_KVAERNO3 || a.alg == A_KVAERNO5) {
            KvInteg<T> I;
            I.alg = a.alg; I.model = a.model; I.n = mi.n; I.pol = Policy{a.fma_stages};
            for (int c = 0; c < mi.n; ++c) { I.u[c] = ui[c]; I.uprev[c] = ui[c]; I.k1[c] = I.k2[c] = I.k1next[c] = (T)0; }
            for (int c = 0; c < MAXN; ++c) I.p[c] = c < mi.np ? pi[c] : (T)0;
            I.t = t0; I.tprev = t0; I.dt = (T)a.dt; I.dtnew = (T)a.dt; I.tf = tf;
            I.qold = (T)1.0e-4; I.abstol = (T)a.abstol; I.reltol = (T)a.reltol;
            I.u_modified = true; I.naccept = I.nreject = I.nf = 0; I.retcode = RC_DEFAULT;
            I.jac_mode = a.jac_mode;
            I.tstops = tstops_T.empty() ? nullptr : tstops_T.data(); I.n_tstops = (int)tstops_T.size(); I.tstops_idx = 0;
            drive<T>(I, a, order, t0, tf, ui, out, saveat);
            na = I.naccept; nr = I.nreject; rc = I.retcode;
        } else {
            rc = sde_solve<T>(a, mi, (uint64_t)i, ui, pi, t0, tf, out, saveat, na);
        }
        if (naccept) naccept[i] = na;
        if (nreject) nreject[i] = nr;
        if (retcode) retcode[i] = rc;
    }
    return 0;
}

}  // namespace

extern "C" {

// dtype: 0 = float32, 1 = float64.  All arrays are host arrays of that dtype.
// us: [n_traj][len][n]   ts: [n_traj][len] (caller pre-fills ts with t0 like lowerlevel_solve.jl:82)
int degk_oracle_solve(int dtype, int model, int alg, int adaptive, int64_t n_traj,
                      const void* u0, int64_t u0_stride, const void* p, int64_t p_stride,
                      const void* tspan, int64_t tspan_stride,
                      double dt, double abstol, double reltol,
                      const void* saveat, int nsave, int save_everystep, uint64_t seed,
                      void* us, void* ts, int64_t len,
                      int32_t* naccept, int32_t* nreject, int32_t* retcode,
                      int fma_stages, int nthreads) {
    SolveArgs a;
    a.model = model; a.alg = alg; a.adaptive = adaptive; a.save_everystep = save_everystep;
    a.nsave = nsave; a.fma_stages = fma_stages; a.n_traj = n_traj; a.len = len;
    a.u0_stride = u0_stride; a.p_stride = p_stride; a.tspan_stride = tspan_stride;
    a.dt = dt; a.abstol = abstol; a.reltol = reltol; a.seed = seed;
    a.max_iters = 10000000;
    if (dtype == 0)
        return solve_T<float>(a, (const float*)u0, (const float*)p, (const float*)tspan,
                              (const float*)saveat, (float*)us, (float*)ts, naccept, nreject,
                              retcode, nthreads);
    return solve_T<double>(a, (const double*)u0, (const double*)p, (const double*)tspan,
                           (const double*)saveat, (double*)us, (double*)ts, naccept, nreject,
                           retcode, nthreads);
}

// same solve with events: tstops (Float64 values, converted to the dtype) and discrete callbacks
// cb_i[4*c + {0,1,2,3}] = condition kind, condition index, affect kind, affect index;
// cb_v[2*c + {0,1}] = condition value, affect value
int degk_oracle_solve_events(int dtype, int model, int alg, int adaptive, int64_t n_traj,
                             const void* u0, int64_t u0_stride, const void* p, int64_t p_stride,
                             const void* tspan, int64_t tspan_stride,
                             double dt, double abstol, double reltol,
                             const void* saveat, int nsave, int save_everystep, uint64_t seed,
                             void* us, void* ts, int64_t len,
                             int32_t* naccept, int32_t* nreject, int32_t* retcode,
                             int fma_stages, int nthreads,
                             const double* tstops, int n_tstops, const int32_t* cb_i, const double* cb_v, int n_cb,
                             int jac_mode, const int32_t* cc_i, const double* cc_v, int n_cc) {
    SolveArgs a;
    a.model = model; a.alg = alg; a.adaptive = adaptive; a.save_everystep = save_everystep;
    a.nsave = nsave; a.fma_stages = fma_stages; a.n_traj = n_traj; a.len = len;
    a.u0_stride = u0_stride; a.p_stride = p_stride; a.tspan_stride = tspan_stride;
    a.dt = dt; a.abstol = abstol; a.reltol = reltol; a.seed = seed;
    a.max_iters = 10000000;
    a.tstops = tstops; a.n_tstops = n_tstops; a.cb_i = cb_i; a.cb_v = cb_v; a.n_cb = n_cb;
    a.jac_mode = jac_mode;
    a.cc_i = cc_i; a.cc_v = cc_v; a.n_cc = n_cc;
    if (dtype == 0)
        return solve_T<float>(a, (const float*)u0, (const float*)p, (const float*)tspan,
                              (const float*)saveat, (float*)us, (float*)ts, naccept, nreject,
                              retcode, nthreads);
    return solve_T<double>(a, (const double*)u0, (const double*)p, (const double*)tspan,
                           (const double*)saveat, (double*)us, (double*)ts, naccept, nreject,
                           retcode, nthreads);
}

int degk_oracle_model_info(int model, int* n, int* np, int* m, int* diag) {
    ModelInfo mi = model_info(model);
    *n = mi.n; *np = mi.np; *m = mi.m; *diag = mi.diag_noise;
    return mi.n ? 0 : -1;
}

// raw Philox block + normals, for the RNG bit-exactness tests
void degk_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                        uint32_t k1, uint32_t* out) { philox4x32_10(c0, c1, c2, c3, k0, k1, out); }
void degk_oracle_normals_f32(uint64_t seed, uint64_t traj, uint32_t step, int m, float* z) {
    normals_for_step<float>(seed, traj, step, m, z);
}
void degk_oracle_normals_f64(uint64_t seed, uint64_t traj, uint32_t step, int m, double* z) {
    normals_for_step<double>(seed, traj, step, m, z);
}
int degk_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
