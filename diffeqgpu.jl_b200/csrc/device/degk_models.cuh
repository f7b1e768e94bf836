// degk_models.cuh -- built-in models, compiled ahead of time into libdegk.so.
//
// A model is a struct of static device functions; the JIT path (degk_jit.cpp) wraps user
// supplied function bodies into exactly this shape.  The bodies below are the reference's
// test problems and the BASELINE.json benchmark systems:
//   lorenz           test/gpu_kernel_de/gpu_ode_regression.jl:4-12 (+ jac test/lower_level_api.jl:19-47)
//   rober            test/gpu_kernel_de/stiff_ode/gpu_ode_mass_matrix.jl:5-22 in ODE form
//   decay            test/gpu_kernel_de/stiff_ode/gpu_ode_regression.jl:5-16
//   gbm / additive   test/gpu_kernel_de/gpu_sde_regression.jl:8-11, 53-55
// NOISE: 0 = none (ODE only), 1 = diagonal g (N values), 2 = general N x M matrix.
#pragma once
#include "degk_common.cuh"

namespace degk {

struct Lorenz {
    static constexpr int N = 3, NP = 3, M = 3, NOISE = 1;
    static constexpr bool HAS_JAC = true, HAS_TGRAD = true;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = p[0] * (u[1] - u[0]);
        du[1] = u[0] * (p[1] - u[2]) - u[1];
        du[2] = u[0] * u[1] - p[2] * u[2];
    }
    template <class T> static DEGK_DEV void jac(T (&J)[N][N], const T (&u)[N], const T* p, T t) {
        J[0][0] = -p[0];        J[0][1] = p[0];   J[0][2] = (T)0;
        J[1][0] = p[1] - u[2];  J[1][1] = (T)-1;  J[1][2] = -u[0];
        J[2][0] = u[1];         J[2][1] = u[0];   J[2][2] = -p[2];
    }
    template <class T> static DEGK_DEV void tgrad(T (&dT)[N], const T (&u)[N], const T* p, T t) {
        dT[0] = dT[1] = dT[2] = (T)0;
    }
    // additive noise g = (3,3,3): gpu_sde_regression.jl:53-55
    template <class T> static DEGK_DEV void g(T (&s)[N], const T (&u)[N], const T* p, T t) {
        s[0] = s[1] = s[2] = (T)3;
    }
};

struct HenonHeiles {   // u = (x, y, px, py); SURVEY §8d C3(ii)
    static constexpr int N = 4, NP = 0, M = 0, NOISE = 0;
    static constexpr bool HAS_JAC = false, HAS_TGRAD = false;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = u[2];
        du[1] = u[3];
        du[2] = -u[0] - (T)2 * u[0] * u[1];
        du[3] = -u[1] - (u[0] * u[0] - u[1] * u[1]);
    }
};

struct Rober {
    static constexpr int N = 3, NP = 3, M = 0, NOISE = 0;
    static constexpr bool HAS_JAC = true, HAS_TGRAD = true;
    static constexpr bool TGRAD_ZERO = true;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];
        du[1] = p[0] * u[0] - p[1] * (u[1] * u[1]) - p[2] * u[1] * u[2];
        du[2] = p[1] * (u[1] * u[1]);
    }
    template <class T> static DEGK_DEV void jac(T (&J)[N][N], const T (&u)[N], const T* p, T t) {
        J[0][0] = -p[0];  J[0][1] = p[2] * u[2];                        J[0][2] = p[2] * u[1];
        J[1][0] = p[0];   J[1][1] = (T)-2 * p[1] * u[1] - p[2] * u[2];  J[1][2] = -(p[2] * u[1]);
        J[2][0] = (T)0;   J[2][1] = (T)2 * p[1] * u[1];                 J[2][2] = (T)0;
    }
    template <class T> static DEGK_DEV void tgrad(T (&dT)[N], const T (&u)[N], const T* p, T t) {
        dT[0] = dT[1] = dT[2] = (T)0;
    }
};

// Robertson in DAE form with mass matrix diag(1, 1, 0): test/gpu_kernel_de/stiff_ode/gpu_ode_mass_matrix.jl:5-31
struct RoberDae {
    static constexpr int N = 3, NP = 3, M = 0, NOISE = 0;
    static constexpr bool HAS_JAC = true, HAS_TGRAD = true, HAS_MASS = true;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];
        du[1] = p[0] * u[0] - p[1] * (u[1] * u[1]) - p[2] * u[1] * u[2];
        du[2] = u[0] + u[1] + u[2] - (T)1;
    }
    template <class T> static DEGK_DEV void jac(T (&J)[N][N], const T (&u)[N], const T* p, T t) {
        J[0][0] = p[0] * (T)-1;  J[0][1] = u[2] * p[2];                                    J[0][2] = p[2] * u[1];
        J[1][0] = p[0];          J[1][1] = u[1] * p[1] * (T)-2 + u[2] * p[2] * (T)-1;      J[1][2] = p[2] * u[1] * (T)-1;
        J[2][0] = (T)1;          J[2][1] = (T)1;                                           J[2][2] = (T)1;
    }
    template <class T> static DEGK_DEV void tgrad(T (&dT)[N], const T (&u)[N], const T* p, T t) { dT[0] = dT[1] = dT[2] = (T)0; }
    template <class T> static DEGK_DEV void mass(T (&Mm)[N][N]) {
        DEGK_UNROLL for (int i = 0; i < N; ++i) DEGK_UNROLL for (int j = 0; j < N; ++j) Mm[i][j] = (T)0;
        Mm[0][0] = (T)1; Mm[1][1] = (T)1;
    }
};

struct Decay {
    static constexpr int N = 1, NP = 1, M = 0, NOISE = 0;
    static constexpr bool HAS_JAC = true, HAS_TGRAD = true;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = -p[0] * u[0];
    }
    // the reference test hard-codes jac = [-1.0f0] (valid for its p = 1)
    template <class T> static DEGK_DEV void jac(T (&J)[N][N], const T (&u)[N], const T* p, T t) {
        J[0][0] = (T)-1;
    }
    template <class T> static DEGK_DEV void tgrad(T (&dT)[N], const T (&u)[N], const T* p, T t) {
        dT[0] = (T)0;
    }
};

struct Linear15 {      // stiff_ode/gpu_ode_regression.jl:24-29, exercises the general LU
    static constexpr int N = 15, NP = 0, M = 0, NOISE = 0;
    static constexpr bool HAS_JAC = true, HAS_TGRAD = true;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        DEGK_UNROLL for (int i = 0; i < N; ++i) du[i] = (T)1.01 * u[i];
    }
    template <class T> static DEGK_DEV void jac(T (&J)[N][N], const T (&u)[N], const T* p, T t) {
        DEGK_UNROLL for (int i = 0; i < N; ++i)
            DEGK_UNROLL for (int j = 0; j < N; ++j) J[i][j] = (i == j) ? (T)1.01 : (T)0;
    }
    template <class T> static DEGK_DEV void tgrad(T (&dT)[N], const T (&u)[N], const T* p, T t) {
        DEGK_UNROLL for (int i = 0; i < N; ++i) dT[i] = (T)0;
    }
};

struct Gbm {           // dX = p1 X dt + p2 X dW, diagonal noise, 3 independent components
    static constexpr int N = 3, NP = 2, M = 3, NOISE = 1;
    static constexpr bool HAS_JAC = false, HAS_TGRAD = false;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        DEGK_UNROLL for (int i = 0; i < N; ++i) du[i] = p[0] * u[i];
    }
    template <class T> static DEGK_DEV void g(T (&s)[N], const T (&u)[N], const T* p, T t) {
        DEGK_UNROLL for (int i = 0; i < N; ++i) s[i] = p[1] * u[i];
    }
};

struct ScalarSde {
    static constexpr int N = 1, NP = 2, M = 1, NOISE = 1;
    static constexpr bool HAS_JAC = false, HAS_TGRAD = false;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = p[0] * u[0];
    }
    template <class T> static DEGK_DEV void g(T (&s)[N], const T (&u)[N], const T* p, T t) {
        s[0] = p[1] * u[0];
    }
};

struct OscT {          // non-autonomous: x'' = -x + p cos t  (exercises stage times and tgrad)
    static constexpr int N = 2, NP = 1, M = 0, NOISE = 0;
    static constexpr bool HAS_JAC = true, HAS_TGRAD = true;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = u[1];
        du[1] = -u[0] + p[0] * cos(t);
    }
    template <class T> static DEGK_DEV void jac(T (&J)[N][N], const T (&u)[N], const T* p, T t) {
        J[0][0] = (T)0; J[0][1] = (T)1; J[1][0] = (T)-1; J[1][1] = (T)0;
    }
    template <class T> static DEGK_DEV void tgrad(T (&dT)[N], const T (&u)[N], const T* p, T t) {
        dT[0] = (T)0; dT[1] = -(p[0] * sin(t));
    }
};

struct GbmNd {         // 2 states, 4 Wiener processes (non-diagonal noise, EM only)
    static constexpr int N = 2, NP = 2, M = 4, NOISE = 2;
    static constexpr bool HAS_JAC = false, HAS_TGRAD = false;
    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {
        du[0] = p[0] * u[0];
        du[1] = p[0] * u[1];
    }
    template <class T> static DEGK_DEV void G(T (&g)[N][M], const T (&u)[N], const T* p, T t) {
        g[0][0] = p[1] * u[0]; g[0][1] = (T)0.5 * p[1] * u[0]; g[0][2] = (T)0; g[0][3] = (T)0.25 * p[1] * u[0];
        g[1][0] = (T)0; g[1][1] = p[1] * u[1]; g[1][2] = (T)0.5 * p[1] * u[1]; g[1][3] = (T)0.25 * p[1] * u[1];
    }
};

}  // namespace degk
