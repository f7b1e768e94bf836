// degk_sde_kernels.cuh -- whole-solve SDE kernels (fixed dt, one thread = one trajectory).
//
//   ALG_EM    Euler-Maruyama, diagonal or general noise   reference perform_step/gpu_em_perform_step.jl:1-74
//   ALG_SIEA  SIEA weak order 2, diagonal noise only      reference perform_step/gpu_siea_perform_step.jl:1-139
//
// RNG.  The reference seeds a backend-dependent device RNG per thread (`Random.seed!(prob.seed)`,
// gpu_em_perform_step.jl:8; degenerate on its CPU backend).  Here the stream is
// counter based and therefore independent of launch geometry and of how an ensemble is sharded
// over GPUs: normal q of global trajectory i (q counts the normals of the whole run: step j with m noise terms uses
// q = j m ... j m + m - 1) is element q & 3 of BoxMuller(Philox4x32-10(key = seed, counter = (q >> 2, i_lo, i_hi)))
// -- see normals_of_block.  Seed and trajectory index occupy different Philox words, so two seeds
// never share a stream (with key = seed ^ i, seeds s and s ^ d would only permute the paths of an ensemble).
// The u32 stream is bit-identical to the oracle's.
#pragma once
#include "degk_common.cuh"

#ifndef DEGK_SDE_MINBLOCKS
#define DEGK_SDE_MINBLOCKS 1
#endif

namespace degk {

enum { ALG_EM = 0, ALG_SIEA = 1 };

DEGK_DEV void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 (&out)[4]) {
    const u32 M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    DEGK_UNROLL for (int r = 0; r < 10; ++r) {
        const u32 hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const u32 hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const u32 n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// u32 -> (0,1]
DEGK_DEV float  u01_(u32 x, float)  { return (float)((x >> 8) + 1u) * 5.9604644775390625e-8f; }
DEGK_DEV double u01_(u32 x, double) { return ((double)x + 1.0) * 2.3283064365386963e-10; }

DEGK_DEV void box_muller(u32 a, u32 b, float& z0, float& z1) {
    float s, c;
#if DEGK_STRICT
    const float u1 = u01_(a, 0.f), u2 = u01_(b, 0.f);
    const float th = 6.283185307179586476925286766559f * u2;
    const float r = sqrtf(-2.0f * logf(u1));
    sincosf(th, &s, &c);
#else
    // u1 is in [2^-24, 1], never subnormal: the MUFU logarithm and square root as they are (lg2 / sqrt.approx.ftz),
    // without the range guards and Newton step of __logf / sqrtf (12 instructions and two branches per pair); the angle
    // 2 pi u2 = n2 (2 pi 2^-24) is one multiply of the integer (one rounding instead of two).  (The same folding for
    // the logarithm, -2 ln2 (lg2 n1 - 24), would cancel near u1 = 1 and is not done.)
    const float u1 = u01_(a, 0.f);
    const float th = 3.7450702829239286e-7f * (float)((b >> 8) + 1u);
    float l2, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));      // sqrt(-2 ln u1)
    __sincosf(th, &s, &c);
#endif
    z0 = r * c;
    z1 = r * s;
}
DEGK_DEV void box_muller(u32 a, u32 b, double& z0, double& z1) {
    const double u1 = u01_(a, 0.0), u2 = u01_(b, 0.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincos(6.283185307179586476925286766559 * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// The normals of a trajectory form ONE sequence over the whole run: step j (0-based) with m noise terms consumes normals
// q = j m ... j m + m - 1, and normal q is element q & 3 of Philox block q >> 2 (elements (0, 1) and (2, 3) are the two
// Box-Muller pairs of the block's four words).  No word of a block is thrown away: m = 3 (BASELINE config 5) needs three
// blocks per four steps instead of four, m = 1 one instead of four.
template <class T>
DEGK_DEV void normals_of_block(u32 k0, u32 k1, u32 g0, u32 g1, u64 blk, T (&zz)[4]) {
    u32 r[4];
    philox4x32_10((u32)blk, (u32)(blk >> 32), g0, g1, k0, k1, r);
    box_muller(r[0], r[1], zz[0], zz[1]);
    box_muller(r[2], r[3], zz[2], zz[3]);
}
// steps per group / blocks per group: the shortest run of steps that consumes whole blocks
template <int MM> struct NormalGroup {
    static constexpr int G = (MM % 4 == 0) ? 1 : (MM % 2 == 0 ? 2 : 4);
    static constexpr int NB = G * MM / 4;
};
// all normals of steps group*G ... group*G + G - 1 (z[g * MM + c]); the NB Philox chains are independent of each other
// and of the state, so they are in flight together
template <class T, int MM>
DEGK_DEV void normals_for_group(u32 k0, u32 k1, u32 g0, u32 g1, u64 group, T (&z)[NormalGroup<MM>::G * MM]) {
    constexpr int NB = NormalGroup<MM>::NB;
    const u64 b0 = group * (u64)NB;
    DEGK_UNROLL for (int b = 0; b < NB; ++b) {
        T zz[4];
        normals_of_block<T>(k0, k1, g0, g1, b0 + (u64)b, zz);
        DEGK_UNROLL for (int q = 0; q < 4; ++q) z[4 * b + q] = zz[q];
    }
}
// the normals of one step wherever it lies in the sequence (the steps behind the last whole group)
template <class T, int MM>
DEGK_DEV void normals_for_step(u32 k0, u32 k1, u32 g0, u32 g1, u32 step, T (&z)[MM]) {
    const u64 q0 = (u64)step * (u64)MM;
    u64 have = ~(u64)0;
    T zz[4] = {(T)0, (T)0, (T)0, (T)0};
    DEGK_UNROLL for (int c = 0; c < MM; ++c) {
        const u64 q = q0 + (u64)c, blk = q >> 2;
        if (blk != have) { normals_of_block<T>(k0, k1, g0, g1, blk, zz); have = blk; }
        const u32 e = (u32)q & 3u;
        z[c] = e == 0 ? zz[0] : (e == 1 ? zz[1] : (e == 2 ? zz[2] : zz[3]));
    }
}

// optional ensemble reduction: sum(u), sum(u^2) over trajectories for output row k
// (replaces the host-side `reduction` of src/solve.jl:123-125 for mean/variance ensembles).
// Must be called by all 32 lanes of the warp; `valid` masks lanes without a trajectory.
template <class T, int N>
DEGK_DEV void reduce_row(const KArgs& a, i64 k, const T (&v)[N], bool valid) {
    DEGK_UNROLL for (int c = 0; c < N; ++c) {
        double s1 = valid ? (double)v[c] : 0.0;
        double s2 = s1 * s1;
        DEGK_UNROLL for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane_id() == 0) {
            atomicAdd(a.reduce + (k * N + c) * 2 + 0, s1);
            atomicAdd(a.reduce + (k * N + c) * 2 + 1, s2);
        }
    }
}

template <class T, class Model, int ALG>
DEGK_DEV void sde_solve_body(const KArgs& a) {
    constexpr int N = Model::N;
    constexpr int MM = (Model::NOISE == 2) ? Model::M : N;
    static_assert(Model::NOISE != 0, "model has no noise term");
    static_assert(ALG == ALG_EM || Model::NOISE == 1, "SIEA supports diagonal noise only");
    const i64 traj = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = traj < a.n_traj;
    const i64 tl = active ? traj : 0;      // idle lanes shadow trajectory 0 but never store
    T u[N], uprev[N];
    T p[Model::NP > 0 ? Model::NP : 1];
    T t0, tf;
    load_problem<T, Model>(a, tl, u, p, t0, tf);
    const T dt = (T)a.dt;
    const T sqdt = sqrt_(dt);
    const bool has_saveat = a.saveat != nullptr;
    const T* saveat = (const T*)a.saveat + (has_saveat ? tl * a.saveat_stride : 0);            // this trajectory's grid
    const bool red = a.reduce != nullptr;
    const u64 gid = (u64)(a.traj_offset + tl);
    const u32 k0 = (u32)a.seed, k1 = (u32)(a.seed >> 32);
    const u32 g0 = (u32)gid, g1 = (u32)(gid >> 32);
    int cur = 0;
    i64 ts_written = 0;
    if (has_saveat) {                       // gpu_em_perform_step.jl:27-37
        cur = 1;
        if (t0 == saveat[0]) {
            cur = 2;
            if (active) { store_u<T, N>(a, traj, 0, u); store_t<T>(a, traj, 0, t0); }
            if (red) reduce_row<T, N>(a, 0, u, active);
        }
    } else {
        if (active) { store_t<T>(a, traj, 0, t0); store_u<T, N>(a, traj, 0, u); }
        if (red) reduce_row<T, N>(a, 0, u, active);
        ts_written = 1;
    }
    T t = t0;
    // n = floor(Int, abs(tf - t0) / abs(dt)) + 1   (gpu_em_perform_step.jl:44)
    const i64 nst = (i64)floor((double)(abs_(tf - t0) / abs_(dt))) + 1;
    // one step: u <- stepper(uprev = u), t <- t + dt
    auto advance_z = [&](const T (&z)[MM]) {
        DEGK_UNROLL for (int c = 0; c < N; ++c) uprev[c] = u[c];
        if constexpr (ALG == ALG_EM) {
            T f[N];
            Model::template f<T>(f, uprev, p, t);
            if constexpr (Model::NOISE == 1) {
                T g[N];
                Model::template g<T>(g, uprev, p, t);
                // u = uprev + f*dt + sqdt*g .* randn   (gpu_em_perform_step.jl:49-50)
                DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = (uprev[c] + f[c] * dt) + (sqdt * g[c]) * z[c];
            } else {
                T G[N][Model::M > 0 ? Model::M : 1];
                Model::template G<T>(G, uprev, p, t);
                DEGK_UNROLL for (int c = 0; c < N; ++c) {
                    T s = (sqdt * G[c][0]) * z[0];
                    DEGK_UNROLL for (int q = 1; q < MM; ++q) s = s + (sqdt * G[c][q]) * z[q];
                    u[c] = (uprev[c] + f[c] * dt) + s;
                }
            }
        } else {
            // SIEA constants: gpu_siea_perform_step.jl:23-47
            const T al1 = (T)0.5, al2 = (T)0.5, ga1 = (T)0.5, la1 = (T)0.25, la2 = (T)-0.25, la3 = (T)0.25,
                    mu1 = (T)0.25, mu2 = (T)0.25, mu3 = (T)-0.25, mu0 = (T)1, mubar0 = (T)1, la0 = (T)1,
                    labar0 = (T)1, nu1 = (T)1, nu2 = (T)0, be2 = (T)1, be3 = (T)0, de2 = (T)-1, de3 = (T)0;
            T k0v[N], g0[N], k1v[N], g1[N], g2[N], dW[N], W2[N], W3[N], arg[N];
            Model::template f<T>(k0v, uprev, p, t);
            Model::template g<T>(g0, uprev, p, t);
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                dW[c] = sqdt * z[c];
                W2[c] = (dW[c] * dW[c]) / sqdt;
                W3[c] = (nu2 * (dW[c] * dW[c] * dW[c])) / dt;
            }
            DEGK_UNROLL for (int c = 0; c < N; ++c)
                arg[c] = ((uprev[c] + (la0 * k0v[c]) * dt) + (nu1 * g0[c]) * dW[c]) + g0[c] * W3[c];
            Model::template f<T>(k1v, arg, p, t + mu0 * dt);
            DEGK_UNROLL for (int c = 0; c < N; ++c)
                arg[c] = ((uprev[c] + (labar0 * k0v[c]) * dt) + (be2 * g0[c]) * sqdt) + (be3 * g0[c]) * W2[c];
            Model::template g<T>(g1, arg, p, t + mubar0 * dt);
            DEGK_UNROLL for (int c = 0; c < N; ++c)
                arg[c] = ((uprev[c] + (labar0 * k0v[c]) * dt) + (de2 * g0[c]) * sqdt) + (de3 * g0[c]) * W2[c];
            Model::template g<T>(g2, arg, p, t + mubar0 * dt);
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                const T un = uprev[c] + (al1 * k0v[c] + al2 * k1v[c]) * dt;
                const T x = (ga1 * g0[c]) * dW[c];
                const T y = ((la1 * dW[c] + la2 * sqdt) + la3 * W2[c]) * g1[c];
                const T w = ((mu1 * dW[c] + mu2 * sqdt) + mu3 * W2[c]) * g2[c];
                u[c] = un + ((x + y) + w);
            }
        }
        t = t + dt;
    };
    // what follows step j (1-based row index j - 1) when rows are kept
    auto after_step = [&](i64 j) {
        if (!has_saveat) {
            if (a.save_everystep) {
                if (active) { store_u<T, N>(a, traj, j - 1, u); store_t<T>(a, traj, j - 1, t); }
                if (red && j - 1 < a.n_rows) reduce_row<T, N>(a, j - 1, u, active);
                ts_written = j;
            }
        } else {
            while (cur <= a.n_saveat && saveat[cur - 1] <= t) {   // linear interpolation, :59-67
                const T savet = saveat[cur - 1];
                const T theta = (savet - (t - dt)) / dt;
                T v[N];
                DEGK_UNROLL for (int c = 0; c < N; ++c) v[c] = uprev[c] + (u[c] - uprev[c]) * theta;
                if (active) { store_u<T, N>(a, traj, cur - 1, v); store_t<T>(a, traj, cur - 1, savet); }
                if (red) reduce_row<T, N>(a, cur - 1, v, active);
                ++cur;
            }
        }
    };
    // Steps run in groups that consume whole Philox blocks (NormalGroup: four steps and three blocks for m = 3); the
    // steps behind the last whole group take their normals one step at a time.  Two copies of the loop: end points only
    // (ensemble moments, BASELINE config 5) has nothing to test or store per step -- the save options cost ~25 of the
    // ~130 instructions of a step.
    constexpr int G = NormalGroup<MM>::G;
    const bool keep_rows = has_saveat || a.save_everystep;
    i64 j = 2;
#define DEGK_SDE_GROUPS(AFTER)                                                                     \
    for (; j + (G - 1) <= nst; j += G) {                                                           \
        T zall[G * MM];                                                                            \
        normals_for_group<T, MM>(k0, k1, g0, g1, (u64)(j - 2) / (u64)G, zall);                     \
        DEGK_UNROLL for (int g = 0; g < G; ++g) {                                                  \
            T z[MM];                                                                               \
            DEGK_UNROLL for (int c = 0; c < MM; ++c) z[c] = zall[g * MM + c];                      \
            advance_z(z);                                                                          \
            AFTER(j + g);                                                                          \
        }                                                                                          \
    }                                                                                              \
    for (; j <= nst; ++j) {                                                                        \
        T z[MM];                                                                                   \
        normals_for_step<T, MM>(k0, k1, g0, g1, (u32)(j - 2), z);                                  \
        advance_z(z);                                                                              \
        AFTER(j);                                                                                  \
    }
#define DEGK_SDE_NOTHING(jj) ((void)0)
    if (keep_rows) { DEGK_SDE_GROUPS(after_step) } else { DEGK_SDE_GROUPS(DEGK_SDE_NOTHING) }
#undef DEGK_SDE_GROUPS
#undef DEGK_SDE_NOTHING
    if (!has_saveat && !a.save_everystep) {
        if (active) { store_u<T, N>(a, traj, 1, u); store_t<T>(a, traj, 1, t); }
        if (red) reduce_row<T, N>(a, 1, u, active);
        ts_written = 2;
    }
    u32 nfail = 0;
    if (active) {
        fill_unwritten_ts<T>(a, traj, has_saveat ? (i64)(cur - 1) : ts_written, t0);
        bool fin = true;
        DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(u[c]);
        if (a.retcode) a.retcode[traj] = fin ? RC_SUCCESS : RC_UNSTABLE;
        if (a.naccept) a.naccept[traj] = (int)(nst - 1);
        if (a.nreject) a.nreject[traj] = 0;
        if (!fin) nfail = 1;
    }
    if (a.totals) {
        u32 ns = active ? (u32)(nst - 1) : 0u;
        DEGK_UNROLL for (int o = 16; o > 0; o >>= 1) {
            ns += __shfl_xor_sync(0xffffffffu, ns, o);
            nfail += __shfl_xor_sync(0xffffffffu, nfail, o);
        }
        if (lane_id() == 0) {
            atomicAdd(a.totals + 0, (u64)ns);
            if (nfail) atomicAdd(a.totals + 2, (u64)nfail);
        }
    }
}

}  // namespace degk
