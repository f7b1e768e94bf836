// degk_ode_saves.cuh -- deferred, batched saves of the adaptive kernel (used by degk_ode_kernels4.cuh).
//
// A lane that crosses a save point only appends a small record {trajectory, save cursor, tprev, h, tnew, uprev} to a
// queue in shared memory.  The queue is LANE-PRIVATE: record k of lane l lives at index l + 32 k of the warp's record
// array, so a push is two predicated 16-byte stores at an address the lane keeps in a register plus one predicated
// add -- no ballot / popcount rank per pass (the per-warp queue of the earlier generations spent 2 VOTE, 4 POPC and
// ~8 integer instructions per pass on ranks; POPC alone occupies the dispatch port for ~4 cycles on sm_100,
// profiles/r2_pipe_probe2.jsonl).  When some lane's private part is nearly full the warp flushes: a prefix sum of the
// per-lane counts, a directory (record index per queue position) in shared memory, then full batches of 32 records
// are processed one per lane -- the stages of that step are re-evaluated (same inputs, same instruction sequence =>
// same bits), interpolated (reference integrator_utils.jl:34-47, `_ode_interpolant`) and stored -- and the < 32
// left-over records are re-homed one per lane, which also levels the counts.  Re-computing a step costs less than 1 %
// of a trajectory (11 saves vs ~170 steps) and runs at full SIMT efficiency; in the first-generation kernel 25 % of
// all issued warp-instructions were the saveat block executed with ~2 of 32 lanes active.
#pragma once
#include "degk_common.cuh"
#include "degk_pack.cuh"

#ifndef DEGK_RETIRE_BATCH
#define DEGK_RETIRE_BATCH 6    // measured on C2 (8.4 M trajectories): 1: 90.0, 2: 94.2, 4: 97.7, 6: 98.5, 8: 98.4, 12: 97.2 G steps/s
#endif

namespace degk {


DEGK_DEV float  lds_(u32 saddr, float)  { float v;  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr)); return v; }
DEGK_DEV double lds_(u32 saddr, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr)); return v; }

// queue record of one deferred save (lives in shared memory)
template <class T, int N>
struct __align__(16) SaveRec {
    int traj;           // index in this launch (n_traj < 2^31, checked by the host)
    u32 cur;            // shared-memory address of the first saveat entry to write (1-based cursor c: sv_saddr + c * sizeof(T))
    T tprev, h, tnew;
    T u[N];             // state at the beginning of the step
};

// copy a record through 16-byte words so that the compiler emits vector shared-memory accesses
template <class R>
DEGK_DEV void rec_copy(R* dst, const R* src) {
    static_assert(sizeof(R) % 16 == 0, "SaveRec must be a multiple of 16 bytes");
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    DEGK_UNROLL for (int i = 0; i < (int)(sizeof(R) / 16); ++i) d[i] = s[i];
}

// ------------------------------------------------------------------------------------------
// steppers whose fixed-dt and adaptive attempts keep different interpolation data (Kvaerno) say so
template <class...> struct replay_void_ { typedef void type; };
template <class M, class = void> struct replay_adaptive_of { static constexpr bool value = false; };
template <class M> struct replay_adaptive_of<M, typename replay_void_<decltype(M::REPLAY_ADAPTIVE)>::type> { static constexpr bool value = M::REPLAY_ADAPTIVE; };

// records per lane of the lane-private queue for W trajectories per thread (must exceed W): 4 W while a record is
// 32 bytes (8 KB per warp for the packed Float32 build), fewer for larger states so that the queue stays a few KB
template <class T, int N, int W> __host__ __device__ constexpr int save_queue_depth() {
    return sizeof(SaveRec<T, N>) <= 32 ? 4 * W : (sizeof(SaveRec<T, N>) <= 64 ? 2 * W : W + 2);
}

// One warp processes up to 32 queued save records, one per lane (scalar method): lane j takes the record whose
// index the directory holds at position first + j.  `sv_saddr` is the shared-memory address of the 1-based staged
// saveat array (entry c at sv_saddr + c * sizeof(T); +inf behind the last entry).
template <class T, class Model, class MethodS>
DEGK_DEV void process_saves(const KArgs& a, const SaveRec<T, Model::N>* q, const unsigned short* dir, int first, int count,
                            u32 sv_saddr) {
    constexpr int N = Model::N;
    const int lane = (int)lane_id();
    if (lane < count) {
        SaveRec<T, N> r;
        rec_copy(&r, q + dir[first + lane]);
        T uprev[N], unew[N], err[N];
        T p[Model::NP > 0 ? Model::NP : 1];
        DEGK_UNROLL for (int c = 0; c < N; ++c) uprev[c] = r.u[c];
        if (Model::NP > 0) {
            const T* pp = (const T*)a.p + (i64)r.traj * a.p_stride;
            DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) p[c] = pp[c];
        }
        const T tprev = r.tprev, h = r.h, tnew = r.tnew;
        typename MethodS::Keep K;
        MethodS::init(K, uprev, p, tprev);                  // FSAL k1 = f(uprev, p, tprev)
        MethodS::template attempt<replay_adaptive_of<MethodS>::value>(K, uprev, p, tprev, h, unew, err);
        MethodS::on_accept(K);
        u32 ca = r.cur;
        i64 row = (i64)((ca - sv_saddr) / (u32)sizeof(T)) - 1;
        for (;;) {                                          // integrator_utils.jl:34-47 (the sentinel ends the loop)
            const T savet = lds_(ca, (T)0);
            if (!(savet <= tnew)) break;
            const T theta = (savet - tprev) / h;
            T v[N];
            MethodS::interp(K, theta, h, uprev, unew, p, tprev, v);
            store_u<T, N>(a, r.traj, row, v);
            store_t<T>(a, r.traj, row, savet);
            ca += (u32)sizeof(T); ++row;
        }
    }
}

// ---- per-half helpers (scalar overloads serve W == 1) ----
DEGK_DEV float  vmaxabs(float a, float b)   { return fmaxf(fabsf(a), fabsf(b)); }
DEGK_DEV double vmaxabs(double a, double b) { return fmax(fabs(a), fabs(b)); }
DEGK_DEV Pk2 vmaxabs(Pk2 a, Pk2 b) { return Pk2(vmaxabs(a.lo(), b.lo()), vmaxabs(a.hi(), b.hi())); }
DEGK_DEV float  vrcp(float x)  { return rcp_(x); }
DEGK_DEV double vrcp(double x) { return 1.0 / x; }
DEGK_DEV Pk2 vrcp(Pk2 x) { return Pk2(rcp_(x.lo()), rcp_(x.hi())); }
DEGK_DEV float  vlog2(float x)  { return log2_(x); }
DEGK_DEV double vlog2(double x) { return log2(x); }
DEGK_DEV Pk2 vlog2(Pk2 x) { return Pk2(log2_(x.lo()), log2_(x.hi())); }
DEGK_DEV float  vexp2(float x)  { return exp2_(x); }
DEGK_DEV double vexp2(double x) { return exp2(x); }
DEGK_DEV Pk2 vexp2(Pk2 x) { return Pk2(exp2_(x.lo()), exp2_(x.hi())); }
DEGK_DEV float  vclamp(float x, float lo, float hi)    { return fminf(fmaxf(x, lo), hi); }   // NaN -> lo
DEGK_DEV double vclamp(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
DEGK_DEV Pk2 vclamp(Pk2 x, float lo, float hi) { return Pk2(vclamp(x.lo(), lo, hi), vclamp(x.hi(), lo, hi)); }

// predicated 16-byte shared-memory store: no branch, inactive lanes store nothing
DEGK_DEV void sts128_if(bool pred, u32 saddr, uint4 w) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.v4.u32 [%1], {%2, %3, %4, %5};\n\t}"
                 :: "r"((u32)pred), "r"(saddr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}
// keep an address in a register: stops the compiler from re-deriving it from %tid / the shared
// window base in every loop iteration (it did: 2 x S2R + 8 integer instructions per iteration)
DEGK_DEV u32 opaque(u32 x) { asm volatile("" : "+r"(x)); return x; }
DEGK_DEV float  opaque_f(float x)  { asm volatile("" : "+f"(x)); return x; }
DEGK_DEV double opaque_f(double x) { asm volatile("" : "+d"(x)); return x; }
// in-place predicated increment `if (pred) ++x` as ONE predicated instruction (see assign_if)
DEGK_DEV void inc_if(bool pred, u32& x) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(x) : "r"((u32)pred));
}
DEGK_DEV void inc_if(bool pred, int& x) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(x) : "r"((u32)pred));
}
// `if (pred) x += d` as one predicated add (d: compile-time constant)
template <u32 D> DEGK_DEV void add_if(bool pred, u32& x) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p add.u32 %0, %0, %2;\n\t}" : "+r"(x) : "r"((u32)pred), "n"(D));
}
// `if (pred) x = shared[saddr]` as one predicated load
DEGK_DEV void lds_if(bool pred, u32 saddr, float& x) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ld.shared.f32 %0, [%2];\n\t}" : "+f"(x) : "r"((u32)pred), "r"(saddr));
}
DEGK_DEV void lds_if(bool pred, u32 saddr, double& x) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ld.shared.f64 %0, [%2];\n\t}" : "+d"(x) : "r"((u32)pred), "r"(saddr));
}

template <class R>
DEGK_DEV void rec_store_if(bool pred, u32 saddr, const R& r) {
    static_assert(sizeof(R) % 16 == 0, "SaveRec must be a multiple of 16 bytes");
    const uint4* s = reinterpret_cast<const uint4*>(&r);
    DEGK_UNROLL for (int i = 0; i < (int)(sizeof(R) / 16); ++i) sts128_if(pred, saddr + 16u * i, s[i]);
}

}  // namespace degk
