// degk_pack.cuh -- Pk2: two Float32 trajectories packed in one 64-bit register pair.
//
// Blackwell (sm_100) adds packed FP32 arithmetic: one FFMA2 / FMUL2 / FADD2 instruction performs
// the operation on both halves of a register pair (PTX fma/mul/add.f32x2).  The FMA pipe
// throughput in FLOP/s is unchanged, but the ISSUE cost per FLOP halves, which is what bounds
// this engine (SASS of the ensemble kernels is issue-bound, not pipe-bound: see DESIGN.md).
// Instantiating the generated stage code and the model RHS with T = Pk2 therefore advances two
// trajectories per thread with half the instructions per trajectory.
//
// ptxas fuses mul.f32x2 + add.f32x2 into FFMA2 whenever the product has a single use, and it
// does so even under --fmad=false and for .rn-qualified operands (checked with CUDA 12.9:
// cuobjdump shows FFMA2 for `mul.rn.f32x2; add.rn.f32x2`).  Un-fused bit-parity arithmetic can
// therefore not be expressed with packed operations, so Pk2 is used by the fast fp mode only;
// the strict mode stays scalar.
#pragma once
#include "degk_common.cuh"

namespace degk {

struct Pk2 {
    unsigned long long v;
    DEGK_DEV Pk2() {}
    DEGK_DEV Pk2(float a, float b) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b)); }
    DEGK_DEV explicit Pk2(float a) { asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(a)); }
    DEGK_DEV explicit Pk2(double a) { const float f = (float)a; asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(f)); }
    DEGK_DEV explicit Pk2(int a) { const float f = (float)a; asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(f)); }
    DEGK_DEV float lo() const { float a; asm("mov.b64 {%0, _}, %1;" : "=f"(a) : "l"(v)); return a; }
    DEGK_DEV float hi() const { float b; asm("mov.b64 {_, %0}, %1;" : "=f"(b) : "l"(v)); return b; }
    DEGK_DEV float get(int s) const { return s ? hi() : lo(); }
    // math functions a model body may call, per half.  Hidden friends: found by argument-dependent
    // lookup only, so they do not hide ::sin / ::log ... for the scalar code in this namespace
    friend DEGK_DEV Pk2 sqrt(Pk2 a) { return Pk2(sqrtf(a.lo()), sqrtf(a.hi())); }
    friend DEGK_DEV Pk2 sin(Pk2 a) { return Pk2(sinf(a.lo()), sinf(a.hi())); }
    friend DEGK_DEV Pk2 cos(Pk2 a) { return Pk2(cosf(a.lo()), cosf(a.hi())); }
    friend DEGK_DEV Pk2 exp(Pk2 a) { return Pk2(expf(a.lo()), expf(a.hi())); }
    friend DEGK_DEV Pk2 log(Pk2 a) { return Pk2(logf(a.lo()), logf(a.hi())); }
};

DEGK_DEV Pk2 operator*(Pk2 a, Pk2 b) { Pk2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEGK_DEV Pk2 operator+(Pk2 a, Pk2 b) { Pk2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEGK_DEV Pk2 operator-(Pk2 a, Pk2 b) { Pk2 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEGK_DEV Pk2 operator-(Pk2 a) { Pk2 r; r.v = a.v ^ 0x8000000080000000ull; return r; }
DEGK_DEV Pk2 fma_(Pk2 a, Pk2 b, Pk2 c) {
    Pk2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
}
// division / abs / sqrt per half (fast mode only: MUFU reciprocal, one packed multiply)
DEGK_DEV Pk2 operator/(Pk2 a, Pk2 b) { return a * Pk2(rcp_(b.lo()), rcp_(b.hi())); }
DEGK_DEV Pk2 abs_(Pk2 a) { Pk2 r; r.v = a.v & 0x7fffffff7fffffffull; return r; }
DEGK_DEV Pk2 sqrt_(Pk2 a) { return Pk2(sqrtf(a.lo()), sqrtf(a.hi())); }

// per-half select: s0 ? a.lo : b.lo , s1 ? a.hi : b.hi
DEGK_DEV Pk2 select2(bool s0, bool s1, Pk2 a, Pk2 b) { return Pk2(s0 ? a.lo() : b.lo(), s1 ? a.hi() : b.hi()); }

DEGK_DEV Pk2 blendm(unsigned m, Pk2 a, Pk2 b) { return select2((m & 1u) != 0, (m & 2u) != 0, a, b); }

// In-place per-slot assignment `if (f[s]) x.slot(s) = y.slot(s)`: one predicated move per slot.
// (Written as x = blendm(...) the compiler builds the selected pair in fresh registers and copies
//  it back into the loop-carried pair: three instructions per slot instead of one.)
DEGK_DEV void assign_if(const bool* f, float& x, float y)   { x = f[0] ? y : x; }
DEGK_DEV void assign_if(const bool* f, double& x, double y) { x = f[0] ? y : x; }
DEGK_DEV void assign_if(const bool* f, Pk2& x, Pk2 y) {
    asm("{\n\t.reg .pred p0, p1;\n\t.reg .b32 a0, a1, b0, b1;\n\t"
        "mov.b64 {a0, a1}, %0;\n\tmov.b64 {b0, b1}, %1;\n\t"
        "setp.ne.u32 p0, %2, 0;\n\tsetp.ne.u32 p1, %3, 0;\n\t"
        "@p0 mov.b32 a0, b0;\n\t@p1 mov.b32 a1, b1;\n\tmov.b64 %0, {a0, a1};\n\t}"
        : "+l"(x.v) : "l"(y.v), "r"((unsigned)f[0]), "r"((unsigned)f[1]));
}

// PackOf<T, W>: the value type that carries W trajectories of element type T
template <class T, int W> struct PackOf;
template <class T> struct PackOf<T, 1> {
    typedef T type;
    static DEGK_DEV T get(T v, int) { return v; }
    static DEGK_DEV T make(const T (&s)[1]) { return s[0]; }
    static DEGK_DEV T set(T, int, T x) { return x; }
};
template <> struct PackOf<float, 2> {
    typedef Pk2 type;
    static DEGK_DEV float get(Pk2 v, int s) { return s ? v.hi() : v.lo(); }
    static DEGK_DEV Pk2 make(const float (&s)[2]) { return Pk2(s[0], s[1]); }
    static DEGK_DEV Pk2 set(Pk2 v, int s, float x) { return s == 0 ? Pk2(x, v.hi()) : Pk2(v.lo(), x); }
};

}  // namespace degk
