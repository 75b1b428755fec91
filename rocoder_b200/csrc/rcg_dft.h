// rcg_dft.h -- complex arithmetic on (re, im) pairs and fully unrolled radix-2/4/8/16 DFTs on register arrays.
//
// A complex value is one cf = (re, im) pair in an aligned 64-bit register pair.  Every operation exists in two
// forms: scalar (two FADD / FMUL / FFMA) and Blackwell's packed FP32 (one FADD2 / FMUL2 / FFMA2 on both halves;
// ptxas folds a swapped operand (im, re) and a scalar broadcast (s, s) into the instruction's operand modifiers
// .LO_HI / .F32, so a complex butterfly with a constant twiddle costs 4 instructions instead of 8 and a radix-16 DFT
// 84 instead of 168).  A packed instruction occupies the whole FMA pipe for two cycles, i.e. the pipe sees the same
// work and only the issue slots are halved (scripts/microbench/f32x2_issue.cu measures the rates) -- see
// RCG_PACKED_FP32 below for what that bought.
// The host build (tests/emul) executes the same code with scalar arithmetic.
#pragma once
#include "rcg_common.h"

namespace rcg {

RCG_HD cf pk_add(cf a, cf b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return make_cf(a.x + b.x, a.y + b.y);
#endif
}
RCG_HD cf pk_neg(cf a) { return make_cf(-a.x, -a.y); }   // folded into the consumer's operand modifier
RCG_HD cf pk_sub(cf a, cf b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, pk_neg(b));
#else
    return make_cf(a.x - b.x, a.y - b.y);
#endif
}
RCG_HD cf pk_mul(cf a, cf b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return make_cf(a.x * b.x, a.y * b.y);
#endif
}
// a * b + c per half; fused on the device (one rounding), two roundings on the host build
RCG_HD cf pk_fma(cf a, cf b, cf c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return make_cf(a.x * b.x + c.x, a.y * b.y + c.y);
#endif
}
RCG_HD cf pk_swap(cf a) { return make_cf(a.y, a.x); }     // (im, re): the .LO_HI operand modifier
RCG_HD cf pk_bcast(float s) { return make_cf(s, s); }     // scalar operand, broadcast by the instruction
RCG_HD cf pk_scale(cf a, float s) { return pk_mul(a, pk_bcast(s)); }

// Per-half multiply / add that no compiler stage may contract: ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into one
// FFMA2 although both carry .rn (it honours .rn on the scalar forms), so wherever the reference's rounding order is
// part of the contract -- window * sample, (y + carried) -- the operation next to the add is a scalar one.
RCG_HD cf ex_mul(cf a, cf b) {
#if defined(__CUDA_ARCH__)
    return make_cf(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#else
    return make_cf(a.x * b.x, a.y * b.y);
#endif
}
RCG_HD cf ex_add(cf a, cf b) {
#if defined(__CUDA_ARCH__)
    return make_cf(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
#else
    return make_cf(a.x + b.x, a.y + b.y);
#endif
}
// a product that only ever feeds another product (never an add): packed where the phase allows it
template <bool PK>
RCG_HD cf mul_chain(cf a, cf b) { if constexpr (PK) return pk_mul(a, b); else return ex_mul(a, b); }

// The same operations spelled per half with plain operators: scalar FADD / FMUL / FFMA (contracted by nvcc where it
// sees a * b + c).  Which form a phase of the frame uses is decided per phase (RCG_PK_*): a packed instruction takes
// BOTH halves of the FMA pipe for two cycles, so it cannot overlap the IMAD.WIDE of the Philox rounds the way scalar
// FP32 (which fits on the light half alone) can -- measured, scripts/microbench/f32x2_issue.cu.
// RCG_PACKED_FP32 = 1 builds the packed forms everywhere (measured on B200, profiles/r2e_packed_fp32_variants.md:
// 17 % fewer warp instructions, issue slots 67 % -> 55 % busy, and NOT faster -- 6.57 vs 6.46 ms on config 2: the
// FMA pipe's cycles, not the issue slots, bound the math phases, and `math pipe throttle` becomes the top stall).
// The default is therefore the scalar form; the per-phase switches exist for that A/B only.
#ifndef RCG_PACKED_FP32
#define RCG_PACKED_FP32 0
#endif
#ifndef RCG_PK_DFT
#define RCG_PK_DFT RCG_PACKED_FP32    // radix-2/4/8/16 butterflies
#endif
#ifndef RCG_PK_TW
#define RCG_PK_TW RCG_PACKED_FP32     // products with pass twiddles
#endif
#ifndef RCG_PK_SPEC
#define RCG_PK_SPEC RCG_PACKED_FP32   // spectral step (shares its phase with the Philox rounds)
#endif
#ifndef RCG_PK_IO
#define RCG_PK_IO RCG_PACKED_FP32     // product chains of the load / overlap-add phases
#endif
// constant-twiddle butterflies of the radix-8/16 DFTs in the factored form w o = c (o + (s/c) rot(o)) (bfly16)
#ifndef RCG_BFLY_FACTORED
#define RCG_BFLY_FACTORED 1
#endif
template <bool PK>
struct Ar {
    static RCG_HD cf add(cf a, cf b) { if constexpr (PK) return pk_add(a, b); else return make_cf(a.x + b.x, a.y + b.y); }
    static RCG_HD cf sub(cf a, cf b) { if constexpr (PK) return pk_sub(a, b); else return make_cf(a.x - b.x, a.y - b.y); }
    static RCG_HD cf mul(cf a, cf b) { if constexpr (PK) return pk_mul(a, b); else return make_cf(a.x * b.x, a.y * b.y); }
    static RCG_HD cf fma(cf a, cf b, cf c) { if constexpr (PK) return pk_fma(a, b, c); else return make_cf(a.x * b.x + c.x, a.y * b.y + c.y); }
    static RCG_HD cf scale(cf a, float s) { return mul(a, pk_bcast(s)); }
};

// ---- complex products with a run-time factor ----------------------------------------------------------------
// x * w = x.re * (w.re, w.im) + x.im * (-w.im, w.re): with the rotated copy iw = i w = (-w.im, w.re) at hand a
// complex product is two packed instructions.  A twiddle that multiplies several values keeps its rotated copy.
struct Tw {
    cf w;    // (re, im)
    cf iw;   // i * w = (-im, re)
};
template <bool PK = (RCG_PK_TW != 0)>
RCG_HD Tw make_tw(cf w) {
    Tw t;
    t.w = w;
    if constexpr (PK) t.iw = pk_mul(pk_swap(w), make_cf(-1.0f, 1.0f));
    else t.iw = make_cf(-w.y, w.x);   // scalar form: the negation folds into the consumers
    return t;
}
// x * w;  CONJ: x * conj(w) = x.re * (w.re, -w.im) + x.im * (w.im, w.re) = x.re * swap(iw) + x.im * swap(w)
template <bool CONJ, bool PK = (RCG_PK_TW != 0)>
RCG_HD cf tw_mul(cf x, const Tw& t) {
    using A = Ar<PK>;
    if constexpr (CONJ) return A::fma(pk_bcast(x.y), pk_swap(t.w), A::mul(pk_bcast(x.x), pk_swap(t.iw)));
    else return A::fma(pk_bcast(x.y), t.iw, A::mul(pk_bcast(x.x), t.w));
}
template <bool PK = (RCG_PK_TW != 0)>
RCG_HD Tw tw_product(const Tw& a, const Tw& b) { return make_tw<PK>(tw_mul<false, PK>(a.w, b)); }

// ---- one radix-2 butterfly with a constant twiddle ----------------------------------------------------------
// y = e + w o, z = e - w o, w = exp(-2 pi i K16 / 16) (INV: its conjugate), K16 in 0..7
template <int K16, bool INV>
RCG_HD void bfly16(cf e, cf o, cf& y, cf& z) {
    using A = Ar<(RCG_PK_DFT != 0)>;
    constexpr float C1 = 0.92387953251128674f;  // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;  // sin(pi/8)
    constexpr float R2 = 0.70710678118654752f;  // sqrt(1/2)
    if constexpr (K16 == 0) {
        y = A::add(e, o); z = A::sub(e, o);
    } else if constexpr (K16 == 4) {
        // w = -i: w o = (o.im, -o.re)   (INV: w = +i: (-o.im, o.re))
        const cf s = INV ? make_cf(-1.0f, 1.0f) : make_cf(1.0f, -1.0f);
        y = A::fma(pk_swap(o), s, e); z = A::fma(pk_swap(o), pk_neg(s), e);
    } else {
        constexpr float c = (K16 == 1) ? C1 : (K16 == 2) ? R2 : (K16 == 3) ? S1 : (K16 == 5) ? -S1 : (K16 == 6) ? -R2 : -C1;
        constexpr float s = (K16 == 1 || K16 == 7) ? S1 : (K16 == 2 || K16 == 6) ? R2 : C1;   // sin(K16 pi / 8) > 0
        // w = c - i s: w o = c o + s (o.im, -o.re)   (INV: w = c + i s: c o + s (-o.im, o.re))
        if constexpr (RCG_BFLY_FACTORED && RCG_PK_DFT == 0) {
            // w o = c u with u = o + (s/c) (o.im, -o.re): two FFMA for u (two FADD when s = +-c), then y = e + c u and
            // z = e - c u are two FFMA each -- 6 instructions instead of 8 (c is never 0 here: K16 = 4 is the case above)
            constexpr float r = s / c;
            const cf rv = INV ? make_cf(-r, r) : make_cf(r, -r);
            const cf u = A::fma(pk_swap(o), rv, o);
            y = A::fma(u, pk_bcast(c), e); z = A::fma(u, pk_bcast(-c), e);
        } else {
            const cf sv = INV ? make_cf(-s, s) : make_cf(s, -s);
            const cf t = A::fma(pk_swap(o), sv, A::scale(o, c));
            y = A::add(e, t); z = A::sub(e, t);
        }
    }
}

// y[k] = sum_a x[a] exp(-+2 pi i a k / R), in place; INV = unnormalised inverse (conjugate kernel)
template <int R, bool INV = false>
struct Dft {
    static RCG_HD void run(cf (&x)[R]) {
        constexpr int Hh = R / 2;
        cf e[Hh], o[Hh];
#pragma unroll
        for (int a = 0; a < Hh; a++) { e[a] = x[2 * a]; o[a] = x[2 * a + 1]; }
        Dft<Hh, INV>::run(e);
        Dft<Hh, INV>::run(o);
        combine<0>(e, o, x);
    }
    template <int K>
    static RCG_HD void combine(cf (&e)[R / 2], cf (&o)[R / 2], cf (&y)[R]) {
        if constexpr (K < R / 2) {
            bfly16<K * (16 / R), INV>(e[K], o[K], y[K], y[K + R / 2]);
            combine<K + 1>(e, o, y);
        }
    }
};
template <bool INV>
struct Dft<1, INV> {
    static RCG_HD void run(cf (&)[1]) {}
};

}  // namespace rcg
