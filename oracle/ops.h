// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sora_b200/ may include, link or call this.
//
// Lane-exact SSE building blocks for the CPU restatement of Sora's fixed-point DSP.
// Every helper states which reference primitive's *semantics* it reproduces
// (kernel/core/inc/vector128.h, cited by line); the code itself is written from scratch
// directly on <immintrin.h> intrinsics so that the hardware defines the lane arithmetic.
#pragma once
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

namespace sbo {

struct c16 { int16_t re, im; };          // COMPLEX16: re in the low half of each 32-bit lane
typedef __m128i v128;                    // 4 x c16  ("vcs")  or 4 x int32 ("vi") or 16 x uint8 ("vub")

static inline v128 ld(const void* p) { return _mm_loadu_si128((const __m128i*)p); }
static inline void st(void* p, v128 v) { _mm_storeu_si128((__m128i*)p, v); }

// swap re<->im inside every complex lane           (vector128.h:992-996 flip)
static inline v128 swap_ri(v128 a) { return _mm_shufflehi_epi16(_mm_shufflelo_epi16(a, 0xb1), 0xb1); }
// exact negate of the real parts (psignw)          (vector128.h:957-967 conjre)
static inline v128 neg_re(v128 a) { return _mm_sign_epi16(a, _mm_set1_epi32(0x00018000)); }
// exact negate of the imaginary parts (psignw)     (vector128.h:982-992 conj0)
static inline v128 neg_im(v128 a) { return _mm_sign_epi16(a, _mm_set1_epi32((int)0x80000001)); }
// one's-complement "conjugate" (im -> ~im)          (vector128.h:971 conj)
static inline v128 conj_xor(v128 a) { return _mm_xor_si128(a, _mm_set1_epi32((int)0xFFFF0000)); }
// approximate multiply by j: (re,im) -> (~im, re)  (vector128.h:1258-1261 mul_j)
static inline v128 mulj_xor(v128 a) { return _mm_xor_si128(swap_ri(a), _mm_set1_epi32(0x0000FFFF)); }

// a * conj(b) as separate int32 re / im vectors    (vector128.h:1031-1037 conj_mul)
static inline void cmul_conj32(v128& re, v128& im, v128 a, v128 b) {
    v128 t = neg_re(swap_ri(b));               // (-b.im, b.re)
    re = _mm_madd_epi16(a, b);                 // a.re*b.re + a.im*b.im
    im = _mm_madd_epi16(t, a);                 // -b.im*a.re + b.re*a.im
}
// a * b as separate int32 re / im vectors          (vector128.h:1072-1078 mul)
static inline void cmul32(v128& re, v128& im, v128 a, v128 b) {
    re = _mm_madd_epi16(a, neg_im(b));         // a.re*b.re - a.im*b.im
    im = _mm_madd_epi16(a, swap_ri(b));        // a.re*b.im + a.im*b.re
}
// truncating repack of int32 re/im -> c16          (vector128.h:876-893 pack)
static inline v128 pack_trunc(v128 re, v128 im) {
    return _mm_or_si128(_mm_and_si128(re, _mm_set1_epi32(0xFFFF)), _mm_slli_epi32(im, 16));
}
// Q15 complex product, truncating pack             (vector128.h:1199-1211 mul(vcs,vcs))
static inline v128 cmul_q15(v128 a, v128 b) {
    v128 re, im; cmul32(re, im, a, b);
    return pack_trunc(_mm_srai_epi32(re, 15), _mm_srai_epi32(im, 15));
}
// FFT twiddle product: a * b >> n with the xor-conjugate   (vector128.h:1235-1246 mul_shift)
static inline v128 cmul_shift_fft(v128 a, v128 b, int n) {
    v128 re = _mm_madd_epi16(a, conj_xor(b));
    v128 im = _mm_madd_epi16(a, swap_ri(b));
    return pack_trunc(_mm_srai_epi32(re, n), _mm_srai_epi32(im, n));
}
// IFFT twiddle product: a * conj(b) >> n            (vector128.h:1215-1231 conj_mul_shift)
static inline v128 cmul_conj_shift(v128 a, v128 b, int n) {
    v128 t = swap_ri(neg_re(a));               // (a.im, -a.re)
    v128 re = _mm_madd_epi16(a, b);
    v128 im = _mm_madd_epi16(t, b);
    return pack_trunc(_mm_srai_epi32(re, n), _mm_srai_epi32(im, n));
}
// |x|^2 per complex lane as int32                   (vector128.h:933 SquaredNorm)
static inline v128 norm2(v128 a) { return _mm_madd_epi16(a, a); }
// wrapping sum of the 4 int32 lanes, lane 0         (vector128.h:378-383,758 hadd(vi))
static inline int32_t hsum32(v128 a) {
    v128 t = _mm_add_epi32(a, _mm_shuffle_epi32(a, 0xb1));
    t = _mm_add_epi32(t, _mm_shuffle_epi32(t, 0x4e));
    return _mm_cvtsi128_si32(t);
}
// wrapping int16 sum of the 4 complex lanes, broadcast to all   (vector128.h:757 hadd(vcs))
static inline v128 hsum_c16(v128 a) {
    v128 t = _mm_add_epi16(a, _mm_shuffle_epi32(a, 0xb1));
    return _mm_add_epi16(t, _mm_shuffle_epi32(t, 0x4e));
}
static inline c16 lane0(v128 a) { uint32_t w = (uint32_t)_mm_cvtsi128_si32(a); c16 c; memcpy(&c, &w, 4); return c; }
static inline v128 splat(c16 c) { uint32_t w; memcpy(&w, &c, 4); return _mm_set1_epi32((int)w); }

} // namespace sbo
