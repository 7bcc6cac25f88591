// sora_b200 — the legacy 802.11b transmit filter on the device: BB11BPMDSpreadFIR4SSE / BB11BPMDSpreadFIR4ASM (kernel/inc/bb/bbb.h:188-200,
// kernel/bb/dot11b/bbb_fir.c), the 37-tap pulse shaper BB11BPMDPacketGenSignal (bbb_tx.c:116-150) runs over the 4x zero-stuffed chip stream.
//
// The reference is a transposed-form filter: a group of four inputs adds its products to the partial sums of the 40 outputs it reaches, with
// saturating 16-bit adds, and a finished output is the saturating sum of its four lanes, (l0 + l2) + (l1 + l3), >> 8, packed to int8.  Read the
// other way round, output sample 4p + r of a frame is a fixed function of the ten input groups p-9 .. p (group q = input samples 8 + 4q .. 8 + 4q + 3:
// the routine starts at the second 16-byte block): lane i sums x[8 + 4q + i] * h[r + 4(p - q) - i] over q.  A lane's sum cannot saturate
// (it meets every fourth tap only: at most 195 * 128 = 24 960), so only the three adds of the lane tree need the clamp, and every output is
// independent: one thread per group of four outputs, ten 8-byte loads, ~80 multiply-adds with compile-time coefficients.
// The two bodies of the reference differ in the outer +-1 taps (legacy 802.11b transmit filter, has the why):
//   variant 0 (FIR37SSE_INTRINSIC, bbb_fir.c:413-566): in an even group outputs 1..3 take x[8+4p] * 1 as their newest contribution (row 0's
//     coefficients instead of rows 1..3); the oldest contribution of output r = 1 is missing and that of r = 2 is -x[2] of the even group at or
//     before p - 9 (instead of +x[2] of group p - 9);
//   variant 1 (FIR37SSE_INLINE, bbb_fir.c:137-386): the plain filter.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb {

__host__ __device__ constexpr int fir37_h(int k) {
    constexpr int H[37] = {1, 0, -1, 0, 1, 0, -1, 0, 2, 0, -3, 0, 5, 0, -11, 0, 54, 128, 163, 128, 54, 0, -11, 0, 5, 0, -3, 0, 2, 0, -1, 0, 1, 0, -1, 0, 1};
    return (k >= 0 && k < 37) ? H[k] : 0;
}
__device__ __forceinline__ int fir37_sat16(int v) { return min(max(v, -32768), 32767); }

#define SB_FIR37_THREADS 128

// grid: x = frame, y = tiles of SB_FIR37_THREADS groups (grid-stride over the frame's groups)
template <int VARIANT>
__global__ void __launch_bounds__(SB_FIR37_THREADS) k_fir37_legacy(const int8_t* __restrict__ in, const uint64_t* __restrict__ off, const uint32_t* __restrict__ len,
                                                                  uint32_t nframes, int8_t* __restrict__ out) {
    const uint32_t f = blockIdx.x;
    if (f >= nframes) return;
    const uint32_t n_in = len[f], ngroups = (n_in >> 3) * 2u;
    const uint2* src = reinterpret_cast<const uint2*>(in + 2ull * off[f]);          // one group of four COMPLEX8 per 8-byte word (off is a multiple of 8 samples)
    uint2* dst = reinterpret_cast<uint2*>(out + 2ull * off[f]);
    const uint32_t nwords = n_in >> 2;                                               // whole groups of the input; samples past n_in read as zero
    for (uint32_t p = blockIdx.y * SB_FIR37_THREADS + threadIdx.x; p < ngroups; p += gridDim.y * SB_FIR37_THREADS) {
        int v[10][4][2];                                                             // v[m] = group p - m
#pragma unroll
        for (int m = 0; m < 10; m++) {
            const int64_t w = (int64_t)p - m + 2;                                    // word index of group q = p - m (sample 8 + 4q)
            uint2 x = make_uint2(0u, 0u);
            if (p >= (uint32_t)m && (uint64_t)w < nwords) x = __ldg(src + w);
            v[m][0][0] = (int8_t)(x.x); v[m][0][1] = (int8_t)(x.x >> 8); v[m][1][0] = (int8_t)(x.x >> 16); v[m][1][1] = (int8_t)(x.x >> 24);
            v[m][2][0] = (int8_t)(x.y); v[m][2][1] = (int8_t)(x.y >> 8); v[m][3][0] = (int8_t)(x.y >> 16); v[m][3][1] = (int8_t)(x.y >> 24);
        }
        int stale2[2] = {0, 0};                                                       // variant 0: x[2] of the even group at or before p - 9
        if (VARIANT == 0) {
            const uint32_t odd = (p + 1u) & 1u;                                       // p - 9 is odd when p is even
            if (!odd) { stale2[0] = v[9][2][0]; stale2[1] = v[9][2][1]; }
            else if (p >= 10u) {
                const int64_t w = (int64_t)p - 10 + 2; uint2 x = make_uint2(0u, 0u);
                if ((uint64_t)w < nwords) x = __ldg(src + w);
                stale2[0] = (int8_t)(x.y); stale2[1] = (int8_t)(x.y >> 8);
            }
        }
        const bool have9 = p >= 9u;                                                   // before that the partial sums are still the zeros they started as
        int y[4][2];
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                int l[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    int s = 0;
#pragma unroll
                    for (int m = 1; m < 9; m++) s += v[m][i][c] * fir37_h(r + 4 * m - i);
                    // oldest contribution (rows 36..39 of the table)
                    if (VARIANT == 1 || r == 0 || r == 3) s += v[9][i][c] * fir37_h(r + 36 - i);
                    else if (r == 2 && i == 2) s += have9 ? -stale2[c] : 0;             // row 38 made from the row-36 product of the even group: -x[2]
                    // newest contribution
                    if (VARIANT == 0 && r > 0) s += ((p & 1u) == 0) ? v[0][i][c] * fir37_h(0 - i) : v[0][i][c] * fir37_h(r - i);
                    else s += v[0][i][c] * fir37_h(r - i);
                    l[i] = s;
                }
                const int t = fir37_sat16(fir37_sat16(l[0] + l[2]) + fir37_sat16(l[1] + l[3])) >> 8;
                y[r][c] = min(max(t, -128), 127) & 0xFF;
            }
        }
        dst[p] = make_uint2((uint32_t)y[0][0] | ((uint32_t)y[0][1] << 8) | ((uint32_t)y[1][0] << 16) | ((uint32_t)y[1][1] << 24),
                            (uint32_t)y[2][0] | ((uint32_t)y[2][1] << 8) | ((uint32_t)y[3][0] << 16) | ((uint32_t)y[3][1] << 24));
    }
}

}  // namespace sb
