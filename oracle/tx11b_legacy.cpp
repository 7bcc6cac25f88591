// sora_b200 oracle — TEST INFRASTRUCTURE ONLY (never on the product path).
// The legacy 802.11b transmit filter: BB11BPMDSpreadFIR4SSE / BB11BPMDSpreadFIR4ASM (kernel/inc/bb/bbb.h:188-200), i.e. the 37-tap
// pulse-shaping FIR of kernel/bb/dot11b/bbb_fir.c that BB11BPMDPacketGenSignal (bbb_tx.c:116-150) runs over the 4x zero-stuffed chip stream.
//
// What the reference computes, restated as a transposed-form filter over groups of four complex int8 samples:
//   * the coefficient table (bbb_fir.c:21-63) is row j, lane i -> h[j - i] for the 37 taps h = 1 0 -1 0 1 0 -1 0 2 0 -3 0 5 0 -11 0 54 128 163 128 54 ...
//     (symmetric); 40 rows, so one group of four inputs updates the partial sums of the 40 outputs it can reach;
//   * the routine starts reading at the SECOND 16-byte block of the source (bbb_fir.c:428 / :170 `[esi + 16]`): output sample n is the filter
//     response aligned to input sample n + 8, the first eight inputs never enter, and eight samples past uiInputSize are read (the caller
//     zeroes 64 bytes of tail, bbb_tx.c:140); here samples beyond n_in read as zero;
//   * every accumulation is a saturating 16-bit add (paddsw): per lane oldest contribution first, then the four lanes as (l0 + l2) + (l1 + l3);
//   * outputs are >> 8 (arithmetic) and packed to int8 with signed saturation;
//   * variant 0 = the intrinsic body FIR37SSE_INTRINSIC (bbb_fir.c:413-566, what BB11BPMDSpreadFIR4SSE calls): rows 37 and 38 of the fresh
//     partial sums are formed from the ALREADY MULTIPLIED row-36 product of the low half-block (bbb_fir.c:485-491, :553-559), so row 37 is
//     always zero and row 38 carries -x[2] of the low group instead of +x[2] of the current one; and the helper macro that forms outputs
//     1..3 of the LOW half-block names its coefficient argument `pTags` but reads `pTaps` (bbb_fir.c:390-395), so those three outputs take
//     x[0] * 1 (row 0) where rows 1..3 were meant.  All of it sits in the +-1 outer taps; oracle/_ref (the compiled reference body,
//     oracle/build_ref.sh) is what this restatement is checked against, sample for sample;
//     variant 1 = the inline-assembly body FIR37SSE_INLINE (bbb_fir.c:137-386, BB11BPMDSpreadFIR4ASM, 32-bit builds), which multiplies the
//     input itself in all four rows.
// Pin status: unpinned by a reference output.  The reference's 802.11b sample files (kernel/HWTest/exe/tx samples/*.mf.bin) were shaped by a
// longer, non-integer filter (a least-squares fit of the taps from the files gives 0.46 -0.60 0.45 -1.20 1.38 -2.57 4.13 -10.94 53.6 127.4 161.9 ...
// with a ripple that runs past 37 taps); this filter follows them within +-3 LSB (tests/test_cpu_oracle_tx11b.py).
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace sbo {

static const int kH[37] = {1, 0, -1, 0, 1, 0, -1, 0, 2, 0, -3, 0, 5, 0, -11, 0, 54, 128, 163, 128, 54, 0, -11, 0, 5, 0, -3, 0, 2, 0, -1, 0, 1, 0, -1, 0, 1};
static inline int coef(int row, int lane) { const int k = row - lane; return (k >= 0 && k < 37) ? kH[k] : 0; }
static inline int sat16(int v) { return v > 32767 ? 32767 : v < -32768 ? -32768 : v; }
static inline int mul16(int a, int b) { return (int)(int16_t)(a * b); }          // pmullw

void fir37_legacy(const int8_t* src, uint32_t n_in, int variant, int8_t* dst) {
    int T[40][4][2]; memset(T, 0, sizeof T);                                     // partial sums of the next 40 outputs: [row][lane][re / im]
    int stale[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};                          // row-36 product of the last LOW half-block
    auto in = [&](uint64_t i, int c) -> int { return i < n_in ? (int)src[2 * i + c] : 0; };
    const uint32_t ngroups = (n_in / 8u) * 2u;
    for (uint32_t p = 0; p < ngroups; p++) {
        int v[4][2];
        for (int i = 0; i < 4; i++) for (int c = 0; c < 2; c++) v[i][c] = in(8ull + 4ull * p + i, c);
        // four finished outputs
        for (int r = 0; r < 4; r++) for (int c = 0; c < 2; c++) {
            int l[4];
            const int row = (variant == 0 && (p & 1u) == 0) ? 0 : r;          // variant 0, low half-block: rows 1..3 are multiplied by row 0 (see header)
            for (int i = 0; i < 4; i++) l[i] = sat16(mul16(v[i][c], coef(row, i)) + T[r][i][c]);
            const int y = sat16(sat16(l[0] + l[2]) + sat16(l[1] + l[3])) >> 8;
            dst[2 * (4ull * p + r) + c] = (int8_t)std::min(127, std::max(-128, y));
        }
        // partial sums move up by four rows and take this group's contribution
        for (int j = 0; j < 32; j++) for (int i = 0; i < 4; i++) for (int c = 0; c < 2; c++) T[j][i][c] = sat16(mul16(v[i][c], coef(j + 4, i)) + T[j + 4][i][c]);
        int p36[4][2];
        for (int i = 0; i < 4; i++) for (int c = 0; c < 2; c++) p36[i][c] = mul16(v[i][c], coef(36, i));
        if ((p & 1u) == 0) memcpy(stale, p36, sizeof stale);
        for (int i = 0; i < 4; i++) for (int c = 0; c < 2; c++) {
            T[32][i][c] = p36[i][c];
            T[33][i][c] = variant == 0 ? mul16(stale[i][c], coef(37, i)) : mul16(v[i][c], coef(37, i));
            T[34][i][c] = variant == 0 ? mul16(stale[i][c], coef(38, i)) : mul16(v[i][c], coef(38, i));
            T[35][i][c] = mul16(v[i][c], coef(39, i));
        }
    }
}

}  // namespace sbo

extern "C" void sbo_fir37_legacy(const int8_t* src, uint32_t n_in, int variant, int8_t* dst) { sbo::fir37_legacy(src, n_in, variant, dst); }
