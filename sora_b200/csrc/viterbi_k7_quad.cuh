// sora_b200 — batched K=7 (133,171) soft Viterbi, v2 "quad" mapping for sm_100a.
//
// Same arithmetic contract as viterbi_k7.cuh (bit-exact with kernel/bb/Brick11/src/viterbicore.h:269-556 driven like
// kernel/bb/Brick11/src/viterbi.hpp:104-237), different machine mapping:
//
//   * 4 lanes decode one code block; each lane keeps 16 of the 64 path metrics in 8 registers, two per register as
//     16-bit halves, so add / compare-select run on Blackwell's native 16x2 integer SIMD (VIADD.16x2, VIMNMX.U16x2);
//     the reference's uint8 wrap is "& 0x00FE00FE" / "(& 0x00FF00FF) | 0x00010001" folded into one LOP3 each.
//   * The trellis is processed IN PLACE: the butterfly (p, p+32) -> (2p, 2p+1) writes its two results into the two
//     slots its inputs came from.  A state index is therefore a 6-bit rotation of its physical address
//     A = lane[2] | reg[3] | half[1]:  state(A, t) = rol6(A, t mod 6).  The pairing dimension walks through the address
//     bits with period 6:  t%6 = 0,1 -> partner lane (one __shfl_xor per register: the path-metric exchange),
//     2,3,4 -> another register of the same lane (pure SIMD, no data movement), 5 -> the two halves of one register.
//   * Branch metrics: the 4 possible values of a step live as bytes of one register; each 2-state operand is one
//     PRMT with a compile-time selector (lane-dependent part applied once per step by a second PRMT).
//   * Survivor bits: the LSB of every new metric, 16 per lane per step, stored as one 16-bit word into a
//     [column][block-in-CTA] shared-memory ring (conflict-free, 8 B per code block per step); traceback reads
//     bit ror6(state, column mod 6) of the 64-bit column word.
//   * One lane of the quad runs the windowed traceback, the x^7+x^4+1 descrambler and the CRC-32 / verdict.
#pragma once
#include "viterbi_k7.cuh"

namespace sb {

#define SB_VQ_WARPS 2                      // warps per CTA
#define SB_VQ_FR (8 * SB_VQ_WARPS)         // code blocks per CTA
#define SB_VQ_RING 304                     // columns kept per code block (>= depth + lookahead + 7 + 3 + slack)

__host__ __device__ constexpr int vq_rol6(int a, int t) { return ((a << t) | (a >> (6 - t))) & 63; }
__host__ __device__ constexpr int vq_cls(int p) {            // (cA << 1) | cB of predecessor index p (bit 5 ignored)
    return ((((p >> 1) ^ (p >> 2) ^ (p >> 4)) & 1) << 1) | ((p ^ (p >> 1) ^ (p >> 2)) & 1);
}
// static class of (reg r, half h) at phase T, lane part excluded (GF(2)-linear, so the lane part is XORed in later)
__host__ __device__ constexpr int vq_scls(int T, int r, int h) { return vq_cls(vq_rol6((r << 1) | h, T) & 31); }
__host__ __device__ constexpr int vq_lcls(int T, int q) { return vq_cls(vq_rol6(q << 4, T) & 31); }
// PRMT selector building [byte i0, 0, byte i1, 0] from (Cb, 0)
__host__ __device__ constexpr unsigned vq_sel(int i0, int i1) { return (unsigned)(i0 | (4 << 4) | (i1 << 8) | (4 << 12)); }

struct VqLane {
    unsigned swz[6];       // per-phase byte swizzle applying this lane's class contribution
    unsigned m1[2], o1[2], m2[2], o2[2];   // lane-phase masks: which of (own, partner) is the even branch
};

// one trellis step at compile-time phase T.  Cbase bytes = metric for class (a,b) at byte (a<<1|b) for the even branch
// of the *upper* output (new state with input bit 0); complement class = 3 - index.
template <int T>
__device__ __forceinline__ void vq_step(uint32_t (&R)[8], uint32_t Cbase, const VqLane& L, unsigned qmask) {
    const uint32_t Cb = __byte_perm(Cbase, 0, L.swz[T]);
    const uint32_t EV = 0x00FE00FEu, FF = 0x00FF00FFu, ONE = 0x00010001u;
    if (T <= 1) {                                       // pair = partner lane (xor 2 at T=0, xor 1 at T=1)
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int c0 = vq_scls(T, r, 0), c1 = vq_scls(T, r, 1);
            uint32_t av = __byte_perm(Cb, 0, vq_sel(c0, c1)), bv = __byte_perm(Cb, 0, vq_sel(3 - c0, 3 - c1));
            uint32_t Z = __shfl_xor_sync(qmask, R[r], T == 0 ? 2 : 1);
            uint32_t t1 = R[r] + av, t2 = Z + bv;            // halves stay < 2^16: a plain 32-bit add is a 16x2 add
            // lane holding p (pair bit 0): out = min(t1 & EV, (t2 & FF) | 1);  lane holding p+32: roles swapped
            R[r] = __vminu2((t1 & L.m1[T]) | L.o1[T], (t2 & L.m2[T]) | L.o2[T]);
        }
    } else if (T <= 4) {                                // pair = register r ^ d inside the lane
        const int d = T == 2 ? 4 : T == 3 ? 2 : 1;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (r & d) continue;
            const int c0 = vq_scls(T, r, 0), c1 = vq_scls(T, r, 1);
            uint32_t av = __byte_perm(Cb, 0, vq_sel(c0, c1)), bv = __byte_perm(Cb, 0, vq_sel(3 - c0, 3 - c1));
            uint32_t X = R[r], Y = R[r + d];
            R[r]     = __vminu2((X + av) & EV, ((Y + bv) & FF) | ONE);
            R[r + d] = __vminu2((X + bv) & EV, ((Y + av) & FF) | ONE);
        }
    } else {                                            // pair = the two halves of each register
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int c = vq_scls(5, r, 0);             // class of p (low half); the high half is p+32: complement
            uint32_t ab = __byte_perm(Cb, 0, vq_sel(c, 3 - c)), ba = __byte_perm(Cb, 0, vq_sel(3 - c, c));
            uint32_t t1 = R[r] + ab, t2 = R[r] + ba;      // t1 = [p+a, p32+b], t2 = [p+b, p32+a]
            uint32_t lo = __byte_perm(t1, t2, 0x5410), hi = __byte_perm(t1, t2, 0x7632);   // [t1.lo, t2.lo], [t1.hi, t2.hi]
            R[r] = __vminu2(lo & EV, (hi & FF) | ONE);
        }
    }
}
__device__ __forceinline__ void vq_step_rt(int T, uint32_t (&R)[8], uint32_t Cbase, const VqLane& L, unsigned lane) {
    switch (T) { case 0: vq_step<0>(R, Cbase, L, lane); break; case 1: vq_step<1>(R, Cbase, L, lane); break;
                 case 2: vq_step<2>(R, Cbase, L, lane); break; case 3: vq_step<3>(R, Cbase, L, lane); break;
                 case 4: vq_step<4>(R, Cbase, L, lane); break; default: vq_step<5>(R, Cbase, L, lane); }
}
__device__ __forceinline__ uint32_t vq_decisions(const uint32_t (&R)[8]) {   // bit (2r+h) = LSB of half h of R[r]
    uint32_t acc = R[0] & 0x00010001u;
#pragma unroll
    for (int r = 1; r < 8; r++) acc |= (R[r] << (2 * r)) & (0x00010001u << (2 * r));
    return (acc | (acc >> 15)) & 0xFFFFu;
}
// branch-metric byte vectors (index = cA<<1 | cB)
__device__ __forceinline__ uint32_t vq_bm_ab(int sA, int sB) {
    int x = 2 * (sA + sB), y = 2 * (sA - sB) + 14;          // c00 = tA+tB, c01 = tA+14-tB, c10 = 28-c01, c11 = 28-c00
    return (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)(28 - y) << 16) | ((uint32_t)(28 - x) << 24);
}
__device__ __forceinline__ uint32_t vq_bm_a(int s) { uint32_t c0 = 2 * s, c1 = 14 - 2 * s; return c0 | (c0 << 8) | (c1 << 16) | (c1 << 24); }
__device__ __forceinline__ uint32_t vq_bm_b(int s) { uint32_t c0 = 2 * s, c1 = 14 - 2 * s; return c0 | (c1 << 8) | (c0 << 16) | (c1 << 24); }

template <int CODE_RATE>
__global__ void __launch_bounds__(32 * SB_VQ_WARPS) k_viterbi_quad(const uint8_t* __restrict__ soft, uint64_t soft_stride,
        uint32_t nframes, const FrameInfo* __restrict__ info, VitJob job, DevTables T,
        uint8_t* __restrict__ out, uint64_t out_stride, uint32_t* __restrict__ status_out, uint32_t* __restrict__ crc_out) {
    __shared__ unsigned long long s_ring[SB_VQ_RING][SB_VQ_FR];
    __shared__ uint32_t s_crc[256];
    __shared__ uint8_t s_scr[128];
    __shared__ uint8_t s_win[SB_VQ_FR][48];
    const int lane = threadIdx.x & 31, q = lane & 3;
    const unsigned QM = 0xFu << (lane & 28);           // the 4 lanes of this code block: quads run as independent sub-warps
    const int fb = (threadIdx.x >> 2);                 // code block within the CTA
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = __ldg(T.crc32 + i);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_scr[i] = __ldg(T.scramble + i);
    __syncthreads();
    const uint32_t f = blockIdx.x * SB_VQ_FR + fb;
    // per-frame parameters; frames of other rates (or finished/invalid) are skipped by this instantiation
    uint32_t L = job.frame_len, nsoft = job.nsoft; bool active = f < nframes;
    if (active && info) {
        FrameInfo fi = info[f];
        active = fi.status == E_SUCCESS && fi.code_rate == (uint32_t)CODE_RATE;
        L = fi.length; nsoft = fi.soft_bytes;
    } else if (active) active = job.code_rate == (uint32_t)CODE_RATE;
    if (!active) return;                                // whole quad leaves together (no block-wide sync below)
    constexpr uint32_t GROUP = CODE_RATE == CR_12 ? 2u : CODE_RATE == CR_34 ? 4u : 3u;   // soft bytes per puncture group
    constexpr uint32_t GSTEPS = CODE_RATE == CR_12 ? 1u : CODE_RATE == CR_34 ? 3u : 2u;  // trellis steps per group
    const uint32_t depth = job.depth, look = job.lookahead;
    const uint8_t* sp = soft + (size_t)f * soft_stride;
    uint8_t* op = out + (size_t)f * out_stride;
    const uint32_t out_cap = (uint32_t)(out_stride < 0xFFFFFFFFull ? out_stride : 0xFFFFFFFFull);
    // lane constants
    VqLane LC;
#pragma unroll
    for (int t = 0; t < 6; t++) {
        int lc = q == 0 ? vq_lcls(t, 0) : q == 1 ? vq_lcls(t, 1) : q == 2 ? vq_lcls(t, 2) : vq_lcls(t, 3);
        LC.swz[t] = (unsigned)((0 ^ lc) | ((1 ^ lc) << 4) | ((2 ^ lc) << 8) | ((3 ^ lc) << 12));
    }
    {
        const uint32_t EV = 0x00FE00FEu, FF = 0x00FF00FFu, ONE = 0x00010001u;
        int b0 = (q >> 1) & 1, b1 = q & 1;              // pair bit at T=0 is address bit 5 (lane bit 1), at T=1 address bit 4
        LC.m1[0] = b0 ? FF : EV; LC.o1[0] = b0 ? ONE : 0u; LC.m2[0] = b0 ? EV : FF; LC.o2[0] = b0 ? 0u : ONE;
        LC.m1[1] = b1 ? FF : EV; LC.o1[1] = b1 ? ONE : 0u; LC.m2[1] = b1 ? EV : FF; LC.o2[1] = b1 ? 0u : ONE;
    }
    // initial metrics (viterbilut.h:22-32): state 0 -> 0x00, others 0x30; at t=0 state == address
    uint32_t R[8];
#pragma unroll
    for (int r = 0; r < 8; r++) R[r] = 0x00300030u;
    if (q == 0) R[0] = 0x00300000u;
    const uint32_t end = L * 8u + 16u + 6u;
    uint32_t t = 0, ob = 0, tmod = 0;                   // tmod = t % 6
    uint32_t wcol = 0;                                  // ring slot of column t (t % SB_VQ_RING)
    uint32_t desc_count = 0, desc_reg = 0, byte_count = 0, crc = 0xFFFFFFFFu, fcs = 0, verdict = E_SUCCESS, nraw = 0;
    bool done = false;
    uint16_t* ring16 = (uint16_t*)&s_ring[0][0];
    uint32_t pos_soft = 0;

    auto after_group = [&]() {
        if ((t & 7u) == 0) {                            // viterbi.hpp:177-180 -> viterbicore.h:445-465
            uint32_t m = __vminu2(__vminu2(__vminu2(R[0], R[1]), __vminu2(R[2], R[3])), __vminu2(__vminu2(R[4], R[5]), __vminu2(R[6], R[7])));
            m = min(m & 0xFFFFu, m >> 16);
            m = min(m, __shfl_xor_sync(QM, m, 1)); m = min(m, __shfl_xor_sync(QM, m, 2));
            m &= 0xFEu;
            const uint32_t mv = m * 0x00010001u;
#pragma unroll
            for (int r = 0; r < 8; r++) R[r] -= mv;        // every half >= m: no borrow between halves
        }
        uint32_t nout = 0, la = 0;                      // viterbi.hpp:182-203
        if (!done) {
            if (t >= end) { nout = end - ob - 6u; la = t - end; }
            else if (t >= ob + depth + look + 6u) { nout = depth; la = look + (t - (ob + depth + look + 6u)) % 8u; }
        }
        if (nout) {                                    // nout is uniform inside the quad
            // best state: smallest (metric, state) key over the 64 states (viterbicore.h:468-520)
            uint32_t m = __vminu2(__vminu2(__vminu2(R[0], R[1]), __vminu2(R[2], R[3])), __vminu2(__vminu2(R[4], R[5]), __vminu2(R[6], R[7])));
            m = min(m & 0xFFFFu, m >> 16);
            m = min(m, __shfl_xor_sync(QM, m, 1)); m = min(m, __shfl_xor_sync(QM, m, 2));
            uint32_t best = 64;
#pragma unroll
            for (int r = 0; r < 8; r++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t v = h ? (R[r] >> 16) : (R[r] & 0xFFFFu);
                    uint32_t A = ((uint32_t)q << 4) | (r << 1) | h;
                    uint32_t n = ((A << tmod) | (A >> (6u - tmod))) & 63u;     // state index of this slot at time t
                    if (v == m && n < best) best = n;
                }
            }
            best = min(best, __shfl_xor_sync(QM, best, 1)); best = min(best, __shfl_xor_sync(QM, best, 2));
            __syncwarp(QM);
            if (q == 0) {
                int pos = (int)(best | ((m & 1u) << 6));
                uint32_t col = wcol, cm = tmod;         // ring slot / phase of the column being read
                auto back = [&]() {
                    col = col ? col - 1 : SB_VQ_RING - 1; cm = cm ? cm - 1 : 5;
                    pos = (pos >> 1) & 0x3F;
                    unsigned long long w = s_ring[col][fb];
                    uint32_t A = (((uint32_t)pos | ((uint32_t)pos << 6)) >> cm) & 63u;   // ror6(pos, cm)
                    pos |= (int)((w >> A) & 1ull) << 6;
                };
                for (uint32_t i = 0; i < la; i++) back();
                const uint32_t nbytes = nout >> 3;
                uint8_t* win = s_win[fb];
                for (uint32_t b = 0; b < nbytes; b++) {
                    uint32_t ch = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) { ch = (ch << 1) | (uint32_t)((pos >> 6) & 1); back(); }
                    win[nbytes - 1 - b] = (uint8_t)ch;
                }
                if (job.raw) {
                    for (uint32_t b = 0; b < nbytes; b++) op[(size_t)nraw + b] = win[b];
                } else {
                    for (uint32_t b = 0; b < nbytes; b++) {       // scramble.hpp:323-351, PHY_11a.hpp:655-700
                        uint32_t by = win[b];
                        desc_count++;
                        if (desc_count == 1) continue;
                        if (desc_count == 2) { desc_reg = by >> 1; continue; }
                        desc_reg = s_scr[desc_reg];
                        uint32_t o = by ^ desc_reg; desc_reg >>= 1;
                        if (byte_count < (uint32_t)((int)L - 4)) {
                            if (byte_count < out_cap) op[byte_count] = (uint8_t)o;
                            byte_count++;
                            crc = (crc >> 8) ^ s_crc[(o ^ crc) & 0xFF];
                        } else if (byte_count < L) {
                            if (byte_count < out_cap) op[byte_count] = (uint8_t)o;
                            byte_count++;
                            fcs |= o << (8u * (byte_count - 1u - (L - 4u)));
                            if (byte_count == L) verdict = (~crc == fcs) ? (uint32_t)E_FRAME_OK : (uint32_t)E_CRC32_FAIL;
                        }
                    }
                }
                nraw += nbytes;
            }
            ob += nout;
            if (ob + 6u >= end && t >= end) done = true;
            __syncwarp(QM);
        }
    };
    auto commit = [&]() {                               // after a step: store survivor bits of the new column
        t++; tmod = tmod == 5 ? 0 : tmod + 1; wcol = wcol == SB_VQ_RING - 1 ? 0 : wcol + 1;
        ring16[(wcol * SB_VQ_FR + fb) * 4 + q] = (uint16_t)vq_decisions(R);
    };

    // main loop: 6 trellis steps (one phase cycle) per iteration; the next chunk's soft values are prefetched
    constexpr uint32_t CHUNK_BYTES = 6u / GSTEPS * GROUP;        // 12 (R=1/2), 9 (2/3), 8 (3/4)
    uint32_t w0 = 0, w1 = 0, w2 = 0;                    // current chunk, little-endian bytes
    auto fetch = [&](uint32_t pos, uint32_t& a0, uint32_t& a1, uint32_t& a2) {
        if (pos + CHUNK_BYTES > nsoft) { a0 = a1 = a2 = 0; return; }
        if (CODE_RATE == CR_34) { uint2 v = __ldg((const uint2*)(sp + pos)); a0 = v.x; a1 = v.y; a2 = 0; }
        else if (CODE_RATE == CR_12) { a0 = __ldg((const uint32_t*)(sp + pos)); a1 = __ldg((const uint32_t*)(sp + pos + 4)); a2 = __ldg((const uint32_t*)(sp + pos + 8)); }
        else { uint32_t b[9];
#pragma unroll
               for (int i = 0; i < 9; i++) b[i] = __ldg(sp + pos + i);
               a0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); a1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24); a2 = b[8]; }
    };
    fetch(0, w0, w1, w2);
#define SV(i) ((int)(((i) < 4 ? w0 >> (8 * (i)) : (i) < 8 ? w1 >> (8 * ((i) - 4)) : w2 >> (8 * ((i) - 8))) & 0xFFu))
    while (!done && pos_soft + CHUNK_BYTES <= nsoft) {
        uint32_t n0, n1, n2; fetch(pos_soft + CHUNK_BYTES, n0, n1, n2);
        pos_soft += CHUNK_BYTES;
        if (CODE_RATE == CR_12) {
            vq_step<0>(R, vq_bm_ab(SV(0), SV(1)), LC, QM);   commit(); after_group();
            vq_step<1>(R, vq_bm_ab(SV(2), SV(3)), LC, QM);   commit(); after_group();
            vq_step<2>(R, vq_bm_ab(SV(4), SV(5)), LC, QM);   commit(); after_group();
            vq_step<3>(R, vq_bm_ab(SV(6), SV(7)), LC, QM);   commit(); after_group();
            vq_step<4>(R, vq_bm_ab(SV(8), SV(9)), LC, QM);   commit(); after_group();
            vq_step<5>(R, vq_bm_ab(SV(10), SV(11)), LC, QM); commit(); after_group();
        } else if (CODE_RATE == CR_34) {
            vq_step<0>(R, vq_bm_ab(SV(0), SV(1)), LC, QM); commit();
            vq_step<1>(R, vq_bm_a(SV(2)), LC, QM);         commit();
            vq_step<2>(R, vq_bm_b(SV(3)), LC, QM);         commit(); after_group();
            vq_step<3>(R, vq_bm_ab(SV(4), SV(5)), LC, QM); commit();
            vq_step<4>(R, vq_bm_a(SV(6)), LC, QM);         commit();
            vq_step<5>(R, vq_bm_b(SV(7)), LC, QM);         commit(); after_group();
        } else {
            vq_step<0>(R, vq_bm_ab(SV(0), SV(1)), LC, QM); commit();
            vq_step<1>(R, vq_bm_a(SV(2)), LC, QM);         commit(); after_group();
            vq_step<2>(R, vq_bm_ab(SV(3), SV(4)), LC, QM); commit();
            vq_step<3>(R, vq_bm_a(SV(5)), LC, QM);         commit(); after_group();
            vq_step<4>(R, vq_bm_ab(SV(6), SV(7)), LC, QM); commit();
            vq_step<5>(R, vq_bm_a(SV(8)), LC, QM);         commit(); after_group();
        }
        w0 = n0; w1 = n1; w2 = n2;
    }
#undef SV
    // tail: remaining whole puncture groups that do not fill a 6-step chunk (standalone API with arbitrary nsoft)
    while (!done && pos_soft + GROUP <= nsoft) {
        int s0 = __ldg(sp + pos_soft), s1 = __ldg(sp + pos_soft + 1);
        int s2 = GROUP > 2 ? (int)__ldg(sp + pos_soft + 2) : 0, s3 = GROUP > 3 ? (int)__ldg(sp + pos_soft + 3) : 0;
        pos_soft += GROUP;
        vq_step_rt((int)tmod, R, vq_bm_ab(s0, s1), LC, QM); commit();
        if (GSTEPS >= 2) { vq_step_rt((int)tmod, R, vq_bm_a(s2), LC, QM); commit(); }
        if (GSTEPS >= 3) { vq_step_rt((int)tmod, R, vq_bm_b(s3), LC, QM); commit(); }
        after_group();
    }
    if (q == 0) {
        if (!job.raw) { if (verdict == E_SUCCESS) verdict = E_FAILED; status_out[f] = verdict; crc_out[f] = fcs; }
        else { status_out[f] = nraw; crc_out[f] = 0; }
    }
}

} // namespace sb
