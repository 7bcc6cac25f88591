// sora_b200 — batched K=7 (133,171) soft Viterbi, v3 "history-carrying" kernel for sm_100a.
//
// Arithmetic contract (bit-exact with kernel/bb/Brick11/src/viterbicore.h:269-556 driven like
// kernel/bb/Brick11/src/viterbi.hpp:104-237): path metrics are the reference's uint8 values, whose LSB is the survivor mark and whose upper
// seven bits m7 are the metric proper (every branch metric is even, viterbilut.h, so the mark never carries into m7):
//     new = min_u8( (old[p] + bm) & 0xFE , (old[p+32] + bm') | 1 )   ==   m7' = min(a7, b7) mod 128 with ties to the even candidate,
//     decision d = (b7 < a7).
//
// Machine mapping (same quad / in-place trellis as v2, viterbi_k7_quad.cuh, different bookkeeping):
//   * 4 lanes per code block, 16 metrics per lane, two per register as 16-bit halves; m7 sits in bits 9..15 of its half, so the uint8 wrap is the
//     half's own carry-out.  Every add is a per-half SIMD add (VIADD.16x2, or the add inside VIADDMNMX.U16x2): nothing crosses halves.
//   * The low bits of a half carry the survivor HISTORY of the path that ends in that state: trellis step T of a 6-step block (T = t mod 6)
//     gives the odd candidate bit T (folded into its branch-metric constant) and leaves it clear in the even candidate.  Bits T+1..8 are still
//     zero in both candidates, so a metric tie is decided by bit T exactly like the reference's LSB mark decides it (even wins), bits below T
//     are never reached by the compare, and the min moves the winner's history along for free.  One fused VIADDMNMX.U16x2 (DPX add-min) plus
//     one VIADD.16x2 per register and step is the whole add-compare-select: no role masks, no per-step gathering of decision bits.
//   * After the 6th step every half holds the six decisions of its survivor over the block (register exchange over one phase cycle).  At a
//     block boundary state index == slot address, so four PRMTs line the 16 history bytes of a lane up in address order and one 128-bit
//     store puts them into a shared-memory ring: entry [block][code block][slot] = 64 bytes per 6 columns.
//   * Traceback jumps a whole block per lookup: the history byte h of slot A gives the six decoded bits (bit T = column 6b+T+1) and the
//     predecessor slot six columns earlier is bit-reverse6(h) — K=7 replaces all six state bits in six steps.  A window of 256+24..31 columns
//     costs ~48 byte loads instead of 280 64-bit word look-ups.
//   * The kernel emits the decoded bytes (SERVICE + PSDU, still scrambled).  Descrambler, CRC-32 and the verdict (scramble.hpp:269-355,
//     PHY_11a.hpp:609-702) run in k_sink11a, one thread per frame, after it.
// Measured instruction rates that shaped this (tools/microbench/pipes.cu on B200): VIADDMNMX.U16x2 + IMAD/VIADD issue together at ~1.0 per
// cycle per sub-partition; VIMNMX + two adds (the v2 ACS) at 0.75; LOP3 and PRMT at 0.5.
#pragma once
#include "viterbi_k7_quad.cuh"

namespace sb {

#define SB_VR_WARPS 1                      // warps per CTA
#define SB_VR_FR (8 * SB_VR_WARPS)         // code blocks per CTA
#define SB_VR_NB 50                        // ring entries of 6 columns: depth + lookahead + 7 <= 288 columns = 48 entries, + the partial one + 1

// low four address bits of (register r, half h): r << 1 | h — the order in which the PRMT gather of commit lays a lane's 16 history bytes down
__host__ __device__ constexpr int vr_low4(int r, int h) { return (r << 1) | h; }
__host__ __device__ constexpr int vr_scls(int T, int r, int h) { return vq_cls(vq_rol6(vr_low4(r, h), T) & 31); }
__host__ __device__ constexpr int vr_kcls(int T) { return vq_cls(vq_rol6(1, T) & 31); }     // class difference between the two halves of a register

struct VrLane {
    unsigned swz[6];       // per-phase byte swizzle applying this lane's class contribution
    unsigned mA[2], mB[2]; // lane-pair phases: history mark of this step for my own value / for my partner's (exactly one is non-zero)
};

// one trellis step at compile-time phase T.  Cbase byte (cA<<1|cB) = metric of the even candidate for a predecessor of that class; the
// complement class (3 - index) is the odd candidate's.  The odd candidate (state p+32) carries bit T of the low byte.
template <int T>
__device__ __forceinline__ void vr_step(uint32_t (&R)[8], uint32_t Cbase, const VrLane& L, unsigned qmask) {
    const uint32_t Cb = __byte_perm(Cbase, 0, L.swz[T]);
    constexpr uint32_t ONE = 0x00010001u << T;
    constexpr int K = vr_kcls(T);
    uint32_t V[4];                                      // [0, Cb[c], 0, Cb[c ^ K]]: branch metrics of class c (low half) and its high-half companion
#pragma unroll
    for (int c = 0; c < 4; c++) V[c] = __byte_perm(Cb, 0, vq_sel(c, c ^ K));
    if (T <= 1) {                                       // pair = partner lane (xor 2 at T=0, xor 1 at T=1): the path-metric exchange
        uint32_t Va[4], Vb[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { Va[c] = V[c] + L.mA[T]; Vb[c] = V[c] + L.mB[T]; }
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int c0 = vr_scls(T, r, 0);
            const uint32_t Z = __shfl_xor_sync(qmask, R[r], T == 0 ? 2 : 1);
            R[r] = __viaddmin_u16x2(R[r], Va[c0], __vadd2(Z, Vb[c0 ^ 3]));
        }
    } else if (T <= 4) {                                // pair = register r ^ d inside the lane
        constexpr int d = T == 2 ? 4 : T == 3 ? 2 : 1;
        uint32_t VO[4];
#pragma unroll
        for (int c = 0; c < 4; c++) VO[c] = V[c] + ONE;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (r & d) continue;
            const int c0 = vr_scls(T, r, 0);
            const uint32_t X = R[r], Y = R[r + d];      // X = states p (even role), Y = states p + 32 (odd role)
            R[r]     = __viaddmin_u16x2(X, V[c0],  __vadd2(Y, VO[c0 ^ 3]));
            R[r + d] = __viaddmin_u16x2(Y, VO[c0], __vadd2(X, V[c0 ^ 3]));
        }
    } else {                                            // pair = the two halves of each register: low = p, high = p + 32
        uint32_t W1[4], W2[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { W1[c] = V[c] + (ONE & 0xFFFF0000u); W2[c] = V[c] + (ONE & 0x0000FFFFu); }
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int c = vr_scls(5, r, 0);
            const uint32_t y = __byte_perm(R[r], 0, 0x1032);          // [p + 32, p]
            R[r] = __viaddmin_u16x2(R[r], W1[c], __vadd2(y, W2[c ^ 3]));   // [min(p + a, p32 + b), min(p32 + a, p + b)] = new states 2p, 2p + 1
        }
    }
}

template <int CODE_RATE>
__global__ void __launch_bounds__(32 * SB_VR_WARPS) k_viterbi_re(const uint8_t* __restrict__ soft, uint64_t soft_stride,
        uint32_t nframes, const FrameInfo* __restrict__ info, VitJob job,
        uint8_t* __restrict__ out, uint64_t out_stride, uint32_t raw_off, uint32_t* __restrict__ nraw_out) {
    __shared__ uint4 s_ring[SB_VR_NB][SB_VR_FR][4];    // entry b: history bytes of the 64 slots of every code block after column 6(b+1)
    const int lane = threadIdx.x & 31, q = lane & 3;
    const unsigned QM = 0xFu << (lane & 28);           // the 4 lanes of this code block: quads run as independent sub-warps
    const int fb = (threadIdx.x >> 2);                 // code block within the CTA
    const uint32_t f = blockIdx.x * SB_VR_FR + fb;
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
    uint8_t* op = out + (size_t)f * out_stride + raw_off;
    const uint32_t out_cap = (uint32_t)(out_stride - raw_off < 0xFFFFFFFFull ? out_stride - raw_off : 0xFFFFFFFFull);
    VrLane LC;
#pragma unroll
    for (int t = 0; t < 6; t++) {
        int lc = q == 0 ? vq_lcls(t, 0) : q == 1 ? vq_lcls(t, 1) : q == 2 ? vq_lcls(t, 2) : vq_lcls(t, 3);
        LC.swz[t] = (unsigned)((0 ^ lc) | ((1 ^ lc) << 4) | ((2 ^ lc) << 8) | ((3 ^ lc) << 12));
    }
    {
        const int b0 = (q >> 1) & 1, b1 = q & 1;        // pair bit at T=0 is address bit 5 (lane bit 1), at T=1 address bit 4
        LC.mA[0] = b0 ? 0x00010001u : 0u; LC.mB[0] = b0 ? 0u : 0x00010001u;
        LC.mA[1] = b1 ? 0x00020002u : 0u; LC.mB[1] = b1 ? 0u : 0x00020002u;
    }
    // initial metrics (viterbilut.h:22-32): state 0 -> 0x00, others 0x30; at t=0 state == address; byte value v sits at v << 8
    uint32_t R[8];
#pragma unroll
    for (int r = 0; r < 8; r++) R[r] = 0x30003000u;
    if (q == 0) R[0] = 0x30000000u;
    const uint32_t end = L * 8u + 16u + 6u;
    uint32_t tb = 0, ob = 0;                            // tb = trellis time at the start of the current 6-step block
    uint32_t wslot = 0;                                 // ring entry of the block that starts at tb
    uint32_t next_tb = min(end, depth + look + 6u);     // first time a traceback can fire (viterbi.hpp:182-203)
    uint32_t nraw = 0;
    bool done = false;
    uint4* ring_q = &s_ring[0][fb][q];                  // + entry * (SB_VR_FR * 4)
    const uint8_t* ring_b = (const uint8_t*)&s_ring[0][fb][0];   // + entry * (SB_VR_FR * 64) + slot
    uint32_t pos_soft = 0;

    // the 16 history bytes of this lane (low byte of every half), in address order, into ring entry `e`
    auto store_hist = [&](const uint32_t e) {
        uint4 w;
        w.x = __byte_perm(R[0], R[1], 0x6420); w.y = __byte_perm(R[2], R[3], 0x6420);
        w.z = __byte_perm(R[4], R[5], 0x6420); w.w = __byte_perm(R[6], R[7], 0x6420);
        ring_q[e * (SB_VR_FR * 4)] = w;
    };
    // windowed traceback from slot A0 at time t = tb + k (k = columns into the current block; its histories are in entry wslot), viterbi.hpp:205-237
    auto do_traceback = [&](const uint32_t A0, const uint32_t k, const uint32_t la, const uint32_t nout) {
        __syncwarp(QM);
        if (q == 0) {
            uint32_t A = A0, todo = la + nout, e = wslot;
            uint32_t fifo = 0; int cnt = -(int)la;                  // the first `la` bits are only looked through; at most 13 bits wait
            uint32_t wpos = nraw + (nout >> 3);                      // bytes come out last-first
            auto push = [&](uint32_t bits, uint32_t n) {
                fifo = (fifo << n) | bits; cnt += (int)n;
                while (cnt >= 8) { --wpos; if (wpos < out_cap) op[wpos] = (uint8_t)(fifo >> (cnt - 8)); cnt -= 8; }
            };
            if (k < 6u) {                                            // partial newest block: k decisions in bits 0..k-1, the top k address bits change
                const uint32_t h = ring_b[e * (SB_VR_FR * 64) + A];
                const uint32_t take = min(k, todo), am = (0x3Fu << (6u - k)) & 0x3Fu;
                A = (A & ~am) | ((__brev(h) >> 26) & am);
                push((h & ((1u << k) - 1u)) >> (k - take), take);
                todo -= take; e = e ? e - 1u : SB_VR_NB - 1u;
            }
            while (todo >= 6u) {
                const uint32_t h = ring_b[e * (SB_VR_FR * 64) + A] & 0x3Fu;
                A = __brev(h) >> 26;                                 // predecessor slot six columns earlier
                push(h, 6u);
                todo -= 6u; e = e ? e - 1u : SB_VR_NB - 1u;
            }
            if (todo) {                                              // oldest block of the window: only its newest `todo` columns
                const uint32_t h = ring_b[e * (SB_VR_FR * 64) + A] & 0x3Fu;
                push(h >> (6u - todo), todo);
            }
        }
        nraw += nout >> 3;
        __syncwarp(QM);
    };
    // normalisation + traceback triggers, evaluated after every puncture group at time t = tb + k, phase tm = t % 6 (compile-time in the main loop)
    auto after_group = [&](const uint32_t t, const uint32_t tm, const uint32_t k) {
        if ((t & 7u) == 0) {                            // viterbi.hpp:177-180 -> viterbicore.h:445-465: subtract (smallest byte & 0xFE) = smallest m7
            uint32_t m = __vminu2(__vminu2(__vminu2(R[0], R[1]), __vminu2(R[2], R[3])), __vminu2(__vminu2(R[4], R[5]), __vminu2(R[6], R[7])));
            m = min(m & 0xFFFFu, m >> 16) >> 9;         // smallest m7 of this lane
            m = min(m, __shfl_xor_sync(QM, m, 1)); m = min(m, __shfl_xor_sync(QM, m, 2));
            const uint32_t mv = m * 0x02000200u;
#pragma unroll
            for (int r = 0; r < 8; r++) R[r] -= mv;     // every half >= m << 9: no borrow between halves, histories untouched
        }
        if (t < next_tb) return;
        uint32_t nout, la;                              // viterbi.hpp:182-203
        if (t >= end) { nout = end - ob - 6u; la = t - end; }
        else { nout = depth; la = look + (t - (ob + depth + look + 6u)) % 8u; }
        if (nout) {                                     // uniform inside the quad
            // best state: smallest (byte = m7 << 1 | newest mark, state index) over the 64 slots (viterbicore.h:468-520)
            const uint32_t tn = tm ? tm - 1u : 5u;      // bit of the newest decision
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int r = 0; r < 8; r++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t v = h ? (R[r] >> 16) : (R[r] & 0xFFFFu);
                    const uint32_t A = ((uint32_t)q << 4) | (uint32_t)vr_low4(r, h);
                    const uint32_t ns = ((A << tm) | (A >> (6u - tm))) & 63u;         // state index of this slot at time t
                    best = min(best, ((((v >> 9) << 1) | ((v >> tn) & 1u)) << 8) | ns);
                }
            }
            best = min(best, __shfl_xor_sync(QM, best, 1)); best = min(best, __shfl_xor_sync(QM, best, 2));
            const uint32_t n = best & 63u;
            const uint32_t A0 = ((n >> tm) | (n << (6u - tm))) & 63u;
            if (k < 6u) store_hist(wslot);              // mid-block: the partial histories of the running block (a block end has just stored its own)
            do_traceback(A0, k, la, nout);
            ob += nout;
        }
        if (ob + 6u >= end && t >= end) done = true;
        next_tb = min(end, ob + depth + look + 6u);
        if (next_tb <= t) next_tb = t + 1u;              // a frame shorter than the prefix: re-evaluate at every group
    };
    auto clear_hist = [&]() {
#pragma unroll
        for (int r = 0; r < 8; r++) R[r] &= 0xFE00FE00u;
    };

    // main loop: 6 trellis steps (one phase cycle = one history block) per iteration; the next chunk's soft values are prefetched
    constexpr uint32_t CHUNK_BYTES = 6u / GSTEPS * GROUP;        // 12 (R=1/2), 9 (2/3), 8 (3/4)
    uint32_t w0 = 0, w1 = 0, w2 = 0;                    // current chunk, little-endian bytes
    auto fetch = [&](uint32_t pos, uint32_t& a0, uint32_t& a1, uint32_t& a2) {
        if (pos + CHUNK_BYTES > nsoft) { a0 = a1 = a2 = 0; return; }
        if (CODE_RATE == CR_34) { uint2 v = __ldg((const uint2*)(sp + pos)); a0 = v.x; a1 = v.y; a2 = 0; }
        else if (CODE_RATE == CR_12) { a0 = __ldg((const uint32_t*)(sp + pos)); a1 = __ldg((const uint32_t*)(sp + pos + 4)); a2 = __ldg((const uint32_t*)(sp + pos + 8)); }
        else { uint32_t b[9];
#pragma unroll
               for (int i = 0; i < 9; i++) b[i] = __ldg(sp + pos + i);
               // keep every puncture group inside one word: w0 = b0 b1 b2 -, w1 = b3 b4 b5 -, w2 = b6 b7 b8 -
               a0 = b[0] | (b[1] << 8) | (b[2] << 16); a1 = b[3] | (b[4] << 8) | (b[5] << 16); a2 = b[6] | (b[7] << 8) | (b[8] << 16); }
    };
    fetch(0, w0, w1, w2);
    while (!done && pos_soft + CHUNK_BYTES <= nsoft) {
        uint32_t n0, n1, n2; fetch(pos_soft + CHUNK_BYTES, n0, n1, n2);
        pos_soft += CHUNK_BYTES;
        if (CODE_RATE == CR_12) {
            vr_step<0>(R, vq_bm_ab<0>(w0), LC, QM); after_group(tb + 1, 1, 1);
            vr_step<1>(R, vq_bm_ab<2>(w0), LC, QM); after_group(tb + 2, 2, 2);
            vr_step<2>(R, vq_bm_ab<0>(w1), LC, QM); after_group(tb + 3, 3, 3);
            vr_step<3>(R, vq_bm_ab<2>(w1), LC, QM); after_group(tb + 4, 4, 4);
            vr_step<4>(R, vq_bm_ab<0>(w2), LC, QM); after_group(tb + 5, 5, 5);
            vr_step<5>(R, vq_bm_ab<2>(w2), LC, QM); store_hist(wslot); after_group(tb + 6, 0, 6);
        } else if (CODE_RATE == CR_34) {
            vr_step<0>(R, vq_bm_ab<0>(w0), LC, QM);
            vr_step<1>(R, vq_bm_a<2>(w0), LC, QM);
            vr_step<2>(R, vq_bm_b<3>(w0), LC, QM);  after_group(tb + 3, 3, 3);
            vr_step<3>(R, vq_bm_ab<0>(w1), LC, QM);
            vr_step<4>(R, vq_bm_a<2>(w1), LC, QM);
            vr_step<5>(R, vq_bm_b<3>(w1), LC, QM);  store_hist(wslot); after_group(tb + 6, 0, 6);
        } else {
            vr_step<0>(R, vq_bm_ab<0>(w0), LC, QM);
            vr_step<1>(R, vq_bm_a<2>(w0), LC, QM);  after_group(tb + 2, 2, 2);
            vr_step<2>(R, vq_bm_ab<0>(w1), LC, QM);
            vr_step<3>(R, vq_bm_a<2>(w1), LC, QM);  after_group(tb + 4, 4, 4);
            vr_step<4>(R, vq_bm_ab<0>(w2), LC, QM);
            vr_step<5>(R, vq_bm_a<2>(w2), LC, QM);  store_hist(wslot); after_group(tb + 6, 0, 6);
        }
        clear_hist();
        tb += 6; wslot = wslot == SB_VR_NB - 1u ? 0u : wslot + 1u;
        w0 = n0; w1 = n1; w2 = n2;
    }
    // tail: whole puncture groups that do not fill a 6-step block (standalone API with arbitrary nsoft).
    // Rare and short, so phases are dispatched at run time.
    {
        uint32_t k = 0;                                 // steps into the block at tb (the main loop always leaves it at 0)
        auto step_rt = [&](uint32_t Cbase) {
            switch (k) { case 0: vr_step<0>(R, Cbase, LC, QM); break; case 1: vr_step<1>(R, Cbase, LC, QM); break;
                         case 2: vr_step<2>(R, Cbase, LC, QM); break; case 3: vr_step<3>(R, Cbase, LC, QM); break;
                         case 4: vr_step<4>(R, Cbase, LC, QM); break; default: vr_step<5>(R, Cbase, LC, QM); }
            k++;
            if (k == 6) store_hist(wslot);
        };
        while (!done && pos_soft + GROUP <= nsoft) {
            uint32_t w = __ldg(sp + pos_soft) | ((uint32_t)__ldg(sp + pos_soft + 1) << 8);
            if (GROUP > 2) w |= (uint32_t)__ldg(sp + pos_soft + 2) << 16;
            if (GROUP > 3) w |= (uint32_t)__ldg(sp + pos_soft + 3) << 24;
            pos_soft += GROUP;
            step_rt(vq_bm_ab<0>(w));
            if (GSTEPS >= 2) step_rt(vq_bm_a<2>(w));
            if (GSTEPS >= 3) step_rt(vq_bm_b<3>(w));
            after_group(tb + k, k == 6 ? 0u : k, k);
            if (k == 6) { clear_hist(); k = 0; tb += 6; wslot = wslot == SB_VR_NB - 1u ? 0u : wslot + 1u; }
        }
    }
    if (q == 0) nraw_out[f] = nraw;
}

// Descrambler + frame sink after the Viterbi kernel, one thread per frame (T11aDesc, scramble.hpp:269-355; TBB11aFrameSink, PHY_11a.hpp:609-702):
// raw bytes (SERVICE + scrambled PSDU) sit at row + 14 so that the PSDU starts 16-byte aligned at row + 16; the descrambled PSDU goes to row + 0.
// The first SERVICE byte is dropped, the second seeds the register (byte >> 1), then out = byte ^ lut[reg], reg advances by eight bits per byte.
__global__ void __launch_bounds__(128) k_sink11a(uint8_t* __restrict__ out, uint64_t out_stride, uint32_t nframes, const FrameInfo* __restrict__ info,
                                                 DevTables T, uint32_t* __restrict__ status_io, uint32_t* __restrict__ crc_out) {
    __shared__ uint32_t s_crc[256];                    // CRC-32 (reflected 0xEDB88320, core/inc/CRC32.h:76)
    __shared__ uint8_t s_scr[128];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = __ldg(T.crc32 + i);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_scr[i] = __ldg(T.scramble + i);
    __syncthreads();
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const FrameInfo fi = info[f];
    if (fi.status != E_SUCCESS) return;                 // the front end already reached a terminal code: nothing was decoded
    const uint32_t L = fi.length, nraw = status_io[f];  // bytes the Viterbi kernel produced (SERVICE included)
    uint8_t* row = out + (size_t)f * out_stride;
    uint32_t verdict = E_FAILED, fcs = 0, crc = 0xFFFFFFFFu;
    if (nraw >= 2u) {
        uint32_t reg = row[15] >> 1;
        const uint32_t have = min(nraw - 2u, L);        // PSDU bytes available
        const uint32_t cap = (uint32_t)(out_stride < 0xFFFFFFFFull ? out_stride : 0xFFFFFFFFull);
        for (uint32_t i0 = 0; i0 < have; i0 += 16u) {
            const uint4 v = *(const uint4*)(row + 16 + i0);
            uint32_t w[4] = {v.x, v.y, v.z, v.w}, o4[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t ow = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t i = i0 + 4u * j + b;
                    reg = s_scr[reg];
                    const uint32_t o = ((w[j] >> (8 * b)) & 0xFFu) ^ reg; reg >>= 1;
                    ow |= o << (8 * b);
                    if (i < have) {
                        if (i + 4u < L) crc = (crc >> 8) ^ s_crc[(crc ^ o) & 0xFFu];
                        else if (L >= 4u) fcs |= o << (8u * (i + 4u - L));
                    }
                }
                o4[j] = ow;
            }
            if (i0 + 16u <= have && i0 + 16u <= cap) *(uint4*)(row + i0) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            else for (uint32_t i = i0; i < have && i < cap; i++) row[i] = (uint8_t)(o4[(i - i0) >> 2] >> (8u * ((i - i0) & 3u)));
        }
        if (have == L && L >= 4u) verdict = (~crc == fcs) ? (uint32_t)E_FRAME_OK : (uint32_t)E_CRC32_FAIL;
        else fcs = 0;                                   // frame_crc32 is only set once the last FCS byte arrived (PHY_11a.hpp:688)
    }
    status_io[f] = verdict; crc_out[f] = fcs;
}

} // namespace sb
