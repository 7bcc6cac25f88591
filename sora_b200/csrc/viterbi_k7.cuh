// sora_b200 — batched K=7 (133,171) soft-decision Viterbi for sm_100a, bit-exact with the reference's
// uint8-metric SSE decoder (kernel/bb/Brick11/src/viterbicore.h:269-556 TViterbiCore, driven with the cadence of
// kernel/bb/Brick11/src/viterbi.hpp:104-237 T11aViterbi; descramble = scramble.hpp:269-355 T11aDesc; CRC/verdict =
// PHY_11a.hpp:609-702 TBB11aFrameSink).
//
// v1 mapping ("warp per code block"): lane l owns the trellis butterfly (l, l+32) -> (2l, 2l+1).
//   * ACS is purely lane-local (the two predecessors of both outputs are this lane's two metrics);
//   * the 64 survivor bits of a step are two warp ballots, kept in a 512-column shared-memory ring
//     (the reference keeps the whole 64-byte column; only its LSBs are ever read back);
//   * the new metrics are repacked (2l,2l+1) -> (l, l+32) with two warp shuffles: the path-metric exchange;
//   * uint8 wrap, survivor mark in the LSB, unsigned min, "(steps & 7)==0 after a puncture group" normalisation
//     with the LSB masked, 256-bit windowed traceback with 24..31 look-ahead and the final flush are all mirrored.
// Soft input is staged 16 B per lane per refill (coalesced 512 B per warp) and read back by warp shuffle.
#pragma once
#include "rx11a_kernels.cuh"

namespace sb {

#define SB_VIT_WARPS 8
#define SB_VIT_RING 512

struct VitJob {               // uniform-parameter mode (standalone API); per-frame mode reads FrameInfo instead
    uint32_t code_rate, frame_len, nsoft; uint32_t depth, lookahead; uint32_t raw; // raw=1: emit SERVICE+PSDU bytes undescrambled
};

__global__ void __launch_bounds__(32 * SB_VIT_WARPS) k_viterbi_k7(const uint8_t* __restrict__ soft, uint64_t soft_stride,
        uint32_t nframes, const FrameInfo* __restrict__ info, VitJob job, DevTables T,
        uint8_t* __restrict__ out, uint64_t out_stride, uint32_t* __restrict__ status_out, uint32_t* __restrict__ crc_out) {
    __shared__ uint2 s_ring[SB_VIT_WARPS][SB_VIT_RING];
    __shared__ uint32_t s_crc[256];
    __shared__ uint8_t s_scr[128];
    __shared__ uint8_t s_win[SB_VIT_WARPS][48];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = __ldg(T.crc32 + i);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_scr[i] = __ldg(T.scramble + i);
    __syncthreads();
    const uint32_t f = blockIdx.x * SB_VIT_WARPS + wib;
    if (f >= nframes) return;
    uint32_t code_rate = job.code_rate, L = job.frame_len, nsoft = job.nsoft;
    if (info) {
        FrameInfo fi = info[f];
        if (fi.status != E_SUCCESS) { if (lane == 0) { status_out[f] = fi.status; crc_out[f] = 0; } return; }
        code_rate = fi.code_rate; L = fi.length; nsoft = fi.soft_bytes;
    }
    const uint32_t depth = job.depth, look = job.lookahead;
    const uint8_t* sp = soft + (size_t)f * soft_stride;
    uint8_t* op = out + (size_t)f * out_stride;
    const uint32_t out_cap = (uint32_t)(out_stride < 0xFFFFFFFFull ? out_stride : 0xFFFFFFFFull);
    uint2* ring = s_ring[wib];
    const int cA = ((lane >> 1) ^ (lane >> 2) ^ (lane >> 4)) & 1;
    const int cB = (lane ^ (lane >> 1) ^ (lane >> 2)) & 1;
    const int src_a = lane >> 1, src_b = 16 + (lane >> 1), shb = 8 * (lane & 1);
    int m0 = lane == 0 ? 0x00 : 0x30, m1 = 0x30;
    const uint32_t end = L * 8u + 16u + 6u;
    uint32_t t = 0, ob = 0;
    // lane-0 back-end state (descrambler + frame sink)
    uint32_t desc_count = 0, desc_reg = 0, byte_count = 0, crc = 0xFFFFFFFFu, fcs = 0, verdict = E_SUCCESS, nraw = 0;
    const uint32_t group = code_rate == CR_12 ? 2u : code_rate == CR_34 ? 4u : 3u;
    uint32_t pos_soft = 0;
    uint4 stage = make_uint4(0, 0, 0, 0); uint32_t stage_base = 0xFFFFFFFFu;     // 512 soft bytes per warp refill
    auto soft_at = [&](uint32_t i) -> int {            // i is warp-uniform
        uint32_t r = i - stage_base;                   // 0..511
        uint32_t w = (r & 15u) >> 2;
        uint32_t word = w == 0 ? stage.x : w == 1 ? stage.y : w == 2 ? stage.z : stage.w;
        word = __shfl_sync(FULL, word, r >> 4);
        return (int)((word >> (8u * (r & 3u))) & 0xFFu);
    };
    auto acs = [&](int alpha, int beta) {
        int n0 = min((m0 + alpha) & 0xFE, ((m1 + beta) & 0xFF) | 1);
        int n1 = min((m0 + beta) & 0xFE, ((m1 + alpha) & 0xFF) | 1);
        uint32_t e = __ballot_sync(FULL, n0 & 1), o = __ballot_sync(FULL, n1 & 1);
        t++;
        if (lane == 0) ring[t & (SB_VIT_RING - 1)] = make_uint2(e, o);
        int w = n0 | (n1 << 8);
        int wa = __shfl_sync(FULL, w, src_a), wb = __shfl_sync(FULL, w, src_b);
        m0 = (wa >> shb) & 0xFF; m1 = (wb >> shb) & 0xFF;
    };
    while (pos_soft + group <= nsoft) {
        if (pos_soft + group > stage_base + 512u || stage_base == 0xFFFFFFFFu) {     // refill on a 16-byte boundary
            stage_base = pos_soft & ~15u;
            uint32_t o = stage_base + 16u * lane;
            if (o + 16u <= ((nsoft + 15u) & ~15u)) stage = __ldg((const uint4*)(sp + o)); else stage = make_uint4(0, 0, 0, 0);
        }
        {   // one puncture group (viterbi.hpp:151-173)
            int tA = 2 * soft_at(pos_soft), tB = 2 * soft_at(pos_soft + 1);
            int a = (cA ? 14 - tA : tA) + (cB ? 14 - tB : tB); acs(a, 28 - a);
            if (code_rate != CR_12) { int tC = 2 * soft_at(pos_soft + 2); int a2 = cA ? 14 - tC : tC; acs(a2, 14 - a2); }
            if (code_rate == CR_34) { int tD = 2 * soft_at(pos_soft + 3); int a3 = cB ? 14 - tD : tD; acs(a3, 14 - a3); }
            pos_soft += group;
        }
        if ((t & 7u) == 0) {                           // viterbi.hpp:177-180 -> viterbicore.h:445-465
            int mn = __reduce_min_sync(FULL, min(m0, m1)) & 0xFE;
            m0 = (m0 - mn) & 0xFF; m1 = (m1 - mn) & 0xFF;
        }
        uint32_t nout = 0, la = 0;                     // viterbi.hpp:182-203
        if (t >= end) { nout = end - ob - 6u; la = t - end; }
        else if (t >= ob + depth + look + 6u) { nout = depth; la = look + (t - (ob + depth + look + 6u)) % 8u; }
        if (nout) {
            unsigned key = min(((unsigned)m0 << 8) | ((unsigned)lane << 2), ((unsigned)m1 << 8) | ((unsigned)(lane + 32) << 2));
            key = __reduce_min_sync(FULL, key);        // viterbicore.h:468-520
            __syncwarp();
            if (lane == 0) {
                int pos = (int)(key >> 2) & 0x7F;
                uint32_t col = t;
                for (uint32_t i = 0; i < la; i++) {
                    col--; pos = (pos >> 1) & 0x3F;
                    uint2 d = ring[col & (SB_VIT_RING - 1)];
                    pos |= (int)((((pos & 1) ? d.y : d.x) >> (pos >> 1)) & 1u) << 6;
                }
                const uint32_t nbytes = nout >> 3;     // <= 35 (final flush)
                uint8_t* win = s_win[wib];             // bytes come out last-first; the sink needs them first-first
                for (uint32_t b = 0; b < nbytes; b++) {
                    uint32_t ch = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        ch = (ch << 1) | (uint32_t)((pos >> 6) & 1);
                        col--; pos = (pos >> 1) & 0x3F;
                        uint2 d = ring[col & (SB_VIT_RING - 1)];
                        pos |= (int)((((pos & 1) ? d.y : d.x) >> (pos >> 1)) & 1u) << 6;
                    }
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
            __syncwarp();
        }
        if (ob + 6u >= end && t >= end) break;         // everything flushed; remaining pad bits carry no output
    }
    if (lane == 0) {
        if (!job.raw) { if (verdict == E_SUCCESS) verdict = E_FAILED; status_out[f] = verdict; crc_out[f] = fcs; }
        else { status_out[f] = nraw; crc_out[f] = 0; }
    }
}

} // namespace sb
