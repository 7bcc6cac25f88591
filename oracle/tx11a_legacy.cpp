// ORACLE — TEST INFRASTRUCTURE ONLY (see ops.h).
// CPU restatement of the reference's LEGACY 802.11a transmitter (the extern "C" path behind BB11ATxFrameMod / BB11AModulateACK,
// kernel/inc/bb/bba.h:240-262):
//   kernel/bb/dot11a/dot11/atx_fe.c:22-70 (rate dispatch), atx_tpl_imp.h:5-62 (Dot11aTxFrameEncodeX), atx_tpl.h:6-62 (Scramble11a),
//   atx.h:78-136 (GetSignal, CopyPreamble16_NT, GenerateSignal), inc/bb/mod/ofdmsymbol.h:13-245 (CopyGI, Window, GenerateNNMSymbol),
//   inc/bb/mod/{convenc,ainterleave,amap,addpilot,ifft64x,copynt,upsample}.h and the LUTs under lutst/.
// The LUT-driven stages (convolutional encoder with puncturing, interleaver, Gray mapper) are the standard clause-17 functions — the
// receive restatement, which is pinned by reference captures, decodes what they make — so they are restated as functions here and
// share encode() / the interleaver tables with tx11a.cpp.  What is specific to this transmitter and restated operation by operation:
//   * constellation levels of lutst/mapa_*.c: 10720 | 7580 | 3389, 10169 | 1654, 4962, 8270, 11578 (the brick mapper's differ by one LSB);
//   * pilots +-32 * IFFT_LUT_FACTOR16 = +-10720 (addpilot.h:4-48), polarity lutst/pilotsgn.c, SIGNAL with polarity 0;
//   * IFFT64x: zero-stuffed IFFT<128>, then << 2 (ifft64x.h:6-24); 32-sample cyclic prefix (CopyGI);
//   * Window (ofdmsymbol.h:19-46): 1/4, 1/2, 3/4 ramp on the first three samples, saturating add of the previous symbol's tail to the first
//     four, the first tail sample added to sample 0 once more (plain 16-bit add), next tail = 3/4, 1/2, 1/4 of samples 32..34;
//   * Copy_NT: arithmetic >> 6 and signed saturation to 8 bits (copynt.h:5-25); 8 trailing samples = the last window tail + zeros;
//   * the scrambler always starts from 0xFF (atx_tpl.h:19).
// The 640-sample preamble is a literal table in the reference (lutst/preamble40_11a.c) that no closed form regenerates exactly
// (a float-made table: +-1 against every candidate formula); it is handed in by the caller (tests/golden/preamble40_11a.i16).
// PINNED: reproduces usr/HwVeri/data/ofdm.bin (24 Mbps, 204-byte PSDU) sample for sample — tests/test_cpu_oracle_tx.py.
#include "tx11a.h"
#include "rx11a.h"
#include <cstring>
#include <vector>

namespace sbo {

void tx11a_encode_bits(const std::vector<uint8_t>& bytes, int cr, std::vector<uint8_t>& coded);   // tx11a.cpp: rate-1/2 mother code + puncturing, one coded bit per byte

namespace {
struct LRate { uint32_t kbps; uint8_t code; int nbpsc; int cr; int ndbps; };
const LRate LR[8] = {{6000, 0xB, 1, CR_12, 24}, {9000, 0xF, 1, CR_34, 36}, {12000, 0xA, 2, CR_12, 48}, {18000, 0xE, 2, CR_34, 72},
                     {24000, 0x9, 4, CR_12, 96}, {36000, 0xD, 4, CR_34, 144}, {48000, 0x8, 6, CR_23, 192}, {54000, 0xC, 6, CR_34, 216}};
inline int16_t level(int nbpsc, int bin) {              // bin = Gray-decoded index 0 .. 2^(nbpsc/2) - 1, smallest = most negative
    static const int16_t L2[2] = {-7580, 7580}, L4[4] = {-10169, -3389, 3389, 10169}, L6[8] = {-11578, -8270, -4962, -1654, 1654, 4962, 8270, 11578};
    return nbpsc == 2 ? L2[bin] : nbpsc == 4 ? L4[bin] : L6[bin];
}
inline int16_t sat_add16(int a, int b) { int s = a + b; return (int16_t)(s > 32767 ? 32767 : s < -32768 ? -32768 : s); }
inline int8_t pack_nt(int16_t v) { int s = v >> 6; return (int8_t)(s > 127 ? 127 : s < -128 ? -128 : s); }

struct LegacyTx {
    c16 last[4];                                        // info->cWindow[0..3]
    unsigned pilot_index = 0;
    // one OFDM symbol from ncbps coded bits (one per byte): interleave, map, pilots, IFFT64x, CopyGI, Window, Copy_NT
    void symbol(const uint8_t* coded, int nbpsc, bool pilot_neg, int8_t* out) {
        const Tables& T = tables();
        const int ncbps = 48 * nbpsc;
        const uint16_t* dm = nbpsc == 1 ? T.deint48 : nbpsc == 2 ? T.deint96 : nbpsc == 4 ? T.deint192 : T.deint288;
        uint8_t air[288];
        for (int k = 0; k < ncbps; k++) air[dm[k]] = coded[k];
        alignas(16) c16 f[64]; memset(f, 0, sizeof f);
        const int h = nbpsc / 2;
        auto lev = [&](const uint8_t* b) -> int16_t {                         // first bit on air is the most significant Gray bit
            int g = 0; for (int i = 0; i < h; i++) g = (g << 1) | b[i];
            int bin = 0, acc = 0; for (int i = h - 1; i >= 0; i--) { acc ^= (g >> i) & 1; bin = (bin << 1) | acc; }
            return level(nbpsc, bin);
        };
        int d = 0;
        for (int pass = 0; pass < 2; pass++)                                  // AddPilot's carrier order: -26..-1 then 1..26, pilots skipped
            for (int i = pass ? 1 : 38; i <= (pass ? 26 : 63); i++) {
                if (i == 43 || i == 57 || i == 7 || i == 21) continue;
                const uint8_t* b = air + d * nbpsc; d++;
                if (nbpsc == 1) { f[i].re = b[0] ? 10720 : -10720; f[i].im = 0; }
                else { f[i].re = lev(b); f[i].im = lev(b + h); }
            }
        const int16_t one = 32 * 335;                                         // OFDM_ONE
        const int s = pilot_neg ? -1 : 1;
        f[7].re = (int16_t)(s * one); f[21].re = (int16_t)(-s * one); f[57].re = (int16_t)(s * one); f[43].re = (int16_t)(s * one);
        alignas(16) c16 t[128], o[128]; c16 sym[160];
        memset(t, 0, sizeof t); memcpy(t, f, 32 * sizeof(c16)); memcpy(t + 96, f + 32, 32 * sizeof(c16));
        ifft128((v128*)t, (v128*)o);
        for (int i = 0; i < 128; i++) { sym[32 + i].re = (int16_t)((uint16_t)o[i].re << 2); sym[32 + i].im = (int16_t)((uint16_t)o[i].im << 2); }   // psllw 2
        memcpy(sym, sym + 128, 32 * sizeof(c16));                             // CopyGI
        // Window
        sym[0].re >>= 2; sym[0].im >>= 2; sym[1].re >>= 1; sym[1].im >>= 1;
        sym[2].re = (int16_t)(sym[2].re - (sym[2].re >> 2)); sym[2].im = (int16_t)(sym[2].im - (sym[2].im >> 2));
        for (int i = 0; i < 4; i++) { sym[i].re = sat_add16(sym[i].re, last[i].re); sym[i].im = sat_add16(sym[i].im, last[i].im); }
        sym[0].re = (int16_t)(sym[0].re + last[0].re); sym[0].im = (int16_t)(sym[0].im + last[0].im);
        last[0].re = (int16_t)(sym[32].re - (sym[32].re >> 2)); last[0].im = (int16_t)(sym[32].im - (sym[32].im >> 2));
        last[1].re = (int16_t)(sym[33].re >> 1); last[1].im = (int16_t)(sym[33].im >> 1);
        last[2].re = (int16_t)(sym[34].re >> 2); last[2].im = (int16_t)(sym[34].im >> 2);
        last[3].re = last[3].im = 0;
        for (int i = 0; i < 160; i++) { out[2 * i] = pack_nt(sym[i].re); out[2 * i + 1] = pack_nt(sym[i].im); }
    }
};
}

uint32_t tx11a_legacy_nsym(uint32_t psdu_len, uint32_t rate_kbps) {
    for (auto& r : LR) if (r.kbps == rate_kbps) return (16u + 6u + 8u * psdu_len + (uint32_t)r.ndbps - 1u) / (uint32_t)r.ndbps;
    return 0;
}

// mpdu / len: the frame body; append_crc: BB11ATxFrameMod appends CRC-32 (atx_tpl_imp.h:15, atx_tpl.h:42-47), BB11ATxBufferMod6M (the ACK
// path) sends the buffer as it is.  Returns complex samples written: 640 + 160 (1 + nsym) + 8.
size_t tx11a_legacy_modulate(const uint8_t* mpdu, uint32_t len, int append_crc, uint32_t rate_kbps, const c16* preamble640, int8_t* out, size_t cap) {
    const LRate* ri = nullptr; for (auto& r : LR) if (r.kbps == rate_kbps) ri = &r;
    if (!ri || !mpdu || !preamble640) return 0;
    const Tables& T = tables();
    const uint32_t L = len + (append_crc ? 4u : 0u);
    const uint32_t nsym = tx11a_legacy_nsym(L, rate_kbps);
    const size_t total = 640 + 160 * (size_t)(1 + nsym) + 8;
    if (total > cap || L > 4095u) return 0;
    LegacyTx tx;
    for (int i = 0; i < 640; i++) { out[2 * i] = pack_nt(preamble640[i].re); out[2 * i + 1] = pack_nt(preamble640[i].im); }   // CopyPreamble16_NT
    const c16* pt = preamble640 + 512;
    tx.last[0].re = (int16_t)(pt[0].re - (pt[0].re >> 2)); tx.last[0].im = (int16_t)(pt[0].im - (pt[0].im >> 2));
    tx.last[1].re = (int16_t)(pt[1].re >> 1); tx.last[1].im = (int16_t)(pt[1].im >> 1);
    tx.last[2].re = (int16_t)(pt[2].re >> 2); tx.last[2].im = (int16_t)(pt[2].im >> 2);
    tx.last[3].re = tx.last[3].im = 0;
    {   // SIGNAL (GetSignal, atx.h:78-98): rate code, LENGTH << 5, even parity in bit 17; one BPSK rate-1/2 symbol, pilot polarity 0
        uint32_t sig = ri->code | (L << 5);
        uint32_t p = sig ^ (sig >> 16); p ^= p >> 8; p ^= p >> 4; p ^= p >> 2; p ^= p >> 1; sig |= (p & 1u) << 17;
        std::vector<uint8_t> b = {(uint8_t)sig, (uint8_t)(sig >> 8), (uint8_t)(sig >> 16)}, coded;
        tx11a_encode_bits(b, CR_12, coded);
        tx.symbol(coded.data(), 1, false, out + 2 * 640);
    }
    // Scramble11a: SERVICE (two scrambler bytes), body, CRC-32, tail byte & 0xC0, scrambled zero padding; register starts at 0xFF
    const uint32_t nbytes = (nsym * (uint32_t)ri->ndbps + 7u) / 8u;
    std::vector<uint8_t> data(nbytes + 8, 0);
    memcpy(data.data() + 2, mpdu, len);
    if (append_crc) { uint32_t crc = 0xFFFFFFFFu; for (uint32_t i = 0; i < len; i++) crc = (crc >> 8) ^ T.crc32_lut[mpdu[i] ^ (crc & 0xFF)]; crc = ~crc; memcpy(data.data() + 2 + len, &crc, 4); }
    uint8_t reg = 0xFF;
    for (uint32_t i = 0; i < nbytes; i++) {
        reg = T.scramble_lut[reg >> 1];
        data[i] = (uint8_t)(data[i] ^ reg);
        if (i == 2 + L) data[i] &= 0xC0;
    }
    data.resize(nbytes);
    std::vector<uint8_t> coded; tx11a_encode_bits(data, ri->cr, coded);
    const int ncbps = 48 * ri->nbpsc;
    coded.resize((size_t)nsym * ncbps + 8, 0);
    const Tables& TT = tables(); (void)TT;
    static uint8_t pneg[127];
    static const bool have = [] { unsigned st = 0x7F; uint8_t seq[127]; for (int i = 0; i < 127; i++) { unsigned o = ((st >> 6) ^ (st >> 3)) & 1; st = ((st << 1) | o) & 0x7F; seq[i] = (uint8_t)o; }
                 for (int i = 0; i < 127; i++) pneg[i] = seq[(i + 1) % 127]; return true; }(); (void)have;      // lutst/pilotsgn.c: entry i = p_{i+1}
    unsigned pi = 0;
    for (uint32_t s = 0; s < nsym; s++) {
        tx.symbol(coded.data() + (size_t)s * ncbps, ri->nbpsc, pneg[pi] != 0, out + 2 * (640 + 160 * (size_t)(1 + s)));
        pi++; if (pi == 127) pi = 0;
    }
    int8_t* tail = out + 2 * (640 + 160 * (size_t)(1 + nsym));                // UpsampleTailAndCopyNT: Copy_NT(cWindow, 8)
    for (int i = 0; i < 4; i++) { tail[2 * i] = pack_nt(tx.last[i].re); tail[2 * i + 1] = pack_nt(tx.last[i].im); }
    memset(tail + 8, 0, 8);
    return total;
}

} // namespace sbo
