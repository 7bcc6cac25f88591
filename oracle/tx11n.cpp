// ORACLE — TEST INFRASTRUCTURE ONLY (see ops.h).  802.11n two-stream transmit restatement; see tx11n.h.
#include "tx11n.h"
#include "rx11a.h"
#include "rx11n.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace sbo {
namespace {
const int16_t MOD_BPSK = 30339, MOD_QPSK = 21453, MOD_QAM16 = 9594, MOD_QAM64 = 4681;   // fb11nmod_config.hpp:129-136 (TMap11aBPSK<30339>, TMap11aQPSK<21453>, TMap11aQAM16<9594>, TMap11aQAM64<4681>)
struct McsInfo { int nbpsc, cr, ndbps, enc_in, parse_in; };             // ieee80211const.h:35-55, conv_enc.hpp:66-67,152-153,245-246, streamparser.hpp:13,50
bool mcs_info(uint32_t mcs, McsInfo& m) {
    // the modulator graph carries MCS 8..14 (fb11nmod_config.hpp:146-155: encoders 1/2, 3/4, 2/3 behind TBB11nMRSelect; parsers streamparser.hpp:13,46,80,114)
    switch (mcs) {
        case 8: m = {1, CR_12, 52, 1, 13}; break;    case 9: m = {2, CR_12, 104, 1, 26}; break;   case 10: m = {2, CR_34, 156, 3, 26}; break;
        case 11: m = {4, CR_12, 208, 1, 52}; break;  case 12: m = {4, CR_34, 312, 3, 52}; break;
        case 13: m = {6, CR_23, 416, 2, 78}; break;  case 14: m = {6, CR_34, 468, 3, 78}; break;
        default: return false;
    }
    return true;
}
// K = 7 (133, 171) mother code on a bit string, LSB first, with the three puncture patterns (conv_enc.hpp:20-280)
void encode_bits(const std::vector<uint8_t>& bytes, int cr, std::vector<uint8_t>& coded, unsigned& s) {
    const size_t n = bytes.size() * 8;
    for (size_t i = 0; i < n; i++) {
        const unsigned x = (bytes[i >> 3] >> (i & 7)) & 1;
        const unsigned a = (x ^ (s >> 4) ^ (s >> 3) ^ (s >> 1) ^ s) & 1, b = (x ^ s ^ (s >> 3) ^ (s >> 4) ^ (s >> 5)) & 1;
        s = (s >> 1) | (x << 5);
        const size_t ph = cr == CR_34 ? i % 3 : cr == CR_23 ? i % 2 : 0;
        if (cr == CR_12 || ph == 0) { coded.push_back((uint8_t)a); coded.push_back((uint8_t)b); }
        else if (cr == CR_34) coded.push_back((uint8_t)(ph == 1 ? a : b));
        else coded.push_back((uint8_t)a);
    }
}
// T11Interleave<N_CBPS, N_BPSC, 13, 11, I_SS>: where coded bit k of a stream goes (interleave.hpp:33-60)
int ht_interleave_pos(int k, int ncbps, int nbpsc, int iss /*1, 2*/) {
    const int ncol = 13, nrot = 11, ns = nbpsc / 2 > 1 ? nbpsc / 2 : 1;
    const int i = ncbps / ncol * (k % ncol) + k / ncol;
    const int j = ns * (i / ns) + (i + ncbps - ncol * i / ncbps) % ns;
    return (ncbps + j - (((iss - 1) * 2) % 3 + 3 * ((iss - 1) / 3)) * nrot * nbpsc) % ncbps;
}
const uint8_t* pilot_neg() {                                            // _b_dot11_pilot.h:40-46 == pilot.hpp:10-28: entry i is p(i+1) of the 127-periodic polarity sequence
    static uint8_t t[128];
    static const bool done = [] { unsigned st = 0x7F; uint8_t seq[127]; for (int i = 0; i < 127; i++) { unsigned o = ((st >> 6) ^ (st >> 3)) & 1; st = ((st << 1) | o) & 0x7F; seq[i] = (uint8_t)o; }
                 for (int i = 0; i < 127; i++) t[i] = seq[(i + 1) % 127]; t[127] = 0; return true; }();
    (void)done; return t;
}
// TIFFTxOnly (fft.hpp:62-101): zero-stuffed IFFT<128>, no scaling; TAddGI (gi.hpp:32-41): last 32 samples in front.  csd_vec = TCSD<n>
void ifft_gi(const c16* f64, int csd_vec, c16* out160) {
    alignas(16) c16 t[128], o[128], sft[128];
    memset(t, 0, sizeof t); memcpy(t, f64, 32 * sizeof(c16)); memcpy(t + 96, f64 + 32, 32 * sizeof(c16));
    ifft128((v128*)t, (v128*)o);
    const c16* body = o;
    if (csd_vec) { for (int i = 0; i < 128; i++) sft[(i + 4 * csd_vec) % 128] = o[i]; body = sft; }     // csd.hpp:38-51: cyclic delay by csd_vec vectors of 4
    memcpy(out160, body + 96, 32 * sizeof(c16)); memcpy(out160 + 32, body, 128 * sizeof(c16));
}
}  // namespace

void tx11n_preamble_tables(c16* lstf, c16* lltf, c16* htstf, c16* htltf) {
    // One amplitude reproduces all four literal tables of the reference; it is fitted, not documented there: the tables only determine
    // it to 362.0592 +- 0.0001 (= 512.029 / sqrt 2).  Each table carries the same power: scale = A * sqrt(24 / sum |tone|^2).
    const double A = 362.0592, PI_ = 3.14159265358979323846;
    static const int8_t L[53] = {1,1,-1,-1,1,1,-1,1,-1,1,1,1,1,1,1,-1,-1,1,1,-1,1,-1,1,1,1,1,0,
                                 1,-1,-1,1,1,-1,1,-1,1,-1,-1,-1,-1,-1,1,1,-1,-1,1,-1,1,-1,1,1,1,1};
    double Sre[3][128] = {{0}}, Sim[3][128] = {{0}};
    static const int stf_k[12] = {-24, -20, -16, -12, -8, -4, 4, 8, 12, 16, 20, 24}; static const int stf_s[12] = {1, -1, 1, -1, -1, 1, -1, -1, 1, 1, 1, 1};
    for (int i = 0; i < 12; i++) { Sre[0][(stf_k[i] + 128) % 128] = stf_s[i]; Sim[0][(stf_k[i] + 128) % 128] = stf_s[i]; }
    for (int k = -26; k <= 26; k++) Sre[1][(k + 128) % 128] = L[k + 26];
    for (int k = -28; k <= 28; k++) Sre[2][(k + 128) % 128] = k == -28 || k == -27 ? 1 : k == 27 || k == 28 ? -1 : L[k + 26];
    auto gen = [&](int set, double scale, int n0, int count, c16* out) {
        for (int i = 0; i < count; i++) {
            const int n = n0 + i; double re = 0, im = 0;
            for (int k = 0; k < 128; k++) {
                if (Sre[set][k] == 0 && Sim[set][k] == 0) continue;
                const double ph = 2 * PI_ * (double)(((long long)k * n) % 128) / 128, c = cos(ph), s = sin(ph);
                re += Sre[set][k] * c - Sim[set][k] * s; im += Sre[set][k] * s + Sim[set][k] * c;
            }
            out[i].re = (int16_t)lround(re * scale); out[i].im = (int16_t)lround(im * scale);
        }
    };
    gen(0, A, 0, 320, lstf); gen(1, A * sqrt(24.0 / 52.0), -64, 320, lltf); gen(0, A, -32, 160, htstf); gen(2, A * sqrt(24.0 / 56.0), -32, 160, htltf);
}

uint32_t tx11n_nsym(uint32_t len, uint32_t mcs, uint32_t* signalled) {
    McsInfo m; if (!mcs_info(mcs, m)) return 0;
    const uint32_t bits = (len + 4u) * 8u + 16u + 6u, ns = (bits + m.ndbps - 1) / m.ndbps;             // ht_symbol_count
    if (signalled) *signalled = ns;
    uint32_t total_bytes = (ns * m.ndbps + 7u) / 8u;                                                   // TBB11nSrc: service + frame + FCS + tail + pad (PHY_11n.hpp:88-134)
    total_bytes = (total_bytes + m.enc_in - 1) / m.enc_in * m.enc_in;                                   // MRSelect's FlushPort pads the encoder's input burst
    const uint32_t coded_bytes = m.cr == CR_12 ? 2u * total_bytes : m.cr == CR_34 ? total_bytes / 3u * 4u : total_bytes / 2u * 3u;
    return (coded_bytes + m.parse_in - 1) / m.parse_in;                                                // the encoder's FlushPort pads the stream parser's burst
}

size_t tx11n_modulate(const uint8_t* payload, uint32_t len, uint32_t mcs, uint8_t sc_seed, c16* out0, c16* out1, size_t cap) {
    McsInfo m; if (!mcs_info(mcs, m) || len + 4u > 4095u) return 0;
    uint32_t ns_sig = 0; const uint32_t nsym = tx11n_nsym(len, mcs, &ns_sig);
    const size_t total = 640 + 480 + 480 + 160 * (size_t)nsym;
    if (total > cap) return 0;
    const Tables& T = tables();
    c16* out[2] = {out0, out1};
    // ---- L-STF + L-LTF (preamble11n.hpp:22-37): stream 2 delayed cyclically by 2 vectors = 8 samples (200 ns) ----------------------------
    c16 lstf[320], lltf[320], htstf[160], htltf[160]; tx11n_preamble_tables(lstf, lltf, htstf, htltf);
    memcpy(out[0], lstf, sizeof lstf); memcpy(out[0] + 320, lltf, sizeof lltf);
    for (int i = 0; i < 320; i++) out[1][(i + 8) % 320] = lstf[i];                                     // _b_lstf.h:22-36: the whole 320 rotate
    for (int i = 0; i < 256; i++) out[1][320 + 64 + i] = lltf[64 + i - 8];                             // _b_lltf.h:22-38: the two symbols read 8 samples early ...
    for (int i = 0; i < 64; i++) out[1][320 + i] = out[1][320 + 256 + i];                              // ... and the guard interval is rebuilt from their end
    // ---- L-SIG + HT-SIG (PHY_11n.hpp:244-281, _b_lsig.h, _b_htsig.h) -> TConvEncode_12 -> T11aInterleaveBPSK -> TSigMap11n ----------------
    size_t pos = 640;
    {
        const uint32_t nsym_all = ns_sig + 5u, lsig_len = (nsym_all * 24u - 16u - 6u) / 8u;
        uint32_t lsig = 0xBu | (lsig_len << 5); uint32_t p = lsig ^ (lsig >> 16); p ^= p >> 8; p ^= p >> 4; p ^= p >> 2; p ^= p >> 1; lsig |= (p & 1u) << 17;
        uint8_t hs[6] = {(uint8_t)mcs, (uint8_t)(len + 4u), (uint8_t)((len + 4u) >> 8), 3, 0, 0};
        const uint8_t c8 = crc8_htsig(hs, 4, 2); hs[4] |= (uint8_t)(c8 << 2); hs[5] |= (uint8_t)(c8 >> 6);
        std::vector<uint8_t> b = {(uint8_t)lsig, (uint8_t)(lsig >> 8), (uint8_t)(lsig >> 16), hs[0], hs[1], hs[2], hs[3], hs[4], hs[5]}, coded;
        unsigned st = 0; encode_bits(b, CR_12, coded, st);
        const uint8_t* neg = pilot_neg(); unsigned pidx = 127;
        for (int s3 = 0; s3 < 3; s3++) {
            uint8_t air[48]; for (int k = 0; k < 48; k++) air[T.deint48[k]] = coded[48 * s3 + k];
            alignas(16) c16 f[64]; memset(f, 0, sizeof f); int d = 0;
            for (int pass = 0; pass < 2; pass++)
                for (int i = pass ? 1 : 38; i <= (pass ? 26 : 63); i++) {
                    if (i == 43 || i == 57 || i == 7 || i == 21) continue;
                    const int16_t v = air[d++] ? MOD_BPSK : (int16_t)-MOD_BPSK;
                    if (s3 == 0) f[i].re = v; else f[i].im = v;                                        // mapper11n.hpp:33-43: L-SIG on I, HT-SIG on Q
                }
            const int sg = neg[pidx] ? -1 : 1;                                                         // pilot.hpp:95-115 with BPSK_MOD = 30339
            f[7].re = (int16_t)(sg * MOD_BPSK); f[21].re = (int16_t)(-sg * MOD_BPSK); f[57].re = (int16_t)(sg * MOD_BPSK); f[43].re = (int16_t)(sg * MOD_BPSK);
            pidx++; if (pidx >= 127) pidx = 0;                                                         // 127 (L-SIG: p0), 0, 1
            ifft_gi(f, 0, out[0] + pos); ifft_gi(f, 2, out[1] + pos); pos += 160;                      // TTeeEx: stream 2 through TCSD<2>
        }
    }
    // ---- HT-STF + HT-LTF x 2 (preamble11n.hpp:56-76): P = [[1, -1], [1, 1]], stream 2 delayed by 4 vectors = 16 samples (400 ns) -----------
    memcpy(out[0] + pos, htstf, sizeof htstf); memcpy(out[0] + pos + 160, htltf, sizeof htltf);
    for (int i = 0; i < 160; i++) { out[0][pos + 320 + i].re = (int16_t)-htltf[i].re; out[0][pos + 320 + i].im = (int16_t)-htltf[i].im; }     // _b_htltf.h:49-62
    for (int i = 0; i < 160; i++) out[1][pos + (i + 16) % 160] = htstf[i];                             // _b_htstf.h:22-36
    for (int r = 0; r < 2; r++) {                                                                      // _b_htltf.h:75-94: get_ltf_21 == get_ltf_22
        c16* o = out[1] + pos + 160 + 160 * r;
        for (int i = 0; i < 128; i++) o[32 + i] = htltf[32 + i - 16];
        for (int i = 0; i < 32; i++) o[i] = o[128 + i];
    }
    pos += 480;
    // ---- DATA: TBB11nSrc -> T11aSc -> TBB11nMRSelect -> encoder -> stream parser -> 2 x (interleaver, mapper, T11nAddPilot, IFFT, [CSD], GI) -----
    uint32_t crc = 0xFFFFFFFFu; for (uint32_t i = 0; i < len; i++) crc = (crc >> 8) ^ T.crc32_lut[payload[i] ^ (crc & 0xFF)]; crc = ~crc;
    const uint32_t src_bytes = (ns_sig * m.ndbps + 7u) / 8u;
    std::vector<uint8_t> data(src_bytes, 0);
    if (len) memcpy(data.data() + 2, payload, len);
    memcpy(data.data() + 2 + len, &crc, 4);
    uint8_t reg = sc_seed; const uint32_t tail_at = 2 + len + 4;
    for (uint32_t i = 0; i < src_bytes; i++) { reg = T.scramble_lut[reg >> 1]; data[i] = (uint8_t)(data[i] ^ reg); if (i == tail_at) data[i] &= 0xC0; }   // scramble.hpp:169-262
    while (data.size() % m.enc_in) data.push_back(0);                                                  // FlushPort pad: unscrambled zero bytes into the encoder
    std::vector<uint8_t> coded; unsigned st = 0; encode_bits(data, m.cr, coded, st);
    const int ncbpss = 52 * m.nbpsc, per_sym = 2 * ncbpss;
    coded.resize((size_t)nsym * per_sym, 0);                                                           // the encoder's FlushPort pad, then whole symbols
    const uint8_t* neg = pilot_neg();
    static const int8_t PIL[4][2][4] = {{{1, 1, -1, -1}, {1, -1, -1, 1}}, {{1, -1, -1, 1}, {-1, -1, 1, 1}}, {{-1, -1, 1, 1}, {-1, 1, 1, -1}}, {{-1, 1, 1, -1}, {1, 1, -1, -1}}};   // _b_dot11_pilot.h:29-35
    for (uint32_t s = 0; s < nsym; s++) {
        const uint8_t* cb = coded.data() + (size_t)s * per_sym;
        for (int iss = 0; iss < 2; iss++) {
            uint8_t air[312];
            const int S = m.nbpsc / 2 > 1 ? m.nbpsc / 2 : 1;                                           // _b_stream_parser.h:36-49,140-198,200-275: S bits to stream 1, S to stream 2, round robin
            for (int k = 0; k < ncbpss; k++) air[ht_interleave_pos(k, ncbpss, m.nbpsc, iss + 1)] = cb[(2 * (k / S) + iss) * S + k % S];
            auto level = [&](const uint8_t* b, int mbits, int16_t mod) -> int16_t {                    // mapper11a.hpp:16-41 InitQamMapLut: bit-reverse, Gray -> binary, (2b - (2^M - 1)) * MOD
                unsigned g = 0; for (int t = 0; t < mbits; t++) g = (g << 1) | b[t];                   // first bit on air = LSB of the LUT index = MSB after BitReverseN
                unsigned bin = g; for (unsigned sh = g >> 1; sh; sh >>= 1) bin ^= sh;
                return (int16_t)(((int)bin * 2 - ((1 << mbits) - 1)) * mod);
            };
            alignas(16) c16 f[64]; memset(f, 0, sizeof f); int d = 0;
            for (int pass = 0; pass < 2; pass++)                                                       // pilot_11n.hpp:54-66: -28..-1 then 1..28 without +-7, +-21
                for (int i = pass ? 1 : 36; i <= (pass ? 28 : 63); i++) {
                    if (i == 43 || i == 57 || i == 7 || i == 21) continue;
                    if (m.nbpsc == 1) f[i].re = air[d] ? MOD_BPSK : (int16_t)-MOD_BPSK;
                    else if (m.nbpsc == 2) { f[i].re = air[2 * d] ? MOD_QPSK : (int16_t)-MOD_QPSK; f[i].im = air[2 * d + 1] ? MOD_QPSK : (int16_t)-MOD_QPSK; }
                    else if (m.nbpsc == 4) { f[i].re = level(air + 4 * d, 2, MOD_QAM16); f[i].im = level(air + 4 * d + 2, 2, MOD_QAM16); }
                    else { f[i].re = level(air + 6 * d, 3, MOD_QAM64); f[i].im = level(air + 6 * d + 3, 3, MOD_QAM64); }
                    d++;
                }
            const int sg = neg[(s + 3) % 127] ? -1 : 1; const int8_t* pp = PIL[s & 3][iss];           // _b_dot11_pilot.h:7-17
            f[43].re = (int16_t)(sg * pp[0] * MOD_BPSK); f[57].re = (int16_t)(sg * pp[1] * MOD_BPSK); f[7].re = (int16_t)(sg * pp[2] * MOD_BPSK); f[21].re = (int16_t)(sg * pp[3] * MOD_BPSK);
            ifft_gi(f, iss ? 4 : 0, out[iss] + pos);                                                   // fb11nmod_config.hpp:122 TCSD<4> on stream 2
        }
        pos += 160;
    }
    return total;
}
}  // namespace sbo
