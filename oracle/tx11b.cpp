// ORACLE — TEST INFRASTRUCTURE ONLY (see ops.h).  802.11b transmit restatement; see tx11b.h.
#include "tx11b.h"
#include "tables.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace sbo {
namespace {
struct c8 { int8_t re, im; };
inline c8 mul(c8 a, c8 b) { return c8{(int8_t)(a.re * b.re - a.im * b.im), (int8_t)(a.re * b.im + a.im * b.re)}; }
inline c8 neg(c8 a) { return c8{(int8_t)-a.re, (int8_t)-a.im}; }
inline bool same(c8 a, c8 b) { return a.re == b.re && a.im == b.im; }
const int BARKER[11] = {1, -1, 1, 1, -1, 1, 1, 1, -1, -1, -1};                         // barkerspread.hpp:7
const c8 DQPSK_ENC[4] = {{1, 0}, {0, -1}, {0, 1}, {-1, 0}};                             // cck.hpp:766
const c8 CCK11_D[4] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};                               // cck.hpp:767
const c8 CCK5_D[4][8] = {                                                               // cck.hpp:768-773
    {{0, 1}, {1, 0}, {0, 1}, {-1, 0}, {0, 1}, {1, 0}, {0, -1}, {1, 0}},
    {{0, -1}, {1, 0}, {0, -1}, {-1, 0}, {0, -1}, {1, 0}, {0, 1}, {1, 0}},
    {{0, -1}, {-1, 0}, {0, -1}, {1, 0}, {0, 1}, {1, 0}, {0, -1}, {1, 0}},
    {{0, 1}, {-1, 0}, {0, 1}, {1, 0}, {0, -1}, {1, 0}, {0, 1}, {1, 0}}};
inline unsigned find_dqpsk(c8 v) { for (unsigned i = 0; i < 4; i++) if (same(DQPSK_ENC[i], v)) return i; return 4; }   // tpltrick.h:5

uint16_t crc16_ccitt(const uint8_t* p, unsigned n) {                                    // core/inc/CRC16.h:37-48 (reflected 0x8408, init FFFF, inverted)
    uint16_t c = 0xFFFF;
    for (unsigned i = 0; i < n; i++) { c ^= p[i]; for (int k = 0; k < 8; k++) c = (c & 1) ? (uint16_t)((c >> 1) ^ 0x8408) : (uint16_t)(c >> 1); }
    return (uint16_t)~c;
}
uint32_t crc32_of(const uint8_t* p, unsigned n) {                                       // core/inc/CRC32.h:82-93
    const Tables& T = tables(); uint32_t c = 0xFFFFFFFFu;
    for (unsigned i = 0; i < n; i++) c = (c >> 8) ^ T.crc32_lut[(p[i] ^ c) & 0xFF];
    return ~c;
}
uint8_t rate_code(uint32_t kbps) { return kbps == 1000 ? 0x0A : kbps == 2000 ? 0x14 : kbps == 5500 ? 0x37 : kbps == 11000 ? 0x6E : 0; }   // bb/bbb.h:47-50, DataRate.h:40-43
unsigned chips_per_byte(uint32_t kbps) { return kbps == 1000 ? 88 : kbps == 2000 ? 44 : kbps == 5500 ? 16 : kbps == 11000 ? 8 : 0; }

double shaper_h(int i) {                                                                // pulse.hpp:292-300, with the file's own PI
    const double PI_ = 3.141593;
    return (i == 1 || i == -1) ? 1.0 : 4 * cos(PI_ * i / 2) / PI_ / (1 - i * i);
}
}  // namespace

void tx11b_taps(int16_t* out20) { for (int k = 0; k < 20; k++) out20[k] = (int16_t)(shaper_h(8 - k) * 80 + .5); }

uint32_t tx11b_nchips(uint32_t len, uint32_t rate_kbps) {
    const unsigned cpb = chips_per_byte(rate_kbps); if (!cpb) return 0;
    return 24u * 88u + (len + 4u) * cpb;
}
uint32_t tx11b_nsamples(uint32_t len, uint32_t rate_kbps) {
    const uint32_t nc = tx11b_nchips(len, rate_kbps); if (!nc) return 0;
    return ((nc + 5u) * 4u + 7u) / 8u * 8u;
}

size_t tx11b_modulate(const uint8_t* payload, uint32_t len, uint32_t rate_kbps, uint32_t init_phase, int8_t* out, size_t cap, uint32_t* final_phase) {
    const uint8_t code = rate_code(rate_kbps);
    if (!code || len + 4u > 4095u) return 0;
    const size_t nsamp = tx11b_nsamples(len, rate_kbps);
    if (cap < nsamp) return 0;
    // ---- TBB11bSrc: sync, SFD, PLCP header, MPDU, CRC-32 (PHY_11b.hpp:82-151) -------------------------------------
    const uint32_t size = len + 4u;
    std::vector<uint8_t> bytes(24 + size);
    memset(bytes.data(), 0xFF, 16); bytes[16] = 0xA0; bytes[17] = 0xF3;                 // DOT11B_PLCP_LONG_PREAMBLE_SFD 0xF3A0, little endian
    uint32_t plen, ext = 0;
    if (rate_kbps == 1000) plen = size << 3;
    else if (rate_kbps == 2000) plen = size << 2;
    else if (rate_kbps == 5500) plen = ((size << 4) - 1) / 11 + 1;
    else { plen = ((size << 3) - 1) / 11 + 1; if (plen * 11 - (size << 3) >= 8) ext = 1; }
    bytes[18] = code; bytes[19] = (uint8_t)(ext << 7); bytes[20] = (uint8_t)plen; bytes[21] = (uint8_t)(plen >> 8);
    const uint16_t hc = crc16_ccitt(&bytes[18], 4); bytes[22] = (uint8_t)hc; bytes[23] = (uint8_t)(hc >> 8);
    if (len) memcpy(&bytes[24], payload, len);
    const uint32_t fcs = crc32_of(payload, len);
    for (int i = 0; i < 4; i++) bytes[24 + len + i] = (uint8_t)(fcs >> (8 * i));
    // ---- TSc741 (scramble.hpp:24-40,80-86): bit-serial form of the 256x128 table, seed 0x6C ------------------------
    uint8_t reg = 0x6C;
    for (auto& b : bytes) {
        uint8_t x = b, s = reg, o = 0;
        for (int k = 0; k < 8; k++) { const uint8_t o1 = (uint8_t)((x ^ s ^ (s >> 3)) & 1); s = (uint8_t)((s >> 1) | (o1 << 6)); o = (uint8_t)((o >> 1) | (o1 << 7)); x >>= 1; }
        b = o; reg = (uint8_t)(o >> 1);
    }
    // ---- TBB11bMRSelect + the four spreaders: chips at 11 Mchip/s, each component in {-1, 0, 1} --------------------
    std::vector<c8> chips; chips.reserve(tx11b_nchips(len, rate_kbps));
    uint32_t last_phase = init_phase; unsigned even = 0;                                // CF_DifferentialMap::last_phase; cck.hpp:834 bEven
    int preamble_cnt = 24;                                                              // PHY_11b.hpp:222-231
    for (uint8_t b : bytes) {
        if (preamble_cnt || rate_kbps == 1000) {                                        // barkerspread.hpp:25-41,55-65,91-108
            preamble_cnt--;
            uint8_t phase = (uint8_t)(last_phase & 1), codeb = 0, x = b;
            for (int k = 0; k < 8; k++) { if (x & 1) phase ^= 1; codeb |= (uint8_t)(phase << k); x >>= 1; }
            // (the table's second column is the complement of the first: the same recurrence started from phase 1)
            last_phase = codeb >> 7; last_phase = (last_phase << 1) | last_phase;
            for (int j = 0; j < 8; j++) { const int c = ((codeb >> j) & 1) ? -1 : 1; for (int k = 0; k < 11; k++) chips.push_back(c8{(int8_t)(c * BARKER[k]), 0}); }
        } else if (rate_kbps == 2000) {                                                 // barkerspread.hpp:129-156,160-188,214-222
            static const uint8_t rotate[4][4] = {{0, 1, 2, 3}, {1, 3, 0, 2}, {2, 0, 3, 1}, {3, 2, 1, 0}};
            uint8_t phase = (uint8_t)(last_phase & 3), codeb = 0, x = b;
            for (int k = 0; k < 8; k += 2) { phase = rotate[phase][x & 3]; codeb |= (uint8_t)(phase << k); x >>= 2; }
            last_phase = codeb >> 6;
            for (int j = 0; j < 4; j++) {
                const c8 m = DQPSK_ENC[(codeb >> (2 * j)) & 3];                         // 0: 1, 1: -j, 2: +j, 3: -1
                for (int k = 0; k < 11; k++) chips.push_back(c8{(int8_t)(m.re * BARKER[k]), (int8_t)(m.im * BARKER[k])});
            }
        } else if (rate_kbps == 5500) {                                                 // cck.hpp:893-918,945-953
            c8 v[16]; const unsigned prev = last_phase & 3;
            unsigned half = b % 16u; c8 m1 = DQPSK_ENC[half % 4], m0 = DQPSK_ENC[prev];
            for (int i = 0; i < 8; i++) v[i] = mul(mul(m0, m1), CCK5_D[half / 4][i]);
            half = b / 16u; m1 = DQPSK_ENC[half % 4]; m0 = v[7];
            for (int i = 0; i < 8; i++) v[8 + i] = mul(mul(neg(m0), m1), CCK5_D[half / 4][i]);
            last_phase = find_dqpsk(v[15]);
            for (int i = 0; i < 16; i++) chips.push_back(v[i]);
        } else {                                                                        // cck.hpp:797-828,854-866
            const c8 m0 = DQPSK_ENC[last_phase & 3], m1 = DQPSK_ENC[b % 4], m2 = CCK11_D[(b >> 2) % 4], m3 = CCK11_D[(b >> 4) % 4], m4 = CCK11_D[(b >> 6) % 4];
            const c8 a = mul(m0, m1); c8 v[8];
            v[0] = mul(mul(mul(a, m2), m3), m4); v[1] = mul(mul(a, m3), m4); v[2] = mul(mul(a, m2), m4); v[3] = neg(mul(a, m4));
            v[4] = mul(mul(a, m2), m3); v[5] = mul(a, m3); v[6] = neg(mul(a, m2)); v[7] = a;
            if (even) for (auto& c : v) c = neg(c);                                    // odd-numbered symbols carry an extra pi
            last_phase = find_dqpsk(v[7]); even ^= 1;
            for (int i = 0; i < 8; i++) chips.push_back(v[i]);
        }
    }
    if (final_phase) *final_phase = last_phase;
    // ---- TQuickPulseShaper (pulse.hpp:260-379): out[n][k] = sum_j x[n-j] * h(8 - 4j - k), then five zero-input flush vectors ----
    int16_t h[20]; tx11b_taps(h);
    const size_t nc = chips.size();
    std::vector<int16_t> y((nc + 5) * 8);
    for (size_t n = 0; n < nc + 5; n++)
        for (int k = 0; k < 4; k++) {
            int16_t re = 0, im = 0;
            for (int j = 0; j < 5; j++) { if (n < (size_t)j || n - j >= nc) continue; const c8 x = chips[n - j]; re = (int16_t)(re + x.re * h[4 * j + k]); im = (int16_t)(im + x.im * h[4 * j + k]); }
            y[(4 * n + k) * 2] = re; y[(4 * n + k) * 2 + 1] = im;
        }
    // ---- TPackSample16to8 (stdbrick.hpp:413-445, packsswb) on bursts of 8; Flush pads the last burst with zeros ----------------
    memset(out, 0, nsamp * 2);
    for (size_t i = 0; i < (nc + 5) * 8; i++) { const int16_t v = y[i]; out[i] = (int8_t)(v > 127 ? 127 : v < -128 ? -128 : v); }
    return nsamp;
}
}  // namespace sbo
