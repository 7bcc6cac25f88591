// ORACLE — TEST INFRASTRUCTURE ONLY (see ops.h).
#pragma once
#include "tables.h"
#include <vector>
#include <stddef.h>

namespace sbo {

enum { CR_12 = 0, CR_23 = 1, CR_34 = 2 };   // kernel/bb/Brick11/src/ieee80211const.h:13-18

struct ViterbiCore {
    std::vector<v128> col;      // trellis columns, 4 x 16 uint8 metrics each (LSB = survivor mark)
    v128* cur = nullptr;
    uint32_t steps = 0;
    uint32_t max_steps = 5000 * 8;   // fb11ademod_config.hpp:175 T11aViterbi<5000*8,48,256>
    void reset();
    void step_ab(unsigned sa, unsigned sb);
    void step_one(bool use_b, unsigned s);
    void normalize();
    void traceback(uint8_t* out, uint32_t nbits, uint32_t lookahead);
};

uint32_t viterbi_signal(const uint8_t soft[48]);
size_t viterbi_decode_block(ViterbiCore& v, const uint8_t* soft, size_t nsoft, int code_rate,
                            uint32_t frame_len_bytes, uint32_t depth, uint32_t look, uint8_t* out, uint32_t& ob_count);

} // namespace sbo
