// ORACLE — TEST INFRASTRUCTURE ONLY (see ops.h).
// CPU restatement of the reference's 802.11a brick transmit graphs
//   kernel/bb/demod11/fb11amod_config.hpp:75-118 (CreateModGraph11a_40M) and :150-158 (CreatePreamble11a_40M),
//   driven like kernel/bb/demod11/fb11a_mod.cpp:27-107 (Test11A_FB_Mod: preamble, then SIGNAL + DATA, 32 trailing zero samples).
// PARITY UNPINNED: the reference holds no output vector of this brick graph.  usr/HwVeri/data/ofdm.bin (tests/golden/ofdm.bin) comes
// from the legacy transmitter (different window and IFFT rounding): the restatement matches it to +-1 away from symbol edges,
// which is evidence, not a pin; what is checked is round trip through the pinned receive oracle and table identities.
#pragma once
#include "tables.h"
#include <stddef.h>

namespace sbo {
// preamble as TTS11aSrc builds it (Brick11/src/preamble11a.hpp:22-104), 640 x COMPLEX16 before the 16->8 bit pack
void tx11a_preamble16(c16* out640);
// whole PPDU as COMPLEX8 pairs (re, im): 640 + 160 * (1 + nsym) samples followed by `tail_zeros` zero samples.
// payload = MPDU without FCS (the modulator appends CRC-32).  Returns the number of complex samples written (0 on bad arguments).
size_t tx11a_modulate(const uint8_t* payload, uint32_t len, uint32_t rate_kbps, uint8_t sc_seed, int8_t* out, size_t cap_samples, uint32_t tail_zeros);
// symbol count the reference transmits (TBB11aSrc::GetPadingByte, PHY_11a.hpp:107-123): the tail is counted as a whole byte and
// 9 Mbps pads to a pair of symbols, so this can exceed the standard's N_SYM by one
uint32_t tx11a_nsym(uint32_t len, uint32_t rate_kbps);
// the LEGACY transmitter (tx11a_legacy.cpp): BB11ATxFrameMod / BB11ATxBufferMod6M; pinned by usr/HwVeri/data/ofdm.bin
size_t tx11a_legacy_modulate(const uint8_t* mpdu, uint32_t len, int append_crc, uint32_t rate_kbps, const c16* preamble640, int8_t* out, size_t cap_samples);
uint32_t tx11a_legacy_nsym(uint32_t psdu_len, uint32_t rate_kbps);
}
