// Context facades used at the GPU brick boundary, field-for-field as the reference declares them:
//   CF_Error, CF_MemSamples      kernel/brick/inc/stdfacade.h:14-41
//   CF_RxFrameBuffer             kernel/bb/Brick11/src/ieee80211facade.hpp:106-114
//   CF_11aRxVector               kernel/bb/Brick11/src/ieee80211facade.hpp:213-238
//   CF_11CCA                     kernel/bb/Brick11/src/ieee80211facade.hpp:21-45
//   CF_CFOffset                  kernel/bb/Brick11/src/ieee80211facade.hpp:168-175
//   CF_11bRxVector               kernel/bb/Brick11/src/ieee80211facade.hpp:72-76
//   CF_HTRxVector                kernel/bb/Brick11/src/ieee80211facade.hpp:269-272
//   CF_11aTxVector, CF_TxFrameBuffer, CF_ScramblerSeed   ieee80211facade.hpp:140-146, 87-92, 116-119
//   CF_11bTxVector, CF_DifferentialMap                   ieee80211facade.hpp:78-85, 94-98
//   CF_11nTxVector                                       ieee80211facade.hpp:262-267
// plus the E_ERROR_* codes (stdfacade.h:10-12, ieee80211facade.hpp:10-19).
#pragma once
#include "brick.hpp"

#define E_ERROR_SUCCESS          0x00000000
#define E_ERROR_FRAME_OK         0x00000001
#define E_ERROR_PARAMETER        0x80000001
#define E_ERROR_PLCP_HEADER_FAIL 0x80000005
#define E_ERROR_CRC32_FAIL       0x80000006
#define E_ERROR_CS_TIMEOUT       0x80000007
#define E_ERROR_NOT_SUPPORTED    0x80000003
#define E_ERROR_DATARATE         0x80000002
#define E_ERROR_FAILED           0x8000FFFF
#define BK_ERROR_FAILED          0x8000FFFF
#define CR_12 0
#define CR_23 1
#define CR_34 2

class CF_Error { FACADE_FIELD(ulong, error_code) public: CF_Error() { error_code() = E_ERROR_SUCCESS; } };

class CF_MemSamples {
    FACADE_FIELD(COMPLEX16*, mem_sample_buf) FACADE_FIELD(uint, mem_sample_buf_size) FACADE_FIELD(uint, mem_sample_start_pos)
    FACADE_FIELD(uint, mem_sample_count) FACADE_FIELD(uint, mem_sample_index)
public:
    void Init(COMPLEX16* sbuf, uint sbuf_size) {
        mem_sample_buf() = sbuf; mem_sample_buf_size() = sbuf_size; mem_sample_start_pos() = 0;
        mem_sample_count() = sbuf_size / sizeof(COMPLEX16); mem_sample_index() = 0;
    }
};
class CF_RxFrameBuffer {
    FACADE_FIELD(uchar*, rx_frame_buf) FACADE_FIELD(uint, rx_frame_buf_size)
public:
    void Init(uchar* fbuf, uint fbuf_size) { rx_frame_buf() = fbuf; rx_frame_buf_size() = fbuf_size; }
};
class CF_11aRxVector {
    FACADE_FIELD(ushort, frame_length) FACADE_FIELD(ushort, total_symbols) FACADE_FIELD(ushort, remain_symbols)
    FACADE_FIELD(ulong, data_rate_kbps) FACADE_FIELD(ushort, code_rate) FACADE_FIELD(ulong, crc32)
public:
    void Reset() { frame_length() = 0; total_symbols() = 0; remain_symbols() = 0; data_rate_kbps() = 6000; code_rate() = CR_12; crc32() = 0; }
    CF_11aRxVector() { Reset(); }
};
class CF_11CCA {
public:
    typedef enum { power_clear = 0, power_detected } CCAState;
    FACADE_FIELD(CCAState, cca_state) FACADE_FIELD(uint, cca_pwr_threshold) FACADE_FIELD(uint, cca_pwr_reading) FACADE_FIELD(uint, cca_peak_index)
public:
    CF_11CCA() { cca_pwr_threshold() = 1000 * 1000; cca_pwr_reading() = 0; cca_state() = power_clear; cca_peak_index() = 0; }
    void Reset() { cca_state() = power_clear; }
    void OnPowerDetected() {}
};
class CF_CFOffset { FACADE_FIELD(short, CFO_est) public: void Reset() { CFO_est() = 0; } };
class CF_11bRxVector { FACADE_FIELD(ushort, frame_length) FACADE_FIELD(ulong, data_rate_kbps) FACADE_FIELD(ulong, crc32) };
class CF_HTRxVector { FACADE_FIELD(ushort, ht_frame_length) FACADE_FIELD(ulong, ht_frame_mcs) };
class CF_11aTxVector { FACADE_FIELD(ushort, frame_length) FACADE_FIELD(ulong, data_rate_kbps) FACADE_FIELD(ulong, crc32) FACADE_FIELD(ushort, coding_rate) };
class CF_TxFrameBuffer { FACADE_FIELD(uchar*, mpdu_buf0) FACADE_FIELD(ushort, mpdu_buf_size0) FACADE_FIELD(uchar*, mpdu_buf1) FACADE_FIELD(ushort, mpdu_buf_size1) };
class CF_ScramblerSeed { FACADE_FIELD(uchar, sc_seed) };
class CF_11bTxVector { FACADE_FIELD(ushort, frame_length) FACADE_FIELD(uchar, preamble_type) FACADE_FIELD(uchar, mod_select) FACADE_FIELD(ulong, data_rate_kbps) FACADE_FIELD(ulong, crc32) };
class CF_DifferentialMap { FACADE_FIELD(uint, last_phase) };
class CF_11nTxVector { FACADE_FIELD(ushort, frame_length) FACADE_FIELD(ulong, crc32) FACADE_FIELD(ushort, coding_rate) FACADE_FIELD(ushort, mcs_index) };
