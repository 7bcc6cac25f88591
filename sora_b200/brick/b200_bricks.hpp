// GPU bricks: drop-in replacements for CPU sub-graphs of the reference's dot11 demod graphs, forwarding to the
// C ABI in include/sora_b200.h.  Header-only templates, compiled into the caller exactly like any other brick.
//
//   TB200Dot11aRx<T_CTX, T_NEXT>   replaces ds2 .. fsink of CreateDemodGraph11a_40M
//       (kernel/bb/demod11/fb11ademod_config.hpp:169-242: TDownSample2, TBB11bRxSwitch, TDCRemoveEx, TCCA11a,
//        TDCEstimator, T11aLTS, T11aDataSymbol, TFreqCompensation, TFFT64, TChannelEqualization, TPhaseCompensate,
//        TPilotTrack, TBB11aRxRateSel, T11aDemap*, T11aDeinterleave*, T11aViterbiSig, T11aPLCPParser,
//        TThreadSeparator, T11aViterbi, T11aDesc, TBB11aFrameSink)
//     iport  COMPLEX16 x 28  (what TMemSamples / TRxStream emit: memsource.hpp:33, rxstream.hpp)
//     oport  uchar x 1       (decoded PSDU bytes, as T11aDesc would hand to the frame sink)
//     context facades: CF_Error, CF_11aRxVector, CF_RxFrameBuffer, CF_11CCA, CF_CFOffset — written exactly when and
//     how the CPU bricks write them: after a frame event error_code() becomes E_ERROR_FRAME_OK / E_ERROR_CRC32_FAIL /
//     E_ERROR_PLCP_HEADER_FAIL and the driver loop (fb11a_demod.cpp:29-81) does its Flush()/Reset() as before.
//
//   TB200Dot11bRx<T_CTX, T_NEXT>   replaces everything behind the source of the 802.11b graph
//       (kernel/bb/demod11/fb11bdemod_config.hpp:123-190: TDCRemove, TEnergyDetect, TSymTiming, TBarkerSync, TBB11bDespread,
//        TSFDSync, TBB11bPlcpParser, TDBPSKDemap/TDQPSKDemap, TCCK5P5Decoder/TCCK11Decoder, TDesc741, TBB11bFrameSink)
//     iport COMPLEX16 x 28 at 44 Msps, oport uchar x 1; facades CF_Error, CF_11bRxVector, CF_RxFrameBuffer.
//   TB200Dot11nRx<T_CTX, T_NEXT>   replaces ds2 .. fsink of CreateDemodGraph11n (kernel/bb/demod11/fb11ndemod_config.hpp:167-262)
//     iport COMPLEX16 x 28 x 2 streams (what TMemSamples2 emits, memsource.hpp:189-244), oport uchar x 1;
//     facades CF_Error, CF_11aRxVector, CF_HTRxVector, CF_RxFrameBuffer, CF_CFOffset.
//
//   TB200Dot11aTx<T_CTX, T_NEXT>   source brick replacing both transmit graphs of kernel/bb/demod11/fb11amod_config.hpp:75-158
//   TB200Dot11bTx<T_CTX, T_NEXT>   source brick replacing the transmit graph of kernel/bb/demod11/fb11bmod_config.hpp:19-45
//   TB200Dot11nTx<T_CTX, T_NEXT>   source brick replacing the four transmit graphs of kernel/bb/demod11/fb11nmod_config.hpp:74-171
//       (TTS11aSrc | TBB11aSrc, T11aSc, TBB11aMRSelect, TConvEncode_*, T11aInterleave*, TMap11a*, T11aAddPilot, TIFFTx, TPackSample16to8):
//     one Process() modulates the MPDU in CF_TxFrameBuffer at CF_11aTxVector::data_rate_kbps with CF_ScramblerSeed::sc_seed and pushes
//     the whole PPDU (preamble + SIGNAL + DATA) downstream as COMPLEX8 x 8 bursts, what TPackSample16to8 hands to TModSink.
//
// Batching: the GPU decodes windows of samples.  The 802.11a receive brick buffers incoming 28-sample blocks and decodes a window in
// continuous-capture mode when `SetSlotSamples(n)` new samples have arrived or on Flush(); every frame of the window becomes one event, handed
// to the driver one per poll, and the samples behind the last event stay for the next window.  All bricks of a process share one engine
// (B200Engine); B200StreamBatcher decodes the windows of several graph instances (threads) in one device call.  A throughput-oriented caller
// that owns the capture uses sb200_rx11a_batch / sb200_rx11a_streams directly with thousands of slots per call.
#pragma once
#include "facades.hpp"
#include "../../include/sora_b200.h"
#include <vector>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <cstring>
#include <cstdlib>
#ifndef SB200_BRICK_DEVICE
#define SB200_BRICK_DEVICE 0      /* CUDA device ordinal the bricks of this translation unit bind to */
#endif

// ---- one engine per process and device; windows of many graph instances decoded together --------------------------------------------------
// sb200 handles are thread-compatible, not thread-safe: the shared handle is used under its mutex.  B200StreamBatcher::Decode is what the
// 802.11a receive brick calls with a window of samples: when `participants` graph instances (threads) are configured, the calls of one round are
// collected and go to the device as ONE sb200_rx11a_streams call (every pass of it decodes the next frame of all windows at once), so K radios
// cost K frames per kernel pass instead of one.  A lone caller, or a round that stays incomplete for `wait_us`, is decoded on its own.
struct B200Engine {
    sb200_handle* h = nullptr; std::mutex m;
    static B200Engine& Get(uint32_t cca_pwr_threshold = 0) {
        static B200Engine e;
        std::lock_guard<std::mutex> l(e.m);
        if (!e.h) { sb200_cfg cfg; memset(&cfg, 0, sizeof cfg); cfg.cca_pwr_threshold = cca_pwr_threshold; if (sb200_create(SB200_BRICK_DEVICE, &cfg, &e.h) != SB200_OK) e.h = nullptr; }
        return e;
    }
    ~B200Engine() { sb200_destroy(h); }
};
struct B200Event { sb200_frame_result r; uint32_t end_sample; std::vector<uchar> bytes; };
class B200StreamBatcher {
    // Windows are staged in ONE page-locked arena (sb200_host_alloc): every participant copies its own window into a region it reserved under
    // the lock — K memcpys in K threads, in parallel, while the round is still filling — and the closing thread hands the arena to
    // sb200_rx11a_streams, which sends only the regions in use across PCIe by DMA.  A request that does not fit (the arena grows only between
    // rounds) is staged by the closing thread.
    struct Req { const COMPLEX16* p; size_t n; uint32_t max_events; std::vector<B200Event>* out; int rc; bool done, taken, staged; size_t off; };
    std::mutex m_; std::condition_variable cv_; std::vector<Req*> pend_; unsigned participants_ = 1; unsigned wait_us_ = 2000; bool running_ = false;
    int16_t* arena_ = nullptr; size_t cap_ = 0, used_ = 0, want_ = 0; unsigned copying_ = 0; bool pinned_ = false;   // cap_, used_, want_ in samples
    std::vector<uint64_t> off_; std::vector<uint32_t> len_, sidx_, cnt_; std::vector<sb200_frame_result> res_; std::vector<uchar> bytes_;
    static size_t Round8(size_t n) { return (n + 7) & ~(size_t)7; }
    void Grow(size_t samples) {                             // only with no copy in flight and no region in use
        if (arena_) { if (pinned_) sb200_host_free(arena_); else free(arena_); }
        arena_ = (int16_t*)sb200_host_alloc(samples * 4 + 64); pinned_ = arena_ != nullptr;
        if (!arena_) arena_ = (int16_t*)malloc(samples * 4 + 64);                  // no device: the decode call will fail loudly
        cap_ = arena_ ? samples : 0;
    }
    void Run(std::vector<Req*>& batch) {                  // called without the lock by the thread that closes the round (no copy is in flight any more)
        B200Engine& E = B200Engine::Get(); std::lock_guard<std::mutex> le(E.m);
        uint32_t K = 1; size_t tot = 0; bool all_staged = true;
        for (Req* q : batch) { if (q->max_events > K) K = q->max_events; tot += Round8(q->n); all_staged = all_staged && q->staged; }
        const uint32_t S = (uint32_t)batch.size(); const uint32_t row = 2560;
        if (!all_staged) {                                  // lay everything out again from the callers' own buffers (they are blocked in Decode)
            if (tot > cap_) Grow(tot + tot / 4);
            size_t o = 0;
            for (Req* q : batch) { if (arena_) memcpy(arena_ + 2 * o, q->p, q->n * sizeof(COMPLEX16)); q->off = o; o += Round8(q->n); }
        }
        if (tot > want_) want_ = tot;
        off_.resize(S); len_.resize(S); cnt_.assign(S, 0); sidx_.assign((size_t)S * K, 0); res_.resize((size_t)S * K); bytes_.resize((size_t)S * K * row);
        size_t span = 0;
        for (uint32_t i = 0; i < S; i++) { off_[i] = batch[i]->off; len_[i] = (uint32_t)batch[i]->n; if (batch[i]->off + batch[i]->n > span) span = batch[i]->off + batch[i]->n; }
        int rc = (E.h && arena_) ? sb200_rx11a_streams(E.h, arena_, span, off_.data(), len_.data(), S, K, bytes_.data(), row, res_.data(), sidx_.data(), cnt_.data(), nullptr) : SB200_E_NODEVICE;
        for (uint32_t i = 0; i < S; i++) {
            Req* q = batch[i]; q->rc = rc; q->out->clear();
            if (rc == SB200_OK) for (uint32_t k = 0; k < cnt_[i] && k < q->max_events; k++) {
                B200Event e; e.r = res_[(size_t)i * K + k]; e.end_sample = sidx_[(size_t)i * K + k];
                const uchar* b = bytes_.data() + ((size_t)i * K + k) * row; e.bytes.assign(b, b + (e.r.length < row ? e.r.length : row));
                q->out->push_back(std::move(e));
            }
        }
    }
public:
    static B200StreamBatcher& Get() { static B200StreamBatcher b; return b; }
    ~B200StreamBatcher() { if (arena_) { if (pinned_) sb200_host_free(arena_); else free(arena_); } }
    void Configure(unsigned participants, unsigned wait_us = 2000) { std::lock_guard<std::mutex> l(m_); participants_ = participants ? participants : 1; wait_us_ = wait_us; }
    int Decode(const COMPLEX16* p, size_t n, uint32_t max_events, std::vector<B200Event>& out) {
        Req q{p, n, max_events, &out, SB200_OK, false, false, false, 0};
        std::unique_lock<std::mutex> l(m_);
        if (!running_ && pend_.empty() && copying_ == 0) {   // a round opens: the arena is free, size it for what the last rounds needed
            used_ = 0;
            const size_t need = want_ + want_ / 4;
            if (need > cap_) Grow(need);
        }
        pend_.push_back(&q);
        if (!running_ && arena_ && used_ + Round8(n) <= cap_) {                    // stage my own window while the round fills
            q.off = used_; used_ += Round8(n); copying_++;
            l.unlock(); memcpy(arena_ + 2 * q.off, p, n * sizeof(COMPLEX16)); l.lock();
            q.staged = true; copying_--; cv_.notify_all();
        }
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(wait_us_);
        while (!q.done) {
            const bool late = std::chrono::steady_clock::now() >= deadline;
            if (!running_ && !q.taken && copying_ == 0 && (pend_.size() >= participants_ || late)) {   // this thread closes the round (complete, or waited long enough) and runs it
                std::vector<Req*> batch; batch.swap(pend_);
                for (Req* r : batch) r->taken = true;
                running_ = true;
                l.unlock(); Run(batch); l.lock();
                for (Req* r : batch) r->done = true;
                running_ = false; cv_.notify_all();
            } else if ((late && copying_ == 0) || q.taken) cv_.wait(l);
            else cv_.wait_until(l, deadline);
        }
        return q.rc;
    }
};

DEFINE_LOCAL_CONTEXT(TB200Dot11aRx, CF_Error, CF_11aRxVector, CF_RxFrameBuffer, CF_11CCA, CF_CFOffset);
template <TFILTER_ARGS>
class TB200Dot11aRx : public TFilter<TFILTER_PARAMS> {
    CTX_VAR_RW(ulong, error_code)
    CTX_VAR_RW(ushort, frame_length) CTX_VAR_RW(ushort, total_symbols) CTX_VAR_RW(ulong, data_rate_kbps) CTX_VAR_RW(ushort, code_rate) CTX_VAR_RW(ulong, frame_crc32)
    CTX_VAR_RO(uchar*, rx_frame_buf) CTX_VAR_RO(uint, rx_frame_buf_size)
    CTX_VAR_RO(uint, cca_pwr_threshold) CTX_VAR_RW(uint, cca_peak_index) CTX_VAR_RW(CF_11CCA::CCAState, cca_state)
    CTX_VAR_RW(short, CFO_est)
    bool ok_;
    std::vector<COMPLEX16> buf_;                          // samples received and not yet behind a delivered event
    std::vector<B200Event> ev_; size_t ev_next_;          // events of the last decoded window, delivered one per Process() / Flush()
    size_t window_, decoded_upto_; uint32_t max_events_; bool truncated_;   // truncated_: the last window filled its event list, more may follow
public:
    DEFINE_IPORT(COMPLEX16, 28);
    DEFINE_OPORT(uchar, 1);
    REFERENCE_LOCAL_CONTEXT(TB200Dot11aRx);
    STD_TFILTER_CONSTRUCTOR(TB200Dot11aRx)
        BIND_CONTEXT(CF_Error::error_code, error_code)
        BIND_CONTEXT(CF_11aRxVector::frame_length, frame_length) BIND_CONTEXT(CF_11aRxVector::total_symbols, total_symbols)
        BIND_CONTEXT(CF_11aRxVector::data_rate_kbps, data_rate_kbps) BIND_CONTEXT(CF_11aRxVector::code_rate, code_rate)
        BIND_CONTEXT(CF_11aRxVector::crc32, frame_crc32)
        BIND_CONTEXT(CF_RxFrameBuffer::rx_frame_buf, rx_frame_buf) BIND_CONTEXT(CF_RxFrameBuffer::rx_frame_buf_size, rx_frame_buf_size)
        BIND_CONTEXT(CF_11CCA::cca_pwr_threshold, cca_pwr_threshold) BIND_CONTEXT(CF_11CCA::cca_peak_index, cca_peak_index)
        BIND_CONTEXT(CF_11CCA::cca_state, cca_state) BIND_CONTEXT(CF_CFOffset::CFO_est, CFO_est)
        , ok_(false), ev_next_(0), window_(0), decoded_upto_(0), max_events_(256), truncated_(false)
    {
        ok_ = B200Engine::Get(cca_pwr_threshold).h != nullptr;                                   // no CPU fallback
        if (!ok_) error_code = E_ERROR_FAILED;
    }
    // decode whenever this many new 40 Msps samples have arrived (0 = only on Flush(): a whole dump file is one window, like demod11 -d);
    // every window goes through continuous-capture mode, so all of its frames come out, one event per driver poll
    void SetSlotSamples(size_t n) { window_ = n; }
    void SetMaxEventsPerWindow(uint32_t k) { max_events_ = k ? k : 1; }

    // The driver's Flush(); ctx.Reset(); Reset() after an event (fb11a_demod.cpp:64-72) must not lose what the source already pumped in:
    // samples behind the delivered event and the events still queued stay.
    STD_TFILTER_RESET() { }
    STD_TFILTER_FLUSH() { if (error_code == E_ERROR_SUCCESS) { if (!Deliver() ) return; if (buf_.size() > decoded_upto_ || ev_next_ < ev_.size()) { DecodeWindow(); Deliver(); } } }

    BOOL_FUNC_PROCESS(ipin) {
        while (ipin.check_read()) {
            const COMPLEX16* p = ipin.peek();
            buf_.insert(buf_.end(), p, p + 28);
            ipin.pop();
            if (ev_next_ < ev_.size()) { if (!Deliver()) return false; }
            else if ((window_ && buf_.size() - decoded_upto_ >= window_) || (truncated_ && decoded_upto_ == 0 && buf_.size() >= 28)) { if (!DecodeWindow()) return false; if (!Deliver()) return false; }
        }
        return true;
    }
    bool EventsPending() const { return ev_next_ < ev_.size(); }
private:
    bool DecodeWindow() {                                  // all events of buf_ (a frame cut by its end is not an event: it waits for more samples)
        if (!ok_) { error_code = E_ERROR_FAILED; return false; }
        ev_.clear(); ev_next_ = 0; decoded_upto_ = buf_.size();
        if (buf_.size() < 28) return true;
        if (B200StreamBatcher::Get().Decode(buf_.data(), buf_.size(), max_events_, ev_) != SB200_OK) { error_code = E_ERROR_FAILED; return false; }
        truncated_ = ev_.size() >= max_events_;
        const size_t keep = 5880 * 28;                     // longer than the longest PPDU (2500 B at 6 Mbps = 3.4 ms = 136 000 samples at 40 Msps)
        if (ev_.empty() && buf_.size() > keep) {           // nothing in this window: a frame still arriving can only be inside its last `keep` samples
            const size_t cut = (buf_.size() - keep) / 28 * 28;
            buf_.erase(buf_.begin(), buf_.begin() + cut); decoded_upto_ -= cut;
        }
        return true;
    }
    bool Deliver() {                                       // next queued event -> context + output port; false = an event was delivered (stop pumping, PHY_11a.hpp:694)
        if (ev_next_ >= ev_.size()) return true;
        const B200Event& e = ev_[ev_next_++]; const sb200_frame_result& r = e.r;
        frame_length = (ushort)r.length; total_symbols = (ushort)r.nsym; data_rate_kbps = r.rate_kbps; frame_crc32 = r.crc32;
        code_rate = (ushort)(r.rate_kbps == 48000 ? CR_23 : (r.rate_kbps == 9000 || r.rate_kbps == 18000 || r.rate_kbps == 36000 || r.rate_kbps == 54000) ? CR_34 : CR_12);
        cca_peak_index = r.peak_index; cca_state = CF_11CCA::power_detected; CFO_est = r.cfo_est;
        if (r.status == SB200_FRAME_OK || r.status == SB200_FRAME_CRC32_FAIL) {
            const uint n = (uint)e.bytes.size();
            if (rx_frame_buf && n <= rx_frame_buf_size) memcpy(rx_frame_buf, e.bytes.data(), n);
            for (uint i = 0; i < n; i++) { *opin().append() = e.bytes[i]; this->Next()->Process(opin()); }
        }
        if (ev_next_ == ev_.size()) {                      // window exhausted: drop the samples up to the end of its last event
            const size_t cut = e.end_sample < buf_.size() ? e.end_sample : buf_.size();
            buf_.erase(buf_.begin(), buf_.begin() + cut); decoded_upto_ = truncated_ ? 0 : (decoded_upto_ > cut ? decoded_upto_ - cut : 0);   // a full event list: the rest is not scanned yet
            ev_.clear(); ev_next_ = 0;
        }
        error_code = r.status;                             // the driver polls this after Process() (fb11a_demod.cpp:35)
        return false;
    }
};


DEFINE_LOCAL_CONTEXT(TB200Dot11bRx, CF_Error, CF_11bRxVector, CF_RxFrameBuffer);
template <TFILTER_ARGS>
class TB200Dot11bRx : public TFilter<TFILTER_PARAMS> {
    CTX_VAR_RW(ulong, error_code)
    CTX_VAR_RW(ushort, frame_length) CTX_VAR_RW(ulong, data_rate_kbps) CTX_VAR_RW(ulong, frame_crc32)
    CTX_VAR_RO(uchar*, rx_frame_buf) CTX_VAR_RO(uint, rx_frame_buf_size)
    sb200_handle* h_; std::vector<COMPLEX16> slot_; std::vector<uchar> bytes_; size_t slot_samples_;
    std::vector<sb200_frame_result_11b> ev_; size_t ev_next_ = 0;   // events of the last decoded window (continuous-capture mode), one per driver poll
public:
    DEFINE_IPORT(COMPLEX16, 28);
    DEFINE_OPORT(uchar, 1);
    REFERENCE_LOCAL_CONTEXT(TB200Dot11bRx);
    STD_TFILTER_CONSTRUCTOR(TB200Dot11bRx)
        BIND_CONTEXT(CF_Error::error_code, error_code)
        BIND_CONTEXT(CF_11bRxVector::frame_length, frame_length) BIND_CONTEXT(CF_11bRxVector::data_rate_kbps, data_rate_kbps) BIND_CONTEXT(CF_11bRxVector::crc32, frame_crc32)
        BIND_CONTEXT(CF_RxFrameBuffer::rx_frame_buf, rx_frame_buf) BIND_CONTEXT(CF_RxFrameBuffer::rx_frame_buf_size, rx_frame_buf_size)
        , h_(nullptr), slot_samples_(0)
    {
        h_ = B200Engine::Get().h; if (!h_) error_code = E_ERROR_FAILED;                 // one engine per process (no CPU fallback)
        bytes_.resize(16 * 4096);
    }
    void SetSlotSamples(size_t n) { slot_samples_ = n; }
    STD_TFILTER_RESET() { }                               // samples behind a delivered event and queued events survive the driver's Reset()
    STD_TFILTER_FLUSH() { if (error_code == E_ERROR_SUCCESS) Submit(); }
    bool EventsPending() const { return ev_next_ < ev_.size(); }
    BOOL_FUNC_PROCESS(ipin) {
        while (ipin.check_read()) {
            const COMPLEX16* p = ipin.peek(); slot_.insert(slot_.end(), p, p + 28); ipin.pop();
            if (ev_next_ < ev_.size() || (slot_samples_ && slot_.size() >= slot_samples_)) { if (!Submit()) return false; }
        }
        return true;
    }
private:
    bool Submit() {
        if (!h_) { error_code = E_ERROR_FAILED; return false; }
        if (ev_next_ >= ev_.size()) {                      // decode the window: every event MAC11b_Receive would meet in it (fb11b_demod.cpp:26-75)
            if (slot_.empty()) return true;
            const uint32_t K = 16; uint64_t off = 0; uint32_t len = (uint32_t)slot_.size(), n = 0;
            ev_.assign(K, sb200_frame_result_11b()); ev_next_ = 0;
            int rc;
            { std::lock_guard<std::mutex> engine_lock(B200Engine::Get().m);
              rc = sb200_rx11b_streams(h_, (const int16_t*)slot_.data(), slot_.size(), &off, &len, 1, K, bytes_.data(), 4096, ev_.data(), &n, nullptr); }
            if (rc != SB200_OK) { ev_.clear(); error_code = E_ERROR_FAILED; return false; }
            ev_.resize(n);
            if (n == 0) { slot_.clear(); return true; }            // nothing detected in this window: keep sensing
        }
        const size_t k = ev_next_++; const sb200_frame_result_11b r = ev_[k];
        frame_length = (ushort)r.length; data_rate_kbps = r.rate_kbps; frame_crc32 = r.crc32;
        if (r.status == SB200_FRAME_OK || r.status == SB200_FRAME_CRC32_FAIL) {
            const uint n = r.length ? r.length - 1 : 0;            // TBB11bFrameSink decides on the third FCS byte (PHY_11b.hpp:728-739)
            const uchar* b = bytes_.data() + k * 4096;
            if (rx_frame_buf && n <= rx_frame_buf_size) memcpy(rx_frame_buf, b, n);
            for (uint i = 0; i < n; i++) { *opin().append() = b[i]; this->Next()->Process(opin()); }
        }
        if (ev_next_ == ev_.size()) {                      // window exhausted: what lies behind its last event waits for the next one
            size_t cut = r.sample_index;                   // + the driver's seek past the last FCS byte after a frame (fb11b_demod.cpp:43-61)
            if (r.status == SB200_FRAME_OK || r.status == SB200_FRAME_CRC32_FAIL) cut += r.rate_kbps == 1000 ? 352u : r.rate_kbps == 2000 ? 176u : r.rate_kbps == 5500 ? 64u : 32u;
            if (cut > slot_.size()) cut = slot_.size();
            slot_.erase(slot_.begin(), slot_.begin() + cut); ev_.clear(); ev_next_ = 0;
        }
        error_code = r.status;
        return false;
    }
};

DEFINE_LOCAL_CONTEXT(TB200Dot11nRx, CF_Error, CF_11aRxVector, CF_HTRxVector, CF_RxFrameBuffer, CF_CFOffset);
template <TFILTER_ARGS>
class TB200Dot11nRx : public TFilter<TFILTER_PARAMS> {
    CTX_VAR_RW(ulong, error_code)
    CTX_VAR_RW(ushort, frame_length) CTX_VAR_RW(ushort, total_symbols) CTX_VAR_RW(ushort, code_rate) CTX_VAR_RW(ulong, frame_crc32)
    CTX_VAR_RW(ushort, ht_frame_length) CTX_VAR_RW(ulong, ht_frame_mcs)
    CTX_VAR_RO(uchar*, rx_frame_buf) CTX_VAR_RO(uint, rx_frame_buf_size)
    CTX_VAR_RW(short, CFO_est)
    sb200_handle* h_; std::vector<COMPLEX16> slot_[2]; std::vector<uchar> bytes_; size_t slot_samples_;
    std::vector<sb200_frame_result_11n> ev_; std::vector<uint32_t> ev_end_; size_t ev_next_ = 0;   // events of the last decoded window, one per driver poll
public:
    static const size_t NSTREAM = 2;
    DEFINE_IPORT(COMPLEX16, 28, NSTREAM);
    DEFINE_OPORT(uchar, 1);
    REFERENCE_LOCAL_CONTEXT(TB200Dot11nRx);
    STD_TFILTER_CONSTRUCTOR(TB200Dot11nRx)
        BIND_CONTEXT(CF_Error::error_code, error_code)
        BIND_CONTEXT(CF_11aRxVector::frame_length, frame_length) BIND_CONTEXT(CF_11aRxVector::total_symbols, total_symbols)
        BIND_CONTEXT(CF_11aRxVector::code_rate, code_rate) BIND_CONTEXT(CF_11aRxVector::crc32, frame_crc32)
        BIND_CONTEXT(CF_HTRxVector::ht_frame_length, ht_frame_length) BIND_CONTEXT(CF_HTRxVector::ht_frame_mcs, ht_frame_mcs)
        BIND_CONTEXT(CF_RxFrameBuffer::rx_frame_buf, rx_frame_buf) BIND_CONTEXT(CF_RxFrameBuffer::rx_frame_buf_size, rx_frame_buf_size)
        BIND_CONTEXT(CF_CFOffset::CFO_est, CFO_est)
        , h_(nullptr), slot_samples_(0)
    {
        h_ = B200Engine::Get().h; if (!h_) error_code = E_ERROR_FAILED;                 // one engine per process (no CPU fallback)
        bytes_.resize(16 * 2048);
    }
    void SetSlotSamples(size_t n) { slot_samples_ = n; }
    STD_TFILTER_RESET() { }                               // samples behind a delivered event and queued events survive the driver's Reset()
    STD_TFILTER_FLUSH() { if (error_code == E_ERROR_SUCCESS) Submit(); }
    bool EventsPending() const { return ev_next_ < ev_.size(); }
    BOOL_FUNC_PROCESS(ipin) {
        while (ipin.check_read()) {
            for (size_t s = 0; s < NSTREAM; s++) { const COMPLEX16* p = ipin.peek(s); slot_[s].insert(slot_[s].end(), p, p + 28); }
            ipin.pop();
            if (ev_next_ < ev_.size() || (slot_samples_ && slot_[0].size() >= slot_samples_)) { if (!Submit()) return false; }
        }
        return true;
    }
private:
    bool Submit() {
        if (!h_) { error_code = E_ERROR_FAILED; return false; }
        if (ev_next_ >= ev_.size()) {                      // decode the window in continuous-capture mode (fb11n_demod.cpp:29-81)
            if (slot_[0].empty()) return true;
            const uint32_t K = 16; uint64_t off = 0; uint32_t len = (uint32_t)slot_[0].size(), n = 0;
            ev_.assign(K, sb200_frame_result_11n()); ev_end_.assign(K, 0); ev_next_ = 0;
            int rc;
            { std::lock_guard<std::mutex> engine_lock(B200Engine::Get().m);
              rc = sb200_rx11n_streams(h_, (const int16_t*)slot_[0].data(), (const int16_t*)slot_[1].data(), slot_[0].size(), &off, &len, 1, K,
                                       bytes_.data(), 2048, ev_.data(), ev_end_.data(), &n, nullptr); }
            if (rc != SB200_OK) { ev_.clear(); error_code = E_ERROR_FAILED; return false; }
            ev_.resize(n);
            if (n == 0) { slot_[0].clear(); slot_[1].clear(); return true; }
        }
        const size_t k = ev_next_++; const sb200_frame_result_11n r = ev_[k];
        frame_length = (ushort)r.length; total_symbols = (ushort)r.nsym; frame_crc32 = r.crc32; CFO_est = r.cfo_est;
        ht_frame_mcs = r.mcs; ht_frame_length = (ushort)(r.nsym ? r.length : 0);
        code_rate = (ushort)(r.mcs == 10 ? CR_34 : CR_12);
        if (r.status == SB200_FRAME_OK || r.status == SB200_FRAME_CRC32_FAIL) {
            const uint n = r.length; const uchar* b = bytes_.data() + k * 2048;
            if (rx_frame_buf && n <= rx_frame_buf_size) memcpy(rx_frame_buf, b, n);
            for (uint i = 0; i < n; i++) { *opin().append() = b[i]; this->Next()->Process(opin()); }
        }
        if (ev_next_ == ev_.size()) {
            const size_t cut = ev_end_[k] < slot_[0].size() ? ev_end_[k] : slot_[0].size();
            slot_[0].erase(slot_[0].begin(), slot_[0].begin() + cut); slot_[1].erase(slot_[1].begin(), slot_[1].begin() + cut); ev_.clear(); ev_next_ = 0;
        }
        error_code = r.status;
        return false;
    }
};


DEFINE_LOCAL_CONTEXT(TB200Dot11aTx, CF_Error, CF_11aTxVector, CF_TxFrameBuffer, CF_ScramblerSeed);
template <TSOURCE_ARGS>
class TB200Dot11aTx : public TSource<TSOURCE_PARAMS> {
    CTX_VAR_RW(ulong, error_code)
    CTX_VAR_RO(ushort, frame_length) CTX_VAR_RO(ulong, data_rate_kbps)
    CTX_VAR_RO(uchar*, mpdu_buf0) CTX_VAR_RO(ushort, mpdu_buf_size0) CTX_VAR_RO(uchar*, mpdu_buf1) CTX_VAR_RO(ushort, mpdu_buf_size1)
    CTX_VAR_RO(uchar, sc_seed)
    sb200_handle* h_; std::vector<uchar> mpdu_; std::vector<int8_t> td_;
public:
    DEFINE_OPORT(COMPLEX8, 8);
    REFERENCE_LOCAL_CONTEXT(TB200Dot11aTx);
    STD_TSOURCE_CONSTRUCTOR(TB200Dot11aTx)
        BIND_CONTEXT(CF_Error::error_code, error_code)
        BIND_CONTEXT(CF_11aTxVector::frame_length, frame_length) BIND_CONTEXT(CF_11aTxVector::data_rate_kbps, data_rate_kbps)
        BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf0, mpdu_buf0) BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf_size0, mpdu_buf_size0)
        BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf1, mpdu_buf1) BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf_size1, mpdu_buf_size1)
        BIND_CONTEXT(CF_ScramblerSeed::sc_seed, sc_seed)
        , h_(nullptr)
    {
        h_ = B200Engine::Get().h; if (!h_) error_code = E_ERROR_FAILED;                 // one engine per process (no CPU fallback)
    }
    STD_TSOURCE_RESET() { }
    STD_TSOURCE_FLUSH() { }
    bool Process() override {
        if (!h_) { error_code = E_ERROR_FAILED; return false; }
        std::lock_guard<std::mutex> engine_lock(B200Engine::Get().m);
        if (!mpdu_buf0 || (mpdu_buf_size1 > 0 && !mpdu_buf1) || (uint)mpdu_buf_size0 + mpdu_buf_size1 != frame_length) { error_code = E_ERROR_PARAMETER; return false; }   // TBB11aSrc::Preprocess
        mpdu_.assign(mpdu_buf0, mpdu_buf0 + mpdu_buf_size0); if (mpdu_buf_size1) mpdu_.insert(mpdu_.end(), mpdu_buf1, mpdu_buf1 + mpdu_buf_size1);
        const uint64_t off = 0; const uint32_t len = frame_length; uint32_t ns = 0; const uchar seed = sc_seed;
        const size_t cap = 640 + 160 * (size_t)(3 + (len + 7) * 8 / 24) + 64;              // enough for the slowest rate
        td_.assign(2 * cap, 0);
        if (sb200_tx11a_batch(h_, mpdu_.empty() ? (const uint8_t*)"" : mpdu_.data(), mpdu_.size() ? mpdu_.size() : 1, &off, &len, &seed, 1, (uint32_t)data_rate_kbps, 0, 8,
                              td_.data(), cap, &ns, nullptr) != SB200_OK) { error_code = E_ERROR_PARAMETER; return false; }
        for (uint32_t i = 0; i + 8 <= ns; i += 8) { memcpy(opin().append(), td_.data() + 2 * i, 16); this->Next()->Process(opin()); }
        return false;                                              // one PPDU per Process(), like TBB11aSrc / TTS11aSrc
    }
};


// 802.11b transmit: everything between TBB11bSrc and TModSink of fb11bmod_config.hpp:25-44 as one source brick.  Same context as the
// graph it replaces (CF_11bTxVector, CF_TxFrameBuffer, CF_DifferentialMap, CF_Error); emits COMPLEX8 bursts of 8 like TPackSample16to8.
DEFINE_LOCAL_CONTEXT(TB200Dot11bTx, CF_Error, CF_11bTxVector, CF_TxFrameBuffer, CF_DifferentialMap);
template <TSOURCE_ARGS>
class TB200Dot11bTx : public TSource<TSOURCE_PARAMS> {
    CTX_VAR_RW(ulong, error_code)
    CTX_VAR_RO(ushort, frame_length) CTX_VAR_RO(uchar, preamble_type) CTX_VAR_RO(ulong, data_rate_kbps)
    CTX_VAR_RO(uchar*, mpdu_buf0) CTX_VAR_RO(ushort, mpdu_buf_size0) CTX_VAR_RO(uchar*, mpdu_buf1) CTX_VAR_RO(ushort, mpdu_buf_size1)
    CTX_VAR_RW(uint, last_phase)
    sb200_handle* h_; std::vector<uchar> mpdu_; std::vector<int8_t> td_;
public:
    DEFINE_OPORT(COMPLEX8, 8);
    REFERENCE_LOCAL_CONTEXT(TB200Dot11bTx);
    STD_TSOURCE_CONSTRUCTOR(TB200Dot11bTx)
        BIND_CONTEXT(CF_Error::error_code, error_code)
        BIND_CONTEXT(CF_11bTxVector::frame_length, frame_length) BIND_CONTEXT(CF_11bTxVector::preamble_type, preamble_type)
        BIND_CONTEXT(CF_11bTxVector::data_rate_kbps, data_rate_kbps)
        BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf0, mpdu_buf0) BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf_size0, mpdu_buf_size0)
        BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf1, mpdu_buf1) BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf_size1, mpdu_buf_size1)
        BIND_CONTEXT(CF_DifferentialMap::last_phase, last_phase)
        , h_(nullptr)
    {
        h_ = B200Engine::Get().h; if (!h_) error_code = E_ERROR_FAILED;                 // one engine per process (no CPU fallback)
    }
    STD_TSOURCE_RESET() { }
    STD_TSOURCE_FLUSH() { }
    bool Process() override {
        if (!h_) { error_code = E_ERROR_FAILED; return false; }
        std::lock_guard<std::mutex> engine_lock(B200Engine::Get().m);
        error_code = E_ERROR_SUCCESS;
        if (!mpdu_buf0 || (mpdu_buf_size1 > 0 && !mpdu_buf1) || (uint)mpdu_buf_size0 + mpdu_buf_size1 != frame_length) { error_code = E_ERROR_PARAMETER; return false; }   // TBB11bSrc::Preprocess
        const ulong rate = data_rate_kbps;
        if (rate != 1000 && rate != 2000 && rate != 5500 && rate != 11000) { error_code = E_ERROR_DATARATE; return false; }     // PHY_11b.hpp:89-94
        if (preamble_type == 1) { error_code = E_ERROR_NOT_SUPPORTED; return false; }                                             // PHY_11b.hpp:96-100
        mpdu_.assign(mpdu_buf0, mpdu_buf0 + mpdu_buf_size0); if (mpdu_buf_size1) mpdu_.insert(mpdu_.end(), mpdu_buf1, mpdu_buf1 + mpdu_buf_size1);
        const uint64_t off = 0; const uint32_t len = frame_length; uint32_t ns = 0, fin = 0;
        const size_t cap = ((24 * 88 + ((size_t)len + 4) * 88 + 5) * 4 + 15) / 8 * 8;          // enough for 1 Mbps
        td_.assign(2 * cap + 16, 0);
        int8_t* td = td_.data() + ((16 - ((uintptr_t)td_.data() & 15)) & 15);                  // the entry point wants 16-byte alignment
        if (sb200_tx11b_batch(h_, mpdu_.empty() ? (const uint8_t*)"" : mpdu_.data(), mpdu_.size() ? mpdu_.size() : 1, &off, &len, 1, (uint32_t)rate, last_phase & 3u, 0, 8,
                              td, cap, &ns, &fin, nullptr) != SB200_OK) { error_code = E_ERROR_PARAMETER; return false; }
        for (uint32_t i = 0; i + 8 <= ns; i += 8) { memcpy(opin().append(), td + 2 * i, 16); this->Next()->Process(opin()); }
        last_phase = fin;                                          // the reference never resets it between frames (barkerspread.hpp:99-100, cck.hpp:864)
        return false;
    }
};


// 802.11n two-stream transmit: the preamble, SIG and DATA graphs of fb11nmod_config.hpp as one source brick with a two-stream output port
// (vectors of 4 COMPLEX16 per stream, what the two TModSink1 sinks of the reference take).  Context: CF_11nTxVector, CF_TxFrameBuffer,
// CF_ScramblerSeed, CF_Error — the fields BB11nModCtx::init fills (fb11nmod_config.hpp:31-53).
DEFINE_LOCAL_CONTEXT(TB200Dot11nTx, CF_Error, CF_11nTxVector, CF_TxFrameBuffer, CF_ScramblerSeed);
template <TSOURCE_ARGS>
class TB200Dot11nTx : public TSource<TSOURCE_PARAMS> {
    CTX_VAR_RW(ulong, error_code)
    CTX_VAR_RO(ushort, frame_length) CTX_VAR_RO(ushort, mcs_index)
    CTX_VAR_RO(uchar*, mpdu_buf0) CTX_VAR_RO(ushort, mpdu_buf_size0) CTX_VAR_RO(uchar*, mpdu_buf1) CTX_VAR_RO(ushort, mpdu_buf_size1)
    CTX_VAR_RO(uchar, sc_seed)
    sb200_handle* h_; std::vector<uchar> mpdu_; std::vector<int16_t> td_[2];
public:
    static const size_t NSTREAM = 2;
    DEFINE_OPORT(COMPLEX16, 4, NSTREAM);
    REFERENCE_LOCAL_CONTEXT(TB200Dot11nTx);
    STD_TSOURCE_CONSTRUCTOR(TB200Dot11nTx)
        BIND_CONTEXT(CF_Error::error_code, error_code)
        BIND_CONTEXT(CF_11nTxVector::frame_length, frame_length) BIND_CONTEXT(CF_11nTxVector::mcs_index, mcs_index)
        BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf0, mpdu_buf0) BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf_size0, mpdu_buf_size0)
        BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf1, mpdu_buf1) BIND_CONTEXT(CF_TxFrameBuffer::mpdu_buf_size1, mpdu_buf_size1)
        BIND_CONTEXT(CF_ScramblerSeed::sc_seed, sc_seed)
        , h_(nullptr)
    {
        h_ = B200Engine::Get().h; if (!h_) error_code = E_ERROR_FAILED;                 // one engine per process (no CPU fallback)
    }
    STD_TSOURCE_RESET() { }
    STD_TSOURCE_FLUSH() { }
    bool Process() override {
        if (!h_) { error_code = E_ERROR_FAILED; return false; }
        std::lock_guard<std::mutex> engine_lock(B200Engine::Get().m);
        error_code = E_ERROR_SUCCESS;
        if (!mpdu_buf0 || (mpdu_buf_size1 > 0 && !mpdu_buf1) || (uint)mpdu_buf_size0 + mpdu_buf_size1 != frame_length) { error_code = E_ERROR_PARAMETER; return false; }   // TBB11nSrc::Preprocess
        mpdu_.assign(mpdu_buf0, mpdu_buf0 + mpdu_buf_size0); if (mpdu_buf_size1) mpdu_.insert(mpdu_.end(), mpdu_buf1, mpdu_buf1 + mpdu_buf_size1);
        const uint64_t off = 0; const uint32_t len = frame_length; uint32_t ns = 0; const uchar seed = sc_seed;
        const size_t cap = 1600 + 160 * (size_t)(2 + ((len + 4) * 8 + 22) / 52);            // enough for MCS 8
        td_[0].assign(2 * cap, 0); td_[1].assign(2 * cap, 0);
        if (sb200_tx11n_batch(h_, mpdu_.empty() ? (const uint8_t*)"" : mpdu_.data(), mpdu_.size() ? mpdu_.size() : 1, &off, &len, &seed, 1, (uint32_t)mcs_index, 0,
                              td_[0].data(), td_[1].data(), cap, &ns, nullptr) != SB200_OK) { error_code = E_ERROR_PARAMETER; return false; }
        for (uint32_t i = 0; i + 4 <= ns; i += 4) {
            for (size_t s = 0; s < NSTREAM; s++) memcpy(opin().write(s), td_[s].data() + 2 * i, 16);
            opin().append(); this->Next()->Process(opin());
        }
        return false;                                              // one PPDU per Process(), like the four graphs run back to back (fb11n_mod.cpp:44-70)
    }
};
