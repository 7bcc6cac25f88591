// Builds a brick graph the way kernel/bb/demod11/fb11ademod_config.hpp does, with the CPU sub-chain ds2..fsink replaced
// by the GPU brick, and drives it like RxThread (kernel/bb/demod11/fb11a_demod.cpp:29-81) over a Sora dump file.
//   usage: demo_graph <file.dmp> [legacy14] [--threads K] [--window N] [--repeat R] [--max-events E]   (see main)
#include "b200_bricks.hpp"
#include <cstdio>
#include <string>
#include <vector>
#include <cstdlib>

// --- stock bricks of the reference, re-stated against brick.hpp: TMemSamples (memsource.hpp:17-152), TDropAny (stdbrick.hpp:34-49)
DEFINE_LOCAL_CONTEXT(TMemSamples, CF_MemSamples, CF_Error);
template <TSOURCE_ARGS> class TMemSamples : public TSource<TSOURCE_PARAMS> {
    CTX_VAR_RO(COMPLEX16*, sample_buf) CTX_VAR_RO(uint, sample_count) CTX_VAR_RW(uint, sample_index) CTX_VAR_RW(ulong, error_code)
    uint remain_; COMPLEX16* ptr_;
public:
    DEFINE_OPORT(COMPLEX16, 28);
    REFERENCE_LOCAL_CONTEXT(TMemSamples);
    STD_TSOURCE_CONSTRUCTOR(TMemSamples)
        BIND_CONTEXT(CF_MemSamples::mem_sample_buf, sample_buf) BIND_CONTEXT(CF_MemSamples::mem_sample_count, sample_count)
        BIND_CONTEXT(CF_MemSamples::mem_sample_index, sample_index) BIND_CONTEXT(CF_Error::error_code, error_code)
    { sample_index = 0; remain_ = sample_count; ptr_ = sample_buf; if (!sample_buf) error_code = BK_ERROR_FAILED; }
    STD_TSOURCE_RESET() {}
    STD_TSOURCE_FLUSH() {}
    bool Process() override {
        if (remain_ < 28) return false;                      // whole blocks only ("partial last block")
        memcpy(opin().append(), ptr_, 28 * sizeof(COMPLEX16));
        ptr_ += 28; sample_index += 28; remain_ -= 28;
        return this->Next()->Process(opin());
    }
};
DEFINE_LOCAL_CONTEXT(TByteCounter, CF_VOID);
template <TSINK_ARGS> class TByteCounter : public TSink<TSINK_PARAMS> {
public:
    unsigned long long nbytes = 0;
    DEFINE_IPORT(uchar, 1);
    STD_TSINK_CONSTRUCTOR(TByteCounter) {}
    BOOL_FUNC_PROCESS(pin) { while (pin.check_read()) { nbytes++; pin.pop(); } return true; }
};

struct DemodCtx : LOCAL_CONTEXT(TB200Dot11aRx), LOCAL_CONTEXT(TMemSamples), LOCAL_CONTEXT(TByteCounter) {
    void Reset() { CF_Error::error_code() = E_ERROR_SUCCESS; CF_11CCA::Reset(); CF_11aRxVector::Reset(); CF_CFOffset::Reset(); }
};

// one graph instance driven like RxThread over `iq`; returns the number of frame events (and FRAME_OK among them)
static unsigned g_max_events = 0;                             // --max-events N: events one decoded window may hold (0 = the adaptor's default)
static int run_graph(std::vector<COMPLEX16>& iq, size_t window, bool print, int* ok_frames) {
    DemodCtx* ctx = new DemodCtx(); static thread_local uchar frame[4096];
    ctx->CF_MemSamples::Init(iq.data(), (uint)(iq.size() * sizeof(COMPLEX16)));
    ctx->CF_RxFrameBuffer::Init(frame, sizeof frame);
    ctx->Reset();
    CREATE_BRICK_SINK(fsink, TByteCounter, *ctx);
    CREATE_BRICK_FILTER(gpurx, TB200Dot11aRx, *ctx, fsink);
    CREATE_BRICK_SOURCE(fsrc, TMemSamples, *ctx, gpurx);
    gpurx->SetSlotSamples(window);
    if (g_max_events) gpurx->SetMaxEventsPerWindow(g_max_events);
    ISource* ssrc = fsrc;
    int nframes = 0, nok = 0;
    for (;;) {                                               // RxThread
        bool more = ssrc->Process();
        if (!more && ctx->CF_Error::error_code() == E_ERROR_SUCCESS) ssrc->Flush();   // end of file: decode what is still buffered
        ulong err = ctx->CF_Error::error_code();
        if (err != E_ERROR_SUCCESS) {
            if (print) printf("{\"event\": %d, \"error_code\": \"0x%08X\", \"rate_kbps\": %u, \"length\": %u, \"crc32\": \"0x%08X\", \"bytes_out\": %llu}\n",
                   nframes, (unsigned)err, (unsigned)ctx->CF_11aRxVector::data_rate_kbps(), (unsigned)ctx->CF_11aRxVector::frame_length(),
                   (unsigned)ctx->CF_11aRxVector::crc32(), fsink->nbytes);
            nframes++; if (err == E_ERROR_FRAME_OK) nok++;
            ssrc->Flush(); ctx->Reset(); ssrc->Reset();
            if (!more) continue;                             // the source is dry but the brick may still hold events of its last window
        }
        if (!more) break;
    }
    IReferenceCounting::Release(ssrc);
    delete ctx;
    if (ok_frames) *ok_frames = nok;
    return nframes;
}

//   usage: demo_graph <file.dmp> [legacy14] [--threads K] [--window N] [--repeat R]
//     --threads K   K graph instances in K threads over the same capture (K radios): their windows are decoded together (B200StreamBatcher)
//     --window N    decode whenever N new samples have arrived (default 0: the whole file is one window)
//     prints one JSON line per frame event (one instance) or one summary line (K > 1 or --repeat)
#include <thread>
#include <chrono>
int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.dmp [legacy14] [--threads K] [--window N] [--repeat R]\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb"); if (!f) { perror("open"); return 2; }
    std::vector<COMPLEX16> iq; unsigned char blk[128];
    bool legacy = false; int K = 1, R = 1; size_t window = 0;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (a == "legacy14") legacy = true;
        else if (a == "--threads" && i + 1 < argc) K = atoi(argv[++i]);
        else if (a == "--window" && i + 1 < argc) window = (size_t)atoll(argv[++i]);
        else if (a == "--repeat" && i + 1 < argc) R = atoi(argv[++i]);
        else if (a == "--max-events" && i + 1 < argc) g_max_events = (unsigned)atoi(argv[++i]);
    }
    while (fread(blk, 1, 128, f) == 128) {                  // LoadSoraDumpFile: strip the 16-byte RX_BLOCK descriptor (brickutil.h:21-59)
        const COMPLEX16* s = (const COMPLEX16*)(blk + 16);
        for (int i = 0; i < 28; i++) { COMPLEX16 c = s[i]; if (legacy) { c.re = (short)(c.re << 2); c.im = (short)(c.im << 2); } iq.push_back(c); }
    }
    fclose(f);
    if (K <= 1 && R <= 1) return run_graph(iq, window, true, nullptr) ? 0 : 1;
    B200StreamBatcher::Get().Configure((unsigned)K);
    std::vector<int> nf(K, 0), nok(K, 0);
    run_graph(iq, window, false, nullptr);                   // warm-up: engine creation, table upload, workspace allocation
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < R; r++) {
        std::vector<std::thread> th;
        for (int k = 0; k < K; k++) th.emplace_back([&, k]() { int ok = 0; nf[k] += run_graph(iq, window, false, &ok); nok[k] += ok; });
        for (auto& t : th) t.join();
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long tf = 0, tok = 0; for (int k = 0; k < K; k++) { tf += nf[k]; tok += nok[k]; }
    printf("{\"graph_instances\": %d, \"repeat\": %d, \"window_samples\": %zu, \"samples_per_instance\": %zu, \"events\": %lld, \"frames_ok\": %lld, \"seconds\": %.6f, \"msamples_per_s\": %.3f, \"frames_per_s\": %.1f}\n",
           K, R, window, iq.size(), tf, tok, dt, (double)K * R * iq.size() / dt / 1e6, tf / dt);
    return tf ? 0 : 1;
}
