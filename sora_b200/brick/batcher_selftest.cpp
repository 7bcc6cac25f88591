// CPU-only exercise of B200StreamBatcher (window staging in one arena, rounds closed by count or deadline): the C ABI is stubbed, no device.
// The stub "decodes" a window to the checksum of the samples it was handed, so overlapping or stale arena regions show up as wrong events.
#include "b200_bricks.hpp"
#include <thread>
#include <atomic>
#include <cstdio>
static std::atomic<long> g_calls{0}, g_streams{0};
extern "C" {
int sb200_create(int, const sb200_cfg*, sb200_handle** h) { *h = (sb200_handle*)0x1; return SB200_OK; }
void sb200_destroy(sb200_handle*) {}
void* sb200_host_alloc(size_t n) { return malloc(n); }
void sb200_host_free(void* p) { free(p); }
int sb200_rx11a_streams(sb200_handle*, const int16_t* iq, uint64_t total, const uint64_t* off, const uint32_t* len, uint32_t n, uint32_t K, uint8_t*, uint32_t, sb200_frame_result* res, uint32_t* sidx, uint32_t* cnt, void*) {
    g_calls++; g_streams += n;
    for (uint32_t s = 0; s < n; s++) {           // one event per stream whose "length" is a checksum of the staged samples: catches overlapping / stale regions
        if (off[s] + len[s] > total) return SB200_E_INVALID;
        uint32_t sum = 0; for (uint32_t i = 0; i < len[s]; i++) sum += (uint16_t)iq[2 * (off[s] + i)];
        cnt[s] = 1; memset(&res[(size_t)s * K], 0, sizeof(sb200_frame_result)); res[(size_t)s * K].status = 1; res[(size_t)s * K].crc32 = sum; res[(size_t)s * K].length = 4; sidx[(size_t)s * K] = len[s];
    }
    return SB200_OK;
}
}
int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 8, ROUNDS = argc > 2 ? atoi(argv[2]) : 200;
    B200StreamBatcher::Get().Configure(K, 500);
    std::atomic<long> bad{0};
    std::vector<std::thread> th;
    for (int k = 0; k < K; k++) th.emplace_back([&, k]() {
        std::vector<COMPLEX16> buf; std::vector<B200Event> ev;
        for (int r = 0; r < ROUNDS; r++) {
            const size_t n = 1000 + (size_t)((k * 7919 + r * 104729) % 50000);
            buf.resize(n); uint32_t sum = 0;
            for (size_t i = 0; i < n; i++) { buf[i].re = (short)(i * 31 + k + r); buf[i].im = 0; sum += (uint16_t)buf[i].re; }
            int rc = B200StreamBatcher::Get().Decode(buf.data(), n, 4, ev);
            if (rc != SB200_OK || ev.size() != 1 || ev[0].r.crc32 != sum || ev[0].end_sample != n) bad++;
            if ((k + r) % 5 == 0) std::this_thread::sleep_for(std::chrono::microseconds(300 * (k % 3)));   // stragglers: rounds close on the deadline
        }
    });
    for (auto& t : th) t.join();
    printf("threads %d rounds %d: device calls %ld (%.2f windows per call), bad %ld\n", K, ROUNDS, g_calls.load(), (double)g_streams / g_calls, bad.load());
    return bad ? 1 : 0;
}
