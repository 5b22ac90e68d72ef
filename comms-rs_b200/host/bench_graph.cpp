// Message-rate benchmark of whole node graphs, one OS thread per node (start_nodes!, src/node/mod.rs:276-284), at the
// message sizes the reference's examples really use -- the launch-latency regime, not the 2^28-sample roofline runs:
//
//   cfg1      BASELINE configs[0] as specified (examples/single_thread_bpsk.rs:19,39): BPSK symbols in messages of 4096,
//             x4 / 32-tap RRC pulse shaping (UpsampleNode + BatchFirNode fused), state carried across messages.
//               host edges  : source -> BatchFirNode (Vec in, Vec out: cb_fir_run) -> sink
//               device edges: source (pooled pinned) -> H2DNode -> BatchFirDevNode -> D2HNode -> sink
//   fm_radio  examples/fm_radio.rs:144-164 at its own read size (131072 IQ samples = 262144 bytes per message):
//               source (pooled pinned bytes) -> H2DNode -> FmFrontDevNode (convert, filt1, /5, FM) -> FirRealDevNode
//               (real second stage, /5) -> D2HNode -> sink
//
// Buffers on every device edge come from the library's pool with a high-water mark: the source blocks in
// cb_pool_throttle when it runs too far ahead (the reference's channels are unbounded, src/node/mod.rs:152).
// Prints one JSON line per graph: messages/s, us per message, units/s, pool statistics, and whether the device-edge
// graph's output equals the host-edge graph's bit for bit (same kernels, same batch sizes).
//   usage: bench_graph [messages_cfg1] [messages_fm]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "comms_b200_nodes.hpp"

using namespace comms_b200;
using Clock = std::chrono::steady_clock;

template <class T>
struct VecSource : Node {  // sends `count` messages cycling over `msgs`
    std::vector<std::vector<T>> msgs;
    size_t count, pos = 0;
    NodeSender<std::vector<T>> output;
    VecSource(std::vector<std::vector<T>> m, size_t c) : msgs(std::move(m)), count(c) {}
    bool is_connected() const override { return true; }
    Result<bool> call() override
    {
        if (pos == count) return Result<bool>::Err(NodeError::DataEnd);
        for (auto &s : output) if (!s.send(msgs[pos % msgs.size()])) return Result<bool>::Err(NodeError::CommError);
        ++pos;
        return Result<bool>::Ok(true);
    }
};

struct PinnedSource : Node {  // the same messages as pooled pinned buffers (what an SDR / file reader node would fill)
    std::vector<std::vector<uint8_t>> msgs;
    size_t elem, count, pos = 0;
    NodeSender<Buf> output;
    PinnedSource(std::vector<std::vector<uint8_t>> m, size_t elem_bytes, size_t c) : msgs(std::move(m)), elem(elem_bytes), count(c) {}
    bool is_connected() const override { return true; }
    Result<bool> call() override
    {
        if (pos == count) return Result<bool>::Err(NodeError::DataEnd);
        const auto &m = msgs[pos % msgs.size()];
        if (cb_pool_throttle(0)) return Result<bool>::Err(NodeError::PermanentError);  // back-pressure: wait for the consumers
        Buf b = Buf::pinned(m.size(), m.size() / elem);
        std::memcpy(b.ptr(), m.data(), m.size());
        for (auto &s : output) if (!s.send(b)) return Result<bool>::Err(NodeError::CommError);
        ++pos;
        return Result<bool>::Ok(true);
    }
};

// order-sensitive checksum at memory speed (the sink must not be the slowest node of the graph)
static uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull)
{
    const uint64_t *w = static_cast<const uint64_t *>(p);
    uint64_t a = h, b = h ^ 0x9e3779b97f4a7c15ull, c = ~h, d = h * 3;
    size_t i = 0;
    for (; i + 4 <= n / 8; i += 4) {
        a = (a << 5 | a >> 59) + w[i];
        b = (b << 7 | b >> 57) ^ w[i + 1];
        c = (c << 11 | c >> 53) + w[i + 2];
        d = (d << 13 | d >> 51) ^ w[i + 3];
    }
    for (; i < n / 8; ++i) a = (a << 5 | a >> 59) + w[i];
    const unsigned char *t = static_cast<const unsigned char *>(p);
    for (size_t k = (n / 8) * 8; k < n; ++k) a = (a ^ t[k]) * 1099511628211ull;
    return (a ^ (b << 1) ^ (c << 2) ^ (d << 3)) * 1099511628211ull;
}

static void pool_json(const char *name, int is_device)
{
    size_t live = 0, cached = 0;
    uint64_t hits = 0, misses = 0, waits = 0;
    cb_pool_stats(is_device, &live, &cached, &hits, &misses, &waits);
    std::printf("\"%s\": {\"live_bytes\": %zu, \"cached_bytes\": %zu, \"hits\": %llu, \"misses\": %llu, \"waits\": %llu}", name, live, cached,
                (unsigned long long)hits, (unsigned long long)misses, (unsigned long long)waits);
}

int main(int argc, char **argv)
{
    const size_t m1 = argc > 1 ? (size_t)atol(argv[1]) : 256 * 16, m2 = argc > 2 ? (size_t)atol(argv[2]) : 1024;
    if (cb_init(0) != CB_OK) { std::printf("no CUDA device: %s\n", cb_last_error()); return 2; }
    int fails = 0;
    // back-pressure: the source is held while more than 8 MiB of device or of pinned messages are in flight
    cb_pool_configure(1, (size_t)8 << 20, (size_t)256 << 20, 20000);
    cb_pool_configure(0, (size_t)8 << 20, (size_t)256 << 20, 20000);

    {   // ---------------------------------------------------------------- cfg 1: 4096-symbol messages
        const size_t nb = 4096, L = 4;
        std::vector<c32> taps;
        rrc_taps(32, 4.0, 0.25, taps);
        // 256 distinct messages of BPSK symbols from PrnGen(0xB8, 0x01) (prns.rs:179-180), b -> 2b-1
        std::vector<uint8_t> bits(256 * nb);
        uint64_t st = 0x01;
        cb_prn_bits(0xB8, &st, 8, bits.size(), bits.data());
        std::vector<std::vector<c32>> msgs(256, std::vector<c32>(nb));
        std::vector<std::vector<uint8_t>> raw(256, std::vector<uint8_t>(nb * sizeof(c32)));
        for (size_t m = 0; m < 256; ++m) {
            for (size_t i = 0; i < nb; ++i) msgs[m][i] = c32(2.f * bits[m * nb + i] - 1.f, 0.f);
            std::memcpy(raw[m].data(), msgs[m].data(), raw[m].size());
        }
        uint64_t sum_host = 0, sum_dev = 0;
        double us_host = 0, us_dev = 0;
        {   // host edges
            VecSource<c32> src(msgs, m1);
            BatchFirNode fir(taps, nullptr, 1, (uint32_t)L);
            connect_nodes(src, fir);
            struct S { NodeReceiver<std::vector<c32>> input; } snk;
            connect_nodes(fir, snk);
            auto t0 = Clock::now();
            auto th = start_nodes(src, fir);
            size_t got = 0;
            const size_t warm = m1 / 4;  // the clock starts once the graph is in steady state
            while (auto v = snk.input->recv()) {
                sum_host = fnv(v->data(), v->size() * sizeof(c32), sum_host + 1);
                if (++got == warm) t0 = Clock::now();
            }
            const double dt = std::chrono::duration<double>(Clock::now() - t0).count() * got / (double)(got - warm);
            for (auto &t : th) t.join();
            us_host = 1e6 * dt / got;
            std::printf("{\"graph\": \"cfg1_pulse4\", \"edges\": \"host Vec (cb_fir_run per message)\", \"messages\": %zu, \"symbols_per_message\": %zu, "
                        "\"seconds_steady_state_scaled\": %.4f, \"us_per_message\": %.2f, \"messages_per_s\": %.0f, \"Msymbols_per_s\": %.1f}\n",
                        got, nb, dt, us_host, got / dt, got * nb / dt / 1e6);
            if (got != m1) ++fails;
        }
        {   // device edges
            PinnedSource src(raw, sizeof(c32), m1);
            H2DNode up(sizeof(c32));
            BatchFirDevNode fir(taps, nullptr, 1, (uint32_t)L);
            D2HNode down(sizeof(c32));
            struct S { NodeReceiver<Buf> input; } snk;
            connect_nodes(src, up);
            connect_nodes(up, fir);
            connect_nodes(fir, down);
            connect_nodes(down, snk);
            auto t0 = Clock::now();
            auto th = start_nodes(src, up, fir, down);
            size_t got = 0;
            const size_t warm = m1 / 4;
            while (auto v = snk.input->recv()) {
                cb_buf_sync(v->raw());
                sum_dev = fnv(v->ptr(), v->len * sizeof(c32), sum_dev + 1);
                if (++got == warm) t0 = Clock::now();
            }
            const double dt = std::chrono::duration<double>(Clock::now() - t0).count() * got / (double)(got - warm);
            for (auto &t : th) t.join();
            us_dev = 1e6 * dt / got;
            std::printf("{\"graph\": \"cfg1_pulse4\", \"edges\": \"device (pooled pinned -> H2D -> BatchFirDevNode -> D2H)\", \"messages\": %zu, "
                        "\"symbols_per_message\": %zu, \"seconds\": %.4f, \"us_per_message\": %.2f, \"messages_per_s\": %.0f, \"Msymbols_per_s\": %.1f, "
                        "\"bit_identical_to_host_edges\": %s, ", got, nb, dt, us_dev, got / dt, got * nb / dt / 1e6, sum_dev == sum_host ? "true" : "false");
            pool_json("pool_device", 1);
            std::printf(", ");
            pool_json("pool_pinned", 0);
            std::printf("}\n");
            if (got != m1 || sum_dev != sum_host) ++fails;
        }
    }
    {   // ---------------------------------------------------------------- fm_radio at 131072-sample messages
        const size_t nb = 131072;
        std::vector<c32> lp(63);
        for (int k = 0; k < 63; ++k) {  // low-pass of the same shape as examples/fm_radio.rs:30-52 (63 real taps)
            const double t = (k - 31) / 5.0, sinc = t == 0 ? 1.0 : std::sin(M_PI * t) / (M_PI * t);
            lp[k] = c32((float)(sinc * (0.54 - 0.46 * std::cos(2 * M_PI * k / 62)) / 5.0), 0.f);
        }
        std::vector<std::vector<uint8_t>> raw(16, std::vector<uint8_t>(2 * nb));
        uint32_t r = 12345;
        for (auto &m : raw) for (auto &b : m) { r = r * 1664525u + 1013904223u; b = (uint8_t)(r >> 24); }
        PinnedSource src(raw, 2, m2);
        H2DNode up(2);
        FmFrontDevNode front(lp, 5);
        FirRealDevNode back(lp, 5);
        D2HNode down(sizeof(float));
        struct S { NodeReceiver<Buf> input; } snk;
        connect_nodes(src, up);
        connect_nodes(up, front);
        connect_nodes(front, back);
        connect_nodes(back, down);
        connect_nodes(down, snk);
        auto t0 = Clock::now();
        auto th = start_nodes(src, up, front, back, down);
        size_t got = 0, audio = 0;
        uint64_t sum = 0;
        const size_t warm = m2 / 4;
        while (auto v = snk.input->recv()) {
            cb_buf_sync(v->raw());
            sum = fnv(v->ptr(), v->len * sizeof(float), sum + 1);
            audio += v->len;
            if (++got == warm) t0 = Clock::now();
        }
        const double dt = std::chrono::duration<double>(Clock::now() - t0).count() * got / (double)(got - warm);
        for (auto &t : th) t.join();
        std::printf("{\"graph\": \"fm_radio\", \"edges\": \"device (pooled pinned bytes -> H2D -> FmFrontDevNode -> FirRealDevNode -> D2H)\", "
                    "\"messages\": %zu, \"iq_samples_per_message\": %zu, \"audio_samples\": %zu, \"seconds\": %.4f, \"us_per_message\": %.2f, "
                    "\"messages_per_s\": %.0f, \"Msamples_per_s\": %.1f, \"realtime_factor_at_1.14MSps\": %.0f, \"checksum\": \"%016llx\", ",
                    got, nb, audio, dt, 1e6 * dt / got, got / dt, got * nb / dt / 1e6, got * nb / dt / 1.14e6, (unsigned long long)sum);
        pool_json("pool_device", 1);
        std::printf(", ");
        pool_json("pool_pinned", 0);
        std::printf("}\n");
        if (got != m2 || audio != m2 * 5243) ++fails;  // ceil(ceil(131072/5)/5) = 5243 per message
    }
    cb_pool_configure(1, 0, (size_t)1 << 30, 10000);
    cb_pool_configure(0, 0, (size_t)1 << 30, 10000);
    std::printf(fails ? "bench_graph FAILED\n" : "bench_graph ok\n");
    return fails ? 1 : 0;
}
