// Graph-level conformance of the GPU nodes, written like the reference's own node tests:
// a tiny source node and a check node around the node under test, wired with connect_nodes,
// one thread per node (src/filter/fir_node.rs:233-337, src/mixer.rs:158-245,
// src/fft/fft_node.rs:179-262, tests/node_test.rs:8-50).  Expected values are the reference's
// golden vectors.  Needs a B200; exits non-zero on any mismatch.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "comms_b200_nodes.hpp"

using namespace comms_b200;

template <class T>
struct SomeSamples : Node {  // source: sends each element once, then stops (DataEnd downstream)
    std::vector<T> samples;
    size_t pos = 0;
    NodeSender<T> output;
    explicit SomeSamples(std::vector<T> s) : samples(std::move(s)) {}
    bool is_connected() const override { return true; }
    Result<bool> call() override
    {
        if (pos == samples.size()) return Result<bool>::Err(NodeError::DataEnd);
        for (auto &s : output) if (!s.send(samples[pos])) return Result<bool>::Err(NodeError::CommError);
        ++pos;
        return Result<bool>::Ok(true);
    }
};

template <class T>
struct Collect {  // check node: drains its input on the calling thread
    NodeReceiver<T> input;
    std::vector<T> got;
    void drain() { while (auto v = input->recv()) got.push_back(std::move(*v)); }
};

static int fails = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { ++fails; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

static bool close_to(c32 a, std::complex<double> b, double tol) { return std::abs(std::complex<double>(a) - b) < tol; }

int main()
{
    if (cb_init(0) != CB_OK) { std::printf("no CUDA device: %s\n", cb_last_error()); return 2; }
    const std::vector<c32> fir_in = {{1, 2}, {3, 4}, {5, 6}, {7, 8}, {9, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const std::vector<c32> fir_taps = {{9, 0}, {8, 7}, {6, 5}, {4, 3}, {2, 1}};
    const std::vector<c32> fir_expect = {{9, 18}, {21, 59}, {37, 124}, {57, 205}, {81, 204}, {78, 196}, {62, 115}, {42, 50}, {18, 9}, {0, 0}};

    {   // FirNode, sample at a time (fir_node.rs:259-314)
        SomeSamples<c32> src(fir_in);
        FirNode fir(fir_taps);
        Collect<c32> chk;
        connect_nodes(src, fir);
        connect_nodes(fir, chk);
        auto th = start_nodes(src, fir);
        chk.drain();
        for (auto &t : th) t.join();
        EXPECT(chk.got.size() == fir_expect.size(), "FirNode count %zu", chk.got.size());
        for (size_t i = 0; i < chk.got.size() && i < fir_expect.size(); ++i) EXPECT(chk.got[i] == fir_expect[i], "FirNode[%zu]", i);
    }
    {   // BatchFirNode, batches of two (fir_node.rs:368-425); fan-out to two edges (clone per edge)
        std::vector<std::vector<c32>> batches;
        for (size_t i = 0; i < fir_in.size(); i += 2) batches.push_back({fir_in[i], fir_in[i + 1]});
        SomeSamples<std::vector<c32>> src(batches);
        BatchFirNode fir(fir_taps);
        Collect<std::vector<c32>> a, b;
        connect_nodes(src, fir);
        connect_nodes(fir, a);
        connect_nodes(fir, b);
        auto th = start_nodes(src, fir);
        a.drain();
        b.drain();
        for (auto &t : th) t.join();
        std::vector<c32> flat;
        for (auto &v : a.got) flat.insert(flat.end(), v.begin(), v.end());
        EXPECT(flat == fir_expect, "BatchFirNode output");
        EXPECT(a.got == b.got, "fan-out edges differ");
        auto st = fir.state();  // newest first
        EXPECT(st.size() == 5 && st[0] == c32(0, 0) && st[4] == c32(0, 0), "state after run");
    }
    {   // PulseNode with rect taps x4 (pulse.rs:129-183)
        const std::vector<c32> sym = {{-1, -1}, {1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        SomeSamples<c32> src(sym);
        PulseNode pulse(std::vector<c32>(4, c32(1, 0)), 4);
        Collect<std::vector<c32>> chk;
        connect_nodes(src, pulse);
        connect_nodes(pulse, chk);
        auto th = start_nodes(src, pulse);
        chk.drain();
        for (auto &t : th) t.join();
        EXPECT(chk.got.size() == 5, "PulseNode messages %zu", chk.got.size());
        for (size_t i = 0; i < chk.got.size(); ++i) {
            EXPECT(chk.got[i].size() == 4, "PulseNode vec len");
            for (auto v : chk.got[i]) EXPECT(v == sym[i], "PulseNode value");
        }
    }
    {   // MixerNode::new(0.123, None) and (0.123, Some(0.1)) (mixer.rs:184-223, 274-313)
        const std::vector<c32> in = {{1, 2}, {3, 4}, {5, 6}, {7, 8}, {9, 0}};
        const std::complex<double> e0[] = {{1.0, 2.0}, {2.486574736, 4.337850399}, {3.388313374, 7.036997405}, {3.643356072, 9.986288426}, {7.932508585, 4.251506503}};
        const std::complex<double> e1[] = {{0.795337332, 2.089841747}, {2.041089794, 4.564422467}, {2.668858427, 7.340108630}, {2.628189174, 10.300127265}, {7.468436663, 5.022196114}};
        for (int v = 0; v < 2; ++v) {
            SomeSamples<c32> src(in);
            MixerNode mix(0.123, v ? std::optional<double>(0.1) : std::nullopt);
            Collect<c32> chk;
            connect_nodes(src, mix);
            connect_nodes(mix, chk);
            auto th = start_nodes(src, mix);
            chk.drain();
            for (auto &t : th) t.join();
            EXPECT(chk.got.size() == 5, "MixerNode count");
            for (size_t i = 0; i < chk.got.size(); ++i) EXPECT(close_to(chk.got[i], (v ? e1 : e0)[i], 2e-6), "MixerNode[%d][%zu]", v, i);
        }
    }
    {   // FFTBatchNode::new(10, false) and FFTSampleNode (fft_node.rs:194-244, 286-332)
        std::vector<c32> x;
        for (int k = 1; k <= 10; ++k) x.push_back(c32(0.1f * k, 0.1f * k));
        const std::complex<double> e[] = {{5.5, 5.5}, {-2.03884, 1.03884}, {-1.18819, 0.18819}, {-0.86327, -0.13673}, {-0.66246, -0.33754},
                                          {-0.5, -0.5}, {-0.33754, -0.66246}, {-0.13673, -0.86327}, {0.18819, -1.18819}, {1.03884, -2.03884}};
        SomeSamples<std::vector<c32>> src({x});
        FFTBatchNode fft(10, false);
        Collect<std::vector<c32>> chk;
        connect_nodes(src, fft);
        connect_nodes(fft, chk);
        auto th = start_nodes(src, fft);
        chk.drain();
        for (auto &t : th) t.join();
        EXPECT(chk.got.size() == 1 && chk.got[0].size() == 10, "FFTBatchNode shape");
        if (chk.got.size() == 1) for (size_t i = 0; i < 10; ++i) EXPECT(close_to(chk.got[0][i], e[i], 1e-5), "FFTBatchNode[%zu]", i);

        SomeSamples<c32> ssrc(x);
        FFTSampleNode sfft(10, false);
        Collect<std::vector<c32>> schk;
        connect_nodes(ssrc, sfft);
        connect_nodes(sfft, schk);
        auto th2 = start_nodes(ssrc, sfft);
        schk.drain();
        for (auto &t : th2) t.join();
        EXPECT(schk.got.size() == 1, "FFTSampleNode emits once per fft_size samples (aggregate), got %zu", schk.got.size());
        if (schk.got.size() == 1) for (size_t i = 0; i < 10; ++i) EXPECT(close_to(schk.got[0][i], e[i], 1e-5), "FFTSampleNode[%zu]", i);

        // wrong frame length: DataError, the node stops and the graph winds down (DataEnd downstream)
        SomeSamples<std::vector<c32>> bad({std::vector<c32>(7), x});
        FFTBatchNode fft2(10, false);
        Collect<std::vector<c32>> bchk;
        connect_nodes(bad, fft2);
        connect_nodes(fft2, bchk);
        auto th3 = start_nodes(bad, fft2);
        bchk.drain();
        for (auto &t : th3) t.join();
        EXPECT(bchk.got.empty(), "bad frame must stop the node");
        EXPECT(!fft2.run(std::vector<c32>(7)).is_ok() && fft2.run(std::vector<c32>(7)).err == NodeError::DataError, "DataError mapping");
    }
    {   // Decimate / Upsample (resample_node.rs:139-175), then fm_radio-style chain: FIR -> Decimate -> FMDemod
        DecimateNode<int> d2(2), d100(100), d0(0);
        UpsampleNode<int> u4(4);
        std::vector<int> six = {1, 2, 3, 4, 5, 6};
        EXPECT(*d2.run(six).ok == (std::vector<int>{1, 3, 5}), "decimate 2");
        EXPECT(*d100.run(six).ok == (std::vector<int>{1}), "decimate 100");
        EXPECT(*d0.run(six).ok == six, "decimate 0");
        EXPECT(*u4.run({1, 2}).ok == (std::vector<int>{1, 0, 0, 0, 2, 0, 0, 0}), "upsample 4");

        std::vector<std::vector<c32>> batches(4, std::vector<c32>(1000));
        double ph = 0;
        for (auto &b : batches) for (auto &s : b) { s = c32((float)std::cos(ph), (float)std::sin(ph)); ph += 0.05; }
        SomeSamples<std::vector<c32>> src(batches);
        std::vector<c32> lp(31, c32(1.0f / 31, 0));
        BatchFirNode filt(lp);
        DecimateNode<c32> dec(5);
        FMDemodNode fm;
        Collect<std::vector<float>> chk;
        connect_nodes(src, filt);
        connect_nodes(filt, dec);
        connect_nodes(dec, fm);
        connect_nodes(fm, chk);
        auto th = start_nodes(src, filt, dec, fm);
        chk.drain();
        for (auto &t : th) t.join();
        EXPECT(chk.got.size() == 4, "chain messages");
        size_t n = 0;
        for (auto &v : chk.got) {
            EXPECT(v.size() == 200, "chain batch length %zu", v.size());
            for (size_t i = 0; i < v.size(); ++i, ++n) if (n >= 8) EXPECT(std::fabs(v[i] - 0.25f) < 1e-4, "FM value %f at %zu", v[i], n);
        }
    }
    if (fails) { std::printf("%d failure(s)\n", fails); return 1; }
    std::printf("host graph tests ok\n");
    return 0;
}
