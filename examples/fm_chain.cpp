// examples/fm_chain.cpp — a main.cpp-style graph written against the dsp:: block API exactly as one would
// against the reference headers: feed IQ with stream<T>::writeBuf/swap (reference src/main.cpp:74-78), wire
// VFO -> FloatFMDemod -> HandlerSink, start the blocks, collect audio. Compiled against include/dsp and linked
// with libqdsp_b200.so, every run() executes on the GPU and the interior streams stay in HBM.
//
//   g++ -O2 -std=c++17 -Iinclude examples/fm_chain.cpp -Lqdsp_b200 -lqdsp_b200 -lpthread -o fm_chain
//   ./fm_chain in.cf32 out.f32 [block] [fused]     (raw interleaved cf32 in, raw f32 audio out)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <vector>
#include <dsp/demodulator.h>
#include <dsp/sink.h>
#include <dsp/vfo.h>

static std::vector<float> audio;
static std::atomic<long long> received{0};
static void audioHandler(float* data, int count, void* ctx) {
    audio.insert(audio.end(), data, data + count);
    received += count;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.cf32 out.f32 [block] [fused]\n", argv[0]); return 2; }
    const int block = argc > 3 ? atoi(argv[3]) : 819200;
    const bool fused = argc > 4 && !strcmp(argv[4], "fused");
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    if (qdsp_device_count() <= 0) { fprintf(stderr, "no CUDA device: %s\n", qdsp_last_error()); return 3; }

    dsp::stream<dsp::complex_t> input;
    dsp::VFO vfo;
    dsp::FloatFMDemod demod;
    dsp::FusedVFOFloatFMDemod chain;
    dsp::HandlerSink<float> sink;
    if (fused) {
        chain.init(&input, 250e3, 2.4e6, 48e3, 48e3, 5e3);
        sink.init(&chain.out, audioHandler, NULL);
        chain.start();
    } else {
        vfo.init(&input, 250e3, 2.4e6, 48e3, 48e3);
        demod.init(vfo.out, 48e3, 5e3);
        sink.init(&demod.out, audioHandler, NULL);
        vfo.start();
        demod.start();
    }
    sink.start();

    long long fed = 0, expect = 0;
    for (;;) {
        const size_t n = fread(input.writeBuf, sizeof(dsp::complex_t), block, f);
        if (n == 0) { break; }
        if (!input.swap((int)n)) { break; }
        fed += (long long)n;
        expect += (long long)n / 50;  // I = 1, D = 50: every run() block yields count / 50 outputs
    }
    fclose(f);
    while (received.load() < expect) { std::this_thread::yield(); }
    sink.stop();
    if (fused) { chain.stop(); } else { demod.stop(); vfo.stop(); }

    FILE* o = fopen(argv[2], "wb");
    fwrite(audio.data(), sizeof(float), audio.size(), o);
    fclose(o);
    printf("fed %lld samples, wrote %zu audio samples (%s path)\n", fed, audio.size(), fused ? "fused" : "block-by-block");
    return 0;
}
