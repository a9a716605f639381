// examples/am_chain.cpp — a second main.cpp-style graph over the element-wise blocks of the dsp:: API:
//   input -> Squelch(-30 dB) -> AMDemod -> Volume<float>(0.5) -> MonoToStereo -> StereoToMono -> HandlerSink
// Every run() executes on the GPU; the five interior streams stay in HBM, H2D/D2H happen at the two host edges.
//
//   g++ -O2 -std=c++17 -Iinclude examples/am_chain.cpp -Lqdsp_b200 -lqdsp_b200 -lpthread -o am_chain
//   ./am_chain in.cf32 out.f32 [block]
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <vector>
#include <dsp/audio.h>
#include <dsp/demodulator.h>
#include <dsp/processing.h>
#include <dsp/sink.h>

static std::vector<float> audio;
static std::atomic<long long> received{0};
static void audioHandler(float* data, int count, void* ctx) {
    audio.insert(audio.end(), data, data + count);
    received += count;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.cf32 out.f32 [block]\n", argv[0]); return 2; }
    const int block = argc > 3 ? atoi(argv[3]) : 100000;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    if (qdsp_device_count() <= 0) { fprintf(stderr, "no CUDA device: %s\n", qdsp_last_error()); return 3; }

    dsp::stream<dsp::complex_t> input;
    dsp::Squelch squelch(&input, -30.0f);
    dsp::AMDemod am(&squelch.out);
    dsp::Volume<float> vol(&am.out, 0.5f);
    vol.setVolume(0.5f);
    dsp::MonoToStereo m2s(&vol.out);
    dsp::StereoToMono s2m(&m2s.out);
    dsp::HandlerSink<float> sink(&s2m.out, audioHandler, NULL);
    squelch.start(); am.start(); vol.start(); m2s.start(); s2m.start(); sink.start();

    long long fed = 0;
    for (;;) {
        const size_t n = fread(input.writeBuf, sizeof(dsp::complex_t), block, f);
        if (n == 0) { break; }
        if (!input.swap((int)n)) { break; }
        fed += (long long)n;
    }
    fclose(f);
    while (received.load() < fed) { std::this_thread::yield(); }
    sink.stop(); s2m.stop(); m2s.stop(); vol.stop(); am.stop(); squelch.stop();

    FILE* o = fopen(argv[2], "wb");
    fwrite(audio.data(), sizeof(float), audio.size(), o);
    fclose(o);
    printf("fed %lld samples, wrote %zu audio samples\n", fed, audio.size());
    return 0;
}
