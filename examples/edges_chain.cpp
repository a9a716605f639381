// examples/edges_chain.cpp — the host edges of a graph, as the reference wires them (src/dsp/source.h:66-107,
// src/dsp/routing.h:9-60, src/dsp/sink.h:122-178): a HandlerSource whose callback fills `writeBuf` from a file, a
// Splitter fanning the stream out to two branches, one branch = fused VFO -> FloatFMDemod -> FileSink<float>, the
// other = FrequencyXlator -> NullSink. Only the first device block uploads the host buffer, only the FileSink downloads.
//
//   g++ -O2 -std=c++17 -Iinclude examples/edges_chain.cpp -Lqdsp_b200 -lqdsp_b200 -lpthread -o edges_chain
//   ./edges_chain in.cf32 out.f32 block
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <dsp/demodulator.h>
#include <dsp/processing.h>
#include <dsp/routing.h>
#include <dsp/sink.h>
#include <dsp/source.h>
#include <dsp/vfo.h>

struct Feed {
    FILE* f;
    int block;
    std::atomic<long long> fed{0};
    std::atomic<bool> eof{false};
};
static int feedHandler(dsp::complex_t* data, void* ctx) {
    Feed* fd = (Feed*)ctx;
    const size_t n = fread(data, sizeof(dsp::complex_t), fd->block, fd->f);
    if (n == 0) {
        fd->eof = true;
        std::this_thread::sleep_for(std::chrono::milliseconds(20));   // idle: the source is stopped from main()
        return 0;
    }
    fd->fed += (long long)n;
    return (int)n;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s in.cf32 out.f32 block\n", argv[0]); return 2; }
    if (qdsp_device_count() <= 0) { fprintf(stderr, "no CUDA device: %s\n", qdsp_last_error()); return 3; }
    Feed fd;
    fd.f = fopen(argv[1], "rb");
    fd.block = atoi(argv[3]);
    if (!fd.f || fd.block <= 0) { perror(argv[1]); return 1; }

    dsp::HandlerSource<dsp::complex_t> src(feedHandler, &fd);
    dsp::Splitter<dsp::complex_t> split(&src.out);
    dsp::stream<dsp::complex_t> toChain, toXlate;
    split.bindStream(&toChain);
    split.bindStream(&toXlate);
    dsp::FusedVFOFloatFMDemod chain;
    chain.init(&toChain, 250e3, 2.4e6, 48e3, 48e3, 5e3);
    dsp::FileSink<float> fsink(&chain.out, argv[2]);
    dsp::FrequencyXlator<dsp::complex_t> xl(&toXlate, 2.4e6, -250e3);
    dsp::NullSink<dsp::complex_t> nsink(&xl.out);
    if (!fsink.isOpen()) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }

    nsink.start(); xl.start(); fsink.start(); chain.start(); split.start(); src.start();
    while (!fd.eof.load()) { std::this_thread::sleep_for(std::chrono::milliseconds(5)); }
    std::this_thread::sleep_for(std::chrono::milliseconds(300));     // let the tail of the pipeline drain
    src.stop(); split.stop(); chain.stop(); fsink.stop(); xl.stop(); nsink.stop();
    fclose(fd.f);
    printf("fed %lld samples through HandlerSource -> Splitter -> {fused chain -> FileSink, Xlator -> NullSink}\n", fd.fed.load());
    return 0;
}
