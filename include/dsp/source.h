// include/dsp/source.h — SineSource (reference src/dsp/source.h:5-64): the NCO phasor generated on the device, and
// HandlerSource<T> (:66-107): a user callback fills `out.writeBuf` (pinned host memory); the first device block
// downstream uploads it once.
#pragma once
#include <dsp/block.h>

namespace dsp {
    class SineSource : public generic_block<SineSource> {
    public:
        SineSource() {}
        SineSource(int blockSize, float sampleRate, float freq) { init(blockSize, sampleRate, freq); }
        ~SineSource() {
            generic_block<SineSource>::stop();
            if (h) { qdsp_sinesource_destroy(h); }
        }
        void init(int blockSize, float sampleRate, float freq) {
            _blockSize = blockSize;
            _sampleRate = sampleRate;
            _freq = freq;
            if (h) { qdsp_sinesource_destroy(h); }
            h = qdsp_sinesource_create(_sampleRate, _freq);
            generic_block<SineSource>::registerOutput(&out);
        }
        void setBlockSize(int blockSize) {
            std::lock_guard<std::mutex> lck(generic_block<SineSource>::ctrlMtx);
            generic_block<SineSource>::tempStop();
            _blockSize = blockSize;
            generic_block<SineSource>::tempStart();
        }
        int getBlockSize() { return _blockSize; }
        void setSampleRate(float sampleRate) {
            _sampleRate = sampleRate;
            qdsp_sinesource_configure(h, _sampleRate, _freq);
        }
        float getSampleRate() { return _sampleRate; }
        void setFrequency(float freq) {
            _freq = freq;
            qdsp_sinesource_configure(h, _sampleRate, _freq);
        }
        float getFrequency() { return _freq; }
        int run() override {
            out.acquireWriteDev(cuStream);
            if (qdsp_sinesource_process(h, out.writeDev(), _blockSize, cuStream) < 0) { return -1; }
            if (!out.swapDevice(_blockSize, cuStream)) { return -1; }
            return _blockSize;
        }

        stream<complex_t> out;

    private:
        int _blockSize = 0;
        float _sampleRate = 1, _freq = 0;
        qdsp_sinesource* h = nullptr;
    };

    template <class T>
    class HandlerSource : public generic_block<HandlerSource<T>> {
        using base = generic_block<HandlerSource<T>>;

    public:
        HandlerSource() {}
        HandlerSource(int (*handler)(T* data, void* ctx), void* ctx) { init(handler, ctx); }
        ~HandlerSource() { base::stop(); }
        void init(int (*handler)(T* data, void* ctx), void* ctx) {
            _handler = handler;
            _ctx = ctx;
            base::registerOutput(&out);
        }
        void setHandler(int (*handler)(T* data, void* ctx), void* ctx) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            _handler = handler;
            _ctx = ctx;
            base::tempStart();
        }
        int run() override {
            const int count = _handler(out.writeBuf, _ctx);
            if (count < 0) { return -1; }
            if (!out.swap(count)) { return -1; }
            return count;
        }

        stream<T> out;

    private:
        int (*_handler)(T* data, void* ctx) = nullptr;
        void* _ctx = nullptr;
    };
}
