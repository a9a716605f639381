// include/dsp/processing.h — FrequencyXlator<T>, AGC, FeedForwardAGC<T>, ComplexAGC (reference
// src/dsp/processing.h:8-300). The element-wise helpers of that header that are off the hot path (DelayImag,
// Volume, Squelch, Packer, Threshold) are not part of this library (out of scope).
#pragma once
#include <type_traits>
#include <dsp/block.h>

namespace dsp {
    template <class T>
    class FrequencyXlator : public generic_block<FrequencyXlator<T>> {
        using base = generic_block<FrequencyXlator<T>>;
        static_assert(std::is_same<T, complex_t>::value, "the reference implements FrequencyXlator for complex_t only");

    public:
        FrequencyXlator() {}
        FrequencyXlator(stream<complex_t>* in, float sampleRate, float freq) { init(in, sampleRate, freq); }
        ~FrequencyXlator() {
            base::stop();
            if (h) { qdsp_xlator_destroy(h); }
        }
        void init(stream<complex_t>* in, float sampleRate, float freq) {
            _in = in;
            _sampleRate = sampleRate;
            _freq = freq;
            if (h) { qdsp_xlator_destroy(h); }
            h = qdsp_xlator_create(_sampleRate, _freq);
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        // (sic) the reference names its setInput "setInputSize" (processing.h:26); both spellings work here
        void setInputSize(stream<complex_t>* in) { setInput(in); }
        void setInput(stream<complex_t>* in) { base::rebindInput(_in, in); }
        void setSampleRate(float sampleRate) {
            _sampleRate = sampleRate;
            qdsp_xlator_set_frequency(h, _sampleRate, _freq);
        }
        float getSampleRate() { return _sampleRate; }
        void setFrequency(float freq) {
            _freq = freq;
            qdsp_xlator_set_frequency(h, _sampleRate, _freq);
        }
        float getFrequency() { return _freq; }
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const long long n = qdsp_xlator_process(h, _in->readDev(), out.writeDev(), count, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, base::cuStream)) { return -1; }
            return count;
        }

        stream<complex_t> out;

    private:
        float _sampleRate = 1, _freq = 0;
        stream<complex_t>* _in = nullptr;
        qdsp_xlator* h = nullptr;
    };

    class AGC : public generic_block<AGC> {
    public:
        AGC() {}
        AGC(stream<float>* in, float fallRate, float sampleRate) { init(in, fallRate, sampleRate); }
        ~AGC() {
            generic_block<AGC>::stop();
            if (h) { qdsp_agc_destroy(h); }
        }
        void init(stream<float>* in, float fallRate, float sampleRate) {
            _in = in;
            _sampleRate = sampleRate;
            _fallRate = fallRate;
            rebuild();
            generic_block<AGC>::registerInput(_in);
            generic_block<AGC>::registerOutput(&out);
        }
        void setInput(stream<float>* in) { generic_block<AGC>::rebindInput(_in, in); }
        void setSampleRate(float sampleRate) {
            std::lock_guard<std::mutex> lck(generic_block<AGC>::ctrlMtx);
            _sampleRate = sampleRate;
            rebuild();
        }
        void setFallRate(float fallRate) {
            std::lock_guard<std::mutex> lck(generic_block<AGC>::ctrlMtx);
            _fallRate = fallRate;
            rebuild();
        }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const int one = count;  // one run() call == one block of the decay schedule (processing.h:123)
            const long long n = qdsp_agc_process(h, _in->readDev(), out.writeDev(), count, &one, 1, 0, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<float> out;

    private:
        void rebuild() {
            float level = 0.0f;
            if (h) { qdsp_agc_get_state(h, &level); qdsp_agc_destroy(h); }
            h = qdsp_agc_create(_fallRate, _sampleRate);
            qdsp_agc_set_state(h, level);
        }
        float _fallRate = 0, _sampleRate = 1;
        stream<float>* _in = nullptr;
        qdsp_agc* h = nullptr;
    };

    template <class T>
    class FeedForwardAGC : public generic_block<FeedForwardAGC<T>> {
        using base = generic_block<FeedForwardAGC<T>>;

    public:
        FeedForwardAGC() {}
        FeedForwardAGC(stream<T>* in) { init(in); }
        ~FeedForwardAGC() {
            base::stop();
            if (h) { qdsp_ffagc_destroy(h); }
        }
        void init(stream<T>* in) {
            _in = in;
            h = qdsp_ffagc_create(std::is_same<T, float>::value ? QDSP_F32 : QDSP_CF32);
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        void setInput(stream<T>* in) { base::rebindInput(_in, in); }
        // emits the valid (toProcess) outputs; the reference swaps `count` elements of which only toProcess are
        // fresh (processing.h:221) -- consumers here see exactly the fresh ones
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const long long n = qdsp_ffagc_process(h, _in->readDev(), out.writeDev(), count, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (n == 0) { return count; }
            if (!out.swapDevice((int)n, base::cuStream)) { return -1; }
            return (int)n;
        }

        stream<T> out;

    private:
        stream<T>* _in = nullptr;
        qdsp_ffagc* h = nullptr;
    };

    class ComplexAGC : public generic_block<ComplexAGC> {
    public:
        ComplexAGC() {}
        ComplexAGC(stream<complex_t>* in, float setPoint, float maxGain, float rate) { init(in, setPoint, maxGain, rate); }
        ~ComplexAGC() {
            generic_block<ComplexAGC>::stop();
            if (h) { qdsp_cagc_destroy(h); }
        }
        void init(stream<complex_t>* in, float setPoint, float maxGain, float rate) {
            _in = in;
            _setPoint = setPoint;
            _maxGain = maxGain;
            _rate = rate;
            rebuild();
            generic_block<ComplexAGC>::registerInput(_in);
            generic_block<ComplexAGC>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<ComplexAGC>::rebindInput(_in, in); }
        void setSetPoint(float setPoint) { _setPoint = setPoint; rebuild(); }
        void setMaxGain(float maxGain) { _maxGain = maxGain; rebuild(); }
        void setRate(float rate) { _rate = rate; rebuild(); }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const long long n = qdsp_cagc_process(h, _in->readDev(), out.writeDev(), count, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<complex_t> out;

    private:
        void rebuild() {
            float gain = 1.0f;
            if (h) { qdsp_cagc_get_state(h, &gain); qdsp_cagc_destroy(h); }
            h = qdsp_cagc_create(_setPoint, _maxGain, _rate);
            qdsp_cagc_set_state(h, gain);
        }
        float _setPoint = 1.0f, _maxGain = 65535.0f, _rate = 1e-3f;
        stream<complex_t>* _in = nullptr;
        qdsp_cagc* h = nullptr;
    };
}
