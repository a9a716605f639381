// include/dsp/filter.h — FIR<T> and BFMDeemp with the reference's interface (src/dsp/filter.h:9-173);
// run() bodies enqueue sm_100a kernels through the C ABI instead of looping over VOLK dot products.
#pragma once
#include <type_traits>
#include <vector>
#include <dsp/block.h>
#include <dsp/window.h>

namespace dsp {
    template <class T>
    class FIR : public generic_block<FIR<T>> {
        using base = generic_block<FIR<T>>;
        static_assert(std::is_same<T, float>::value || std::is_same<T, complex_t>::value, "FIR<float> or FIR<complex_t>");

    public:
        FIR() {}
        FIR(stream<T>* in, dsp::filter_window::generic_window* window) { init(in, window); }
        ~FIR() {
            base::stop();
            if (h) { qdsp_fir_destroy(h); }
        }

        void init(stream<T>* in, dsp::filter_window::generic_window* window) {
            _in = in;
            loadTaps(window);
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        void setInput(stream<T>* in) { base::rebindInput(_in, in); }
        // the reference swaps taps without stopping the worker (filter.h:43-49, a race); here the worker is
        // paused so the new taps take effect between two run() calls
        void updateWindow(dsp::filter_window::generic_window* window) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            loadTaps(window);
            base::tempStart();
        }
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const long long n = qdsp_fir_process(h, _in->readDev(), out.writeDev(), count, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, base::cuStream)) { return -1; }
            return count;
        }

        stream<T> out;

    private:
        void loadTaps(dsp::filter_window::generic_window* window) {
            const int tapCount = window->getTapCount();
            // RRCTaps::createTaps rounds an even count up and writes tapCount | 1 taps (window.h:186): the reference
            // overflows its buffer by one float there and filters with the first tapCount of them; same taps, no overflow
            std::vector<float> taps((size_t)tapCount | 1);
            window->createTaps(taps.data(), tapCount);
            if (!h) { h = qdsp_fir_create(std::is_same<T, float>::value ? QDSP_F32 : QDSP_CF32, taps.data(), tapCount); }
            else { qdsp_fir_set_taps(h, taps.data(), tapCount); }
        }
        stream<T>* _in = nullptr;
        qdsp_fir* h = nullptr;
    };

    class BFMDeemp : public generic_block<BFMDeemp> {
    public:
        BFMDeemp() {}
        BFMDeemp(stream<stereo_t>* in, float sampleRate, float tau) { init(in, sampleRate, tau); }
        ~BFMDeemp() {
            generic_block<BFMDeemp>::stop();
            if (h) { qdsp_deemp_destroy(h); }
        }
        void init(stream<stereo_t>* in, float sampleRate, float tau) {
            _in = in;
            _sampleRate = sampleRate;
            _tau = tau;
            rebuild();
            generic_block<BFMDeemp>::registerInput(_in);
            generic_block<BFMDeemp>::registerOutput(&out);
        }
        void setInput(stream<stereo_t>* in) { generic_block<BFMDeemp>::rebindInput(_in, in); }
        // like the reference (filter.h:117-127) the setters only change scalars: legal while the worker is in run()
        void setSampleRate(float sampleRate) { _sampleRate = sampleRate; if (h) { qdsp_deemp_set_params(h, _sampleRate, _tau); } }
        void setTau(float tau) { _tau = tau; if (h) { qdsp_deemp_set_params(h, _sampleRate, _tau); } }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            long long n;
            if (bypass) { n = qdsp_copy_d2d(out.writeDev(), _in->readDev(), (size_t)count * sizeof(stereo_t), cuStream) == 0 ? count : -1; }
            else { n = qdsp_deemp_process(h, _in->readDev(), out.writeDev(), count, cuStream); }
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        bool bypass = false;
        stream<stereo_t> out;

    private:
        void rebuild() {   // init() only: no worker thread exists yet
            if (h) { qdsp_deemp_set_params(h, _sampleRate, _tau); return; }
            h = qdsp_deemp_create(_sampleRate, _tau);
        }
        float _sampleRate = 48000.0f, _tau = 50e-6f;
        stream<stereo_t>* _in = nullptr;
        qdsp_deemp* h = nullptr;
    };
}
