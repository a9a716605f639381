// include/dsp/demodulator.h — FloatFMDemod and FMDemod (reference src/dsp/demodulator.h:33-187). The kernel
// evaluates the reference's fast_arctan2 phase-difference formula with its exact float sequence (not an
// atan2 of a conjugate product: SURVEY.md Q6). Also StereoFMDemod, AMDemod and SSBDemod (:189-497); the MSK/PSK
// hier-block demodulators (:499-682) are provided when <dsp/clock_recovery.h> can be included (it needs the
// reference's interpolation table, see there): define QDSP_WITH_HIER_DEMODS or include clock_recovery.h first.
#pragma once
#include <dsp/block.h>

namespace dsp {
    namespace detail {
        template <class SELF, class OUT_T, int STEREO>
        class fm_demod_base : public generic_block<SELF> {
            using base = generic_block<SELF>;

        public:
            ~fm_demod_base() {
                base::stop();
                if (h) { qdsp_fmdemod_destroy(h); }
            }
            void init(stream<complex_t>* in, float sampleRate, float deviation) {
                _in = in;
                _sampleRate = sampleRate;
                _deviation = deviation;
                rebuild();
                base::registerInput(_in);
                base::registerOutput(&out);
            }
            void setInput(stream<complex_t>* in) {
                std::lock_guard<std::mutex> lck(base::ctrlMtx);
                base::tempStop();
                base::unregisterInput(_in);
                _in = in;
                base::registerInput(_in);
                base::tempStart();
            }
            void setSampleRate(float sampleRate) {
                std::lock_guard<std::mutex> lck(base::ctrlMtx);
                base::tempStop();
                _sampleRate = sampleRate;
                rebuild();
                base::tempStart();
            }
            float getSampleRate() { return _sampleRate; }
            void setDeviation(float deviation) {
                std::lock_guard<std::mutex> lck(base::ctrlMtx);
                base::tempStop();
                _deviation = deviation;
                rebuild();
                base::tempStart();
            }
            float getDeviation() { return _deviation; }
            int run() override {
                const int count = _in->readDevice(base::cuStream);
                if (count < 0) { return -1; }
                out.acquireWriteDev(base::cuStream);
                const long long n = qdsp_fmdemod_process(h, _in->readDev(), out.writeDev(), count, base::cuStream);
                _in->flushDevice(base::cuStream);
                if (n < 0) { return -1; }
                if (!out.swapDevice(count, base::cuStream)) { return -1; }
                return count;
            }

            stream<OUT_T> out;

        private:
            void rebuild() {
                float phase = 0.0f;
                if (h) { phase = qdsp_fmdemod_get_phase(h); qdsp_fmdemod_destroy(h); }
                h = qdsp_fmdemod_create(_sampleRate, _deviation, STEREO);
                qdsp_fmdemod_set_phase(h, phase);
            }
            float _sampleRate = 1, _deviation = 1;
            stream<complex_t>* _in = nullptr;
            qdsp_fmdemod* h = nullptr;
        };
    }

    // StereoFMDemod (reference demodulator.h:189-330): FloatFMDemod -> pilot FIR<float> -> AGC -> L/R matrix. The
    // reference wires four threads through a Splitter; here one run() enqueues the four kernels back to back.
    class StereoFMDemod : public generic_block<StereoFMDemod> {
    public:
        StereoFMDemod() {}
        StereoFMDemod(stream<complex_t>* in, float sampleRate, float deviation) { init(in, sampleRate, deviation); }
        ~StereoFMDemod() {
            generic_block<StereoFMDemod>::stop();
            if (h) { qdsp_stereofm_destroy(h); }
        }
        void init(stream<complex_t>* in, float sampleRate, float deviation) {
            _in = in;
            _sampleRate = sampleRate;
            _deviation = deviation;
            if (h) { qdsp_stereofm_destroy(h); }
            h = qdsp_stereofm_create(_sampleRate, _deviation);
            generic_block<StereoFMDemod>::registerInput(_in);
            generic_block<StereoFMDemod>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<StereoFMDemod>::rebindInput(_in, in); }
        float getSampleRate() { return _sampleRate; }
        float getDeviation() { return _deviation; }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const int one = count;
            const long long n = qdsp_stereofm_process(h, _in->readDev(), out.writeDev(), count, &one, 1, 0, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<stereo_t> out;

    private:
        float _sampleRate = 1, _deviation = 1;
        stream<complex_t>* _in = nullptr;
        qdsp_stereofm* h = nullptr;
    };

    class FloatFMDemod : public detail::fm_demod_base<FloatFMDemod, float, 0> {
    public:
        FloatFMDemod() {}
        FloatFMDemod(stream<complex_t>* in, float sampleRate, float deviation) { init(in, sampleRate, deviation); }
    };

    class FMDemod : public detail::fm_demod_base<FMDemod, stereo_t, 1> {
    public:
        FMDemod() {}
        FMDemod(stream<complex_t>* in, float sampleRate, float deviation) { init(in, sampleRate, deviation); }
    };

    // AMDemod (reference demodulator.h:332-378): |x| minus its mean over the run() block
    class AMDemod : public generic_block<AMDemod> {
    public:
        AMDemod() {}
        AMDemod(stream<complex_t>* in) { init(in); }
        ~AMDemod() {
            generic_block<AMDemod>::stop();
            if (h) { qdsp_amdemod_destroy(h); }
        }
        void init(stream<complex_t>* in) {
            _in = in;
            if (!h) { h = qdsp_amdemod_create(); }
            generic_block<AMDemod>::registerInput(_in);
            generic_block<AMDemod>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<AMDemod>::rebindInput(_in, in); }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const int one = count;  // one run() call == one block of the mean (demodulator.h:364-366)
            const long long n = qdsp_amdemod_process(h, _in->readDev(), out.writeDev(), count, &one, 1, 0, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<float> out;

    private:
        stream<complex_t>* _in = nullptr;
        qdsp_amdemod* h = nullptr;
    };

    // SSBDemod (reference demodulator.h:380-497): rotate by +-pi*bandWidth/sampleRate per sample, keep the real part
    class SSBDemod : public generic_block<SSBDemod> {
    public:
        SSBDemod() {}
        SSBDemod(stream<complex_t>* in, float sampleRate, float bandWidth, int mode) { init(in, sampleRate, bandWidth, mode); }
        ~SSBDemod() {
            generic_block<SSBDemod>::stop();
            if (h) { qdsp_ssbdemod_destroy(h); }
        }
        enum { MODE_USB, MODE_LSB, MODE_DSB };
        void init(stream<complex_t>* in, float sampleRate, float bandWidth, int mode) {
            _in = in;
            _sampleRate = sampleRate;
            _bandWidth = bandWidth;
            _mode = mode;
            if (h) { qdsp_ssbdemod_destroy(h); }
            h = qdsp_ssbdemod_create(_sampleRate, _bandWidth, _mode);
            generic_block<SSBDemod>::registerInput(_in);
            generic_block<SSBDemod>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<SSBDemod>::rebindInput(_in, in); }
        void setSampleRate(float sampleRate) { _sampleRate = sampleRate; reconf(); }
        void setBandWidth(float bandWidth) { _bandWidth = bandWidth; reconf(); }
        void setMode(int mode) { _mode = mode; reconf(); }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const long long n = qdsp_ssbdemod_process(h, _in->readDev(), out.writeDev(), count, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<float> out;

    private:
        // the reference changes phaseDelta without stopping the worker (demodulator.h:425-473); same here
        void reconf() { qdsp_ssbdemod_configure(h, _sampleRate, _bandWidth, _mode); }
        int _mode = MODE_USB;
        float _sampleRate = 1, _bandWidth = 0;
        stream<complex_t>* _in = nullptr;
        qdsp_ssbdemod* h = nullptr;
    };
}

#if defined(QDSP_WITH_HIER_DEMODS) || defined(QDSP_INTERP_TAPS)
#include <dsp/clock_recovery.h>
#include <dsp/filter.h>
#include <dsp/pll.h>
#include <dsp/processing.h>
#include <dsp/window.h>

namespace dsp {
    // MSKDemod (reference demodulator.h:499-565): FloatFMDemod -> MMClockRecovery<float>
    class MSKDemod : public generic_hier_block<MSKDemod> {
    public:
        MSKDemod() {}
        MSKDemod(stream<complex_t>* input, float sampleRate, float deviation, float baudRate, float omegaGain = (0.01 * 0.01) / 4,
                 float muGain = 0.01f, float omegaRelLimit = 0.005f) {
            init(input, sampleRate, deviation, baudRate);   // (sic) the reference forwards only these four (:503-505)
        }
        void init(stream<complex_t>* input, float sampleRate, float deviation, float baudRate, float omegaGain = (0.01 * 0.01) / 4,
                  float muGain = 0.01f, float omegaRelLimit = 0.005f) {
            _sampleRate = sampleRate;
            _deviation = deviation;
            _baudRate = baudRate;
            _omegaGain = omegaGain;
            _muGain = muGain;
            _omegaRelLimit = omegaRelLimit;
            demod.init(input, _sampleRate, _deviation);
            recov.init(&demod.out, _sampleRate / _baudRate, _omegaGain, _muGain, _omegaRelLimit);
            out = &recov.out;
            generic_hier_block<MSKDemod>::registerBlock(&demod);
            generic_hier_block<MSKDemod>::registerBlock(&recov);
        }
        void setSampleRate(float sampleRate) {
            _sampleRate = sampleRate;
            demod.setSampleRate(_sampleRate);
            recov.setOmega(_sampleRate / _baudRate, _omegaRelLimit);
        }
        void setDeviation(float deviation) {
            _deviation = deviation;
            demod.setDeviation(deviation);
        }
        void setBaudRate(float baudRate, float omegaRelLimit) {
            _baudRate = baudRate;
            _omegaRelLimit = omegaRelLimit;
            recov.setOmega(_sampleRate / _baudRate, _omegaRelLimit);
        }
        void setMMGains(float omegaGain, float myGain) {
            _omegaGain = omegaGain;
            _muGain = myGain;
            recov.setGains(_omegaGain, _muGain);
        }
        void setOmegaRelLimit(float omegaRelLimit) {
            _omegaRelLimit = omegaRelLimit;
            recov.setOmegaRelLimit(_omegaRelLimit);
        }

        stream<float>* out = NULL;

    private:
        FloatFMDemod demod;
        MMClockRecovery<float> recov;
        float _sampleRate = 1, _deviation = 1, _baudRate = 1, _omegaGain = 0, _muGain = 0, _omegaRelLimit = 0;
    };

    // PSKDemod<ORDER, OFFSET> (reference demodulator.h:567-682):
    // ComplexAGC(1, 65535, agcRate) -> FIR<complex_t>(RRCTaps) -> CostasLoop<ORDER> [-> DelayImag] -> MMClockRecovery<complex_t>
    template <int ORDER, bool OFFSET>
    class PSKDemod : public generic_hier_block<PSKDemod<ORDER, OFFSET>> {
        using hier = generic_hier_block<PSKDemod<ORDER, OFFSET>>;

    public:
        PSKDemod() {}
        PSKDemod(stream<complex_t>* input, float sampleRate, float baudRate, int RRCTapCount = 32, float RRCAlpha = 0.32f,
                 float agcRate = 10e-4, float costasLoopBw = 0.004f, float omegaGain = (0.01 * 0.01) / 4, float muGain = 0.01f,
                 float omegaRelLimit = 0.005f) {
            init(input, sampleRate, baudRate, RRCTapCount, RRCAlpha, agcRate, costasLoopBw, omegaGain, muGain, omegaRelLimit);
        }
        void init(stream<complex_t>* input, float sampleRate, float baudRate, int RRCTapCount = 32, float RRCAlpha = 0.32f,
                  float agcRate = 10e-4, float costasLoopBw = 0.004f, float omegaGain = (0.01 * 0.01) / 4, float muGain = 0.01f,
                  float omegaRelLimit = 0.005f) {
            _RRCTapCount = RRCTapCount;
            _RRCAlpha = RRCAlpha;
            _sampleRate = sampleRate;
            _agcRate = agcRate;
            _costasLoopBw = costasLoopBw;
            _baudRate = baudRate;
            _omegaGain = omegaGain;
            _muGain = muGain;
            _omegaRelLimit = omegaRelLimit;
            agc.init(input, 1.0f, 65535, _agcRate);
            taps.init(_RRCTapCount, _sampleRate, _baudRate, _RRCAlpha);
            rrc.init(&agc.out, &taps);
            demod.init(&rrc.out, _costasLoopBw);
            hier::registerBlock(&agc);
            hier::registerBlock(&rrc);
            hier::registerBlock(&demod);
            if (OFFSET) {
                delay.init(&demod.out);
                recov.init(&delay.out, _sampleRate / _baudRate, _omegaGain, _muGain, _omegaRelLimit);
                hier::registerBlock(&delay);
            } else {
                recov.init(&demod.out, _sampleRate / _baudRate, _omegaGain, _muGain, _omegaRelLimit);
            }
            hier::registerBlock(&recov);
            out = &recov.out;
        }
        void setInput(stream<complex_t>* input) { agc.setInput(input); }
        void setSampleRate(float sampleRate) {
            _sampleRate = sampleRate;
            taps.setSampleRate(_sampleRate);
            rrc.updateWindow(&taps);
            recov.setOmega(_sampleRate / _baudRate, _omegaRelLimit);
        }
        void setBaudRate(float baudRate) {
            _baudRate = baudRate;
            taps.setBaudRate(_baudRate);
            rrc.updateWindow(&taps);
            recov.setOmega(_sampleRate / _baudRate, _omegaRelLimit);
        }
        void setRRCParams(int RRCTapCount, float RRCAlpha) {
            _RRCTapCount = RRCTapCount;
            _RRCAlpha = RRCAlpha;
            taps.setTapCount(_RRCTapCount);
            taps.setAlpha(RRCAlpha);
            rrc.updateWindow(&taps);
        }
        void setAgcRate(float agcRate) {
            _agcRate = agcRate;
            agc.setRate(_agcRate);
        }
        void setCostasLoopBw(float costasLoopBw) {
            _costasLoopBw = costasLoopBw;
            demod.setLoopBandwidth(_costasLoopBw);
        }
        void setMMGains(float omegaGain, float myGain) {
            _omegaGain = omegaGain;
            _muGain = myGain;
            recov.setGains(_omegaGain, _muGain);
        }
        void setOmegaRelLimit(float omegaRelLimit) {
            _omegaRelLimit = omegaRelLimit;
            recov.setOmegaRelLimit(_omegaRelLimit);
        }

        stream<complex_t>* out = NULL;

    private:
        dsp::ComplexAGC agc;
        dsp::RRCTaps taps;
        dsp::FIR<dsp::complex_t> rrc;
        CostasLoop<ORDER> demod;
        DelayImag delay;
        MMClockRecovery<dsp::complex_t> recov;
        int _RRCTapCount = 32;
        float _RRCAlpha = 0.32f, _sampleRate = 1, _agcRate = 1e-3f, _baudRate = 1, _costasLoopBw = 0.004f;
        float _omegaGain = 0, _muGain = 0, _omegaRelLimit = 0;
    };
}
#endif
