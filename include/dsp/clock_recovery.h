// include/dsp/clock_recovery.h — MMClockRecovery<T> (reference src/dsp/clock_recovery.h:68-243), T = float,
// complex_t or stereo_t. The timing loop runs on the device (sequential-exact: symbols, per-block counts and state are
// bit-identical to the reference); the block reads the count back once per run() to swap its output stream.
//
// The loop interpolates with the reference's baked MMSE table INTERP_TAPS[129][8] (src/dsp/interpolation_taps.h).
// That data file is not reproduced by this library: keep the reference's <dsp/interpolation_taps.h> reachable on the
// include path (after this directory), or define QDSP_INTERP_TAPS to a `const float (*)[8]` of your own.
// EdgeTrigClockRecovery (:7-66) is not part of this library.
#pragma once
#include <type_traits>
#include <dsp/block.h>
#if !defined(QDSP_INTERP_TAPS)
#if defined(__has_include)
#if __has_include(<dsp/interpolation_taps.h>)
#include <dsp/interpolation_taps.h>
#define QDSP_INTERP_TAPS INTERP_TAPS
#endif
#endif
#endif
#if !defined(QDSP_INTERP_TAPS)
#error "dsp/clock_recovery.h needs the reference's dsp/interpolation_taps.h on the include path (or QDSP_INTERP_TAPS)"
#endif

namespace dsp {
    template <class T>
    class MMClockRecovery : public generic_block<MMClockRecovery<T>> {
        using base = generic_block<MMClockRecovery<T>>;

    public:
        MMClockRecovery() {}
        MMClockRecovery(stream<T>* in, float omega, float gainOmega, float muGain, float omegaRelLimit) {
            init(in, omega, gainOmega, muGain, omegaRelLimit);
        }
        ~MMClockRecovery() {
            base::stop();
            if (h) { qdsp_mm_destroy(h); }
        }
        void init(stream<T>* in, float omega, float gainOmega, float muGain, float omegaRelLimit) {
            _in = in;
            if (h) { qdsp_mm_destroy(h); }
            h = qdsp_mm_create(std::is_same<T, float>::value ? QDSP_F32 : QDSP_CF32, omega, gainOmega, muGain, omegaRelLimit,
                               &QDSP_INTERP_TAPS[0][0]);
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        void setOmega(float omega, float omegaRelLimit) {
            base::tempStop();
            qdsp_mm_set_omega(h, omega, omegaRelLimit);
            base::tempStart();
        }
        void setGains(float omegaGain, float muGain) {
            base::tempStop();
            qdsp_mm_set_gains(h, omegaGain, muGain);
            base::tempStart();
        }
        void setOmegaRelLimit(float omegaRelLimit) {
            base::tempStop();
            qdsp_mm_set_omega_rel_limit(h, omegaRelLimit);
            base::tempStart();
        }
        void setInput(stream<T>* in) { base::rebindInput(_in, in); }
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const int one = count;
            // returns after the kernel: the symbol count decides how much the output stream swaps
            const long long n = qdsp_mm_process(h, _in->readDev(), out.writeDev(), count, &one, 1, 0, nullptr, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice((int)n, base::cuStream)) { return -1; }
            return count;
        }

        stream<T> out;

    private:
        stream<T>* _in = nullptr;
        qdsp_mm* h = nullptr;
    };
}
