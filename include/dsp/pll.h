// include/dsp/pll.h — CostasLoop<ORDER>, ORDER in {2, 4, 8} (reference src/dsp/pll.h:7-117). run() uses the
// chunked block-parallel scan (warm-up + 2*pi/ORDER ambiguity stitching); setChunking(0, 0) selects the strictly
// sequential walk.
#pragma once
#include <dsp/block.h>

namespace dsp {
    template <int ORDER>
    class CostasLoop : public generic_block<CostasLoop<ORDER>> {
        using base = generic_block<CostasLoop<ORDER>>;
        static_assert(ORDER == 2 || ORDER == 4 || ORDER == 8, "CostasLoop order must be 2, 4 or 8");

    public:
        CostasLoop() {}
        CostasLoop(stream<complex_t>* in, float loopBandwidth) { init(in, loopBandwidth); }
        ~CostasLoop() {
            base::stop();
            if (h) { qdsp_costas_destroy(h); }
        }
        void init(stream<complex_t>* in, float loopBandwidth) {
            _in = in;
            _loopBandwidth = loopBandwidth;
            rebuild();
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) {
            base::tempStop();
            base::unregisterInput(_in);
            _in = in;
            base::registerInput(_in);
            base::tempStart();
        }
        void setLoopBandwidth(float loopBandwidth) {
            base::tempStop();
            _loopBandwidth = loopBandwidth;
            rebuild();
            base::tempStart();
        }
        void setChunking(int chunk, int warmup) {
            _chunk = chunk;
            _warmup = warmup;
            qdsp_costas_set_chunking(h, chunk, warmup);
        }
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const long long n = qdsp_costas_process(h, _in->readDev(), out.writeDev(), count, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, base::cuStream)) { return -1; }
            return count;
        }

        stream<complex_t> out;

    private:
        void rebuild() {
            float st[4] = {0.0f, 0.0f, 1.0f, 0.0f};
            if (h) { qdsp_costas_get_state(h, st); qdsp_costas_destroy(h); }
            h = qdsp_costas_create(ORDER, _loopBandwidth);
            qdsp_costas_set_state(h, st);
            qdsp_costas_set_chunking(h, _chunk, _warmup);
        }
        float _loopBandwidth = 1.0f;
        int _chunk = 4096, _warmup = 4096;
        stream<complex_t>* _in = nullptr;
        qdsp_costas* h = nullptr;
    };
}
