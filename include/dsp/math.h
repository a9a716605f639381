// include/dsp/math.h — Add<T>, Substract<T> (sic), Multiply<T> (reference src/dsp/math.h:6-145): two input
// streams, one output; a pair of blocks with different counts is dropped like the reference does (math.h:26-30).
// T = float, complex_t or (Add / Substract) stereo_t; Multiply<complex_t> is the complex product.
#pragma once
#include <type_traits>
#include <dsp/block.h>

namespace dsp {
    namespace detail {
        template <class SELF, class T, int OP>
        class binary_block : public generic_block<SELF> {
            using base = generic_block<SELF>;
            static_assert(std::is_same<T, float>::value || std::is_same<T, complex_t>::value ||
                              (std::is_same<T, stereo_t>::value && OP != QDSP_MATH_MUL),
                          "the reference instantiates this block for float, complex_t (and stereo_t for Add/Substract)");

        public:
            ~binary_block() { base::stop(); }
            void init(stream<T>* a, stream<T>* b) {
                _a = a;
                _b = b;
                base::registerInput(_a);
                base::registerInput(_b);
                base::registerOutput(&out);
            }
            int run() override {
                const int a_count = _a->readDevice(base::cuStream);
                if (a_count < 0) { return -1; }
                const int b_count = _b->readDevice(base::cuStream);
                if (b_count < 0) { return -1; }
                if (a_count != b_count) {
                    _a->flushDevice(base::cuStream);
                    _b->flushDevice(base::cuStream);
                    return 0;
                }
                out.acquireWriteDev(base::cuStream);
                const long long n = qdsp_math_process(OP, std::is_same<T, float>::value ? QDSP_F32 : QDSP_CF32, _a->readDev(),
                                                      _b->readDev(), out.writeDev(), a_count, base::cuStream);
                _a->flushDevice(base::cuStream);
                _b->flushDevice(base::cuStream);
                if (n < 0) { return -1; }
                if (!out.swapDevice(a_count, base::cuStream)) { return -1; }
                return a_count;
            }

            stream<T> out;

        private:
            stream<T>* _a = nullptr;
            stream<T>* _b = nullptr;
        };
    }

    template <class T>
    class Add : public detail::binary_block<Add<T>, T, QDSP_MATH_ADD> {
    public:
        Add() {}
        Add(stream<T>* a, stream<T>* b) { this->init(a, b); }
    };
    template <class T>
    class Substract : public detail::binary_block<Substract<T>, T, QDSP_MATH_SUB> {
    public:
        Substract() {}
        Substract(stream<T>* a, stream<T>* b) { this->init(a, b); }
    };
    template <class T>
    class Multiply : public detail::binary_block<Multiply<T>, T, QDSP_MATH_MUL> {
    public:
        Multiply() {}
        Multiply(stream<T>* a, stream<T>* b) { this->init(a, b); }
    };
}
