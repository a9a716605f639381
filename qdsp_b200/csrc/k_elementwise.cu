// qdsp_b200/csrc/k_elementwise.cu — streaming element-wise kernels: NCO frequency translator,
// FM quadrature demodulator, pair-average decimator, device-side synthetic IQ, FP32-peak probe.
// All are HBM-bound: 128-bit accesses, grid-stride over a grid sized in multiples of the SM count.
#include "internal.cuh"
#include "kernels.cuh"

namespace qdsp {

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
static int stream_grid(long long work_items, int threads, int ctas_per_sm) {
    long long g = (work_items + threads - 1) / threads;
    const long long cap = (long long)sm_count() * ctas_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- FrequencyXlator: out[n] = in[n] * phase(n)  (reference processing.h:64 + VOLK rotator) -----
// phase(n) is the closed form of the reference's recursive float phasor (same float32 increment,
// no drift, unit magnitude). Each thread rotates 4 consecutive samples: one exact phasor from the
// 64-bit turn counter, three more from the host-rounded powers inc^1..inc^3.
__global__ void __launch_bounds__(256) xlator_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                    long long count, uint64_t phase0, uint64_t step, float2 inc1,
                                                    float2 inc2, float2 inc3) {
    const long long nquad = count >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nquad; q += stride) {
        const long long n = q << 2;
        const float4 a = ldg_stream128(reinterpret_cast<const float4*>(in + n));
        const float4 b = ldg_stream128(reinterpret_cast<const float4*>(in + n + 2));
        const float2 p0 = phasor_from_turns(phase0 + step * (uint64_t)n);
        const float2 p1 = cmul(p0, inc1), p2 = cmul(p0, inc2), p3 = cmul(p0, inc3);
        const float2 y0 = cmul_exact(make_float2(a.x, a.y), p0);
        const float2 y1 = cmul_exact(make_float2(a.z, a.w), p1);
        const float2 y2 = cmul_exact(make_float2(b.x, b.y), p2);
        const float2 y3 = cmul_exact(make_float2(b.z, b.w), p3);
        reinterpret_cast<float4*>(out + n)[0] = make_float4(y0.x, y0.y, y1.x, y1.y);
        reinterpret_cast<float4*>(out + n)[1] = make_float4(y2.x, y2.y, y3.x, y3.y);
    }
    // ragged tail (count % 4) — one thread
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long n = nquad << 2; n < count; n++)
            out[n] = cmul_exact(in[n], phasor_from_turns(phase0 + step * (uint64_t)n));
    }
}
// pointers that are only 8-byte aligned (a sub-window of a buffer, an odd time-shard start): one sample per thread
__global__ void __launch_bounds__(256) xlator_scalar_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                           long long count, uint64_t phase0, uint64_t step) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < count; n += stride)
        out[n] = cmul_exact(ldg_stream64(in + n), phasor_from_turns(phase0 + step * (uint64_t)n));
}
// Debug replay of the reference's recursive float32 rotator (volk_32fc_s32fc_x2_rotator_32fc_generic): one thread
// per 512-sample run starts from the host-supplied phase state of that run and repeats the float recursion
// out = in * phase; phase *= inc with every product and sum rounded separately -- bit-identical to the reference.
__global__ void __launch_bounds__(128) xlator_replay_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                           const PartitionDev part, const long long* __restrict__ run0,
                                                           const float2* __restrict__ ckpt, float2 inc, long long nruns) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    // block of run r: largest b with run0[b] <= r
    int lo = 0, hi = part.nblocks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (run0[mid] <= r) lo = mid;
        else hi = mid - 1;
    }
    const BlkInfo bi = part.get(lo);
    const long long first = bi.in_start + (r - run0[lo]) * 512;
    const long long end = bi.in_start + bi.count;
    const int n = (int)(end - first < 512 ? end - first : 512);
    float2 p = ckpt[r];
    for (int j = 0; j < n; j++) {
        out[first + j] = cmul_exact(in[first + j], p);
        p = cmul_exact(p, inc);
    }
}
int launch_xlator_replay(const float2* in, float2* out, const Partition& part, const long long* run0_dev,
                         const float2* ckpt_dev, float2 inc, long long nruns, cudaStream_t s) {
    if (nruns <= 0) return 0;
    xlator_replay_kernel<<<(unsigned)((nruns + 127) / 128), 128, 0, s>>>(in, out, part.view, run0_dev, ckpt_dev, inc, nruns);
    QDSP_LAUNCH_OK();
    return 0;
}

int launch_xlator(const float2* in, float2* out, long long count, uint64_t phase0, uint64_t step, float2 inc1,
                  float2 inc2, float2 inc3, cudaStream_t s) {
    if (count <= 0) return 0;
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) {
        xlator_scalar_kernel<<<stream_grid(count, 256, 8), 256, 0, s>>>(in, out, count, phase0, step);
        QDSP_LAUNCH_OK();
        return 0;
    }
    xlator_kernel<<<stream_grid(count / 4 + 1, 256, 8), 256, 0, s>>>(in, out, count, phase0, step, inc1, inc2, inc3);
    QDSP_LAUNCH_OK();
    return 0;
}

// ---- FloatFMDemod / FMDemod (reference demodulator.h:87-94 / 164-173) -------------------------------
// out[i] = wrap(fast_arctan2(x[i]) - fast_arctan2(x[i-1])) / phasorSpeed; x[-1]'s angle is the carried
// state. Each thread recomputes the neighbour's angle instead of exchanging it.
__global__ void __launch_bounds__(256) fmdemod_kernel(const float2* __restrict__ in, void* __restrict__ out,
                                                     long long count, float phasor_speed,
                                                     const float* __restrict__ state_in,
                                                     float* __restrict__ state_out, int stereo) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
        const float2 x = in[i];
        const float cur = fast_arctan2_ref(x.y, x.x);
        float prev;
        if (i > 0) {
            const float2 xp = in[i - 1];
            prev = fast_arctan2_ref(xp.y, xp.x);
        } else {
            prev = *state_in;
        }
        const float v = fm_step_ref(cur, prev, phasor_speed);
        if (stereo) reinterpret_cast<float2*>(out)[i] = make_float2(v, v);
        else reinterpret_cast<float*>(out)[i] = v;
        if (i == count - 1) *state_out = cur;
    }
}
int launch_fmdemod(const float2* in, void* out, long long count, float phasor_speed, const float* state_in,
                   float* state_out, int stereo, cudaStream_t s) {
    if (count <= 0) return 0;
    fmdemod_kernel<<<stream_grid(count, 256, 8), 256, 0, s>>>(in, out, count, phasor_speed, state_in, state_out,
                                                              stereo);
    QDSP_LAUNCH_OK();
    return 0;
}

// ---- PowerDecimator (reference resampling.h:220-249) -----------------------------------------------
__global__ void __launch_bounds__(256) power_decim_kernel(const float4* __restrict__ in, float2* __restrict__ out,
                                                         long long n_out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < n_out; m += stride) {
        const float4 v = ldg_stream128(in + m);
        out[m] = make_float2(__fmul_rn(__fadd_rn(v.x, v.z), 0.5f), __fmul_rn(__fadd_rn(v.y, v.w), 0.5f));
    }
}
__global__ void __launch_bounds__(256) power_decim_scalar_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                long long n_out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < n_out; m += stride) {
        const float2 a = ldg_stream64(in + 2 * m), b = ldg_stream64(in + 2 * m + 1);
        out[m] = make_float2(__fmul_rn(__fadd_rn(a.x, b.x), 0.5f), __fmul_rn(__fadd_rn(a.y, b.y), 0.5f));
    }
}
int launch_power_decim(const float2* in, float2* out, long long n_out, int copy_only, cudaStream_t s) {
    if (n_out <= 0) return 0;
    if (copy_only) {
        QDSP_CUDA_OK(cudaMemcpyAsync(out, in, (size_t)n_out * sizeof(float2), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    if ((reinterpret_cast<uintptr_t>(in) & 15) != 0) {   // 8-byte aligned input: no 128-bit loads
        power_decim_scalar_kernel<<<stream_grid(n_out, 256, 8), 256, 0, s>>>(in, out, n_out);
        QDSP_LAUNCH_OK();
        return 0;
    }
    power_decim_kernel<<<stream_grid(n_out, 256, 8), 256, 0, s>>>(reinterpret_cast<const float4*>(in), out, n_out);
    QDSP_LAUNCH_OK();
    return 0;
}

// ---- StereoFMDemod's matrix step (reference demodulator.h:261-272): doubled = pilot*pilot; amb = mpx*doubled;
// out = {mpx + amb, mpx - amb}, every product and sum rounded separately like the VOLK calls it replaces ----------
__global__ void __launch_bounds__(256) stereo_matrix_kernel(const float* __restrict__ mpx, const float* __restrict__ pilot,
                                                           float2* __restrict__ out, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
        const float m = mpx[i], p = pilot[i];
        const float amb = __fmul_rn(m, __fmul_rn(p, p));
        out[i] = make_float2(__fadd_rn(m, amb), __fsub_rn(m, amb));
    }
}
int launch_stereo_matrix(const float* mpx, const float* pilot, float2* out, long long count, cudaStream_t s) {
    if (count <= 0) return 0;
    stereo_matrix_kernel<<<stream_grid(count, 256, 8), 256, 0, s>>>(mpx, pilot, out, count);
    QDSP_LAUNCH_OK();
    return 0;
}

// ---- synthetic IQ (integer recipe identical to qdsp_b200/synth.py) -------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float2 uniform_sample(unsigned long long seed, long long n) {
    const uint64_t z = splitmix64(((uint64_t)seed << 40) ^ (uint64_t)n);
    const float re = (float)(uint32_t)(z & 0xFFFFFF) * 1.1920928955078125e-07f - 1.0f;
    const float im = (float)(uint32_t)((z >> 24) & 0xFFFFFF) * 1.1920928955078125e-07f - 1.0f;
    return make_float2(re, im);
}
__global__ void __launch_bounds__(256) synth_uniform_kernel(float2* __restrict__ out, unsigned long long seed,
                                                           long long start, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride)
        out[i] = uniform_sample(seed, start + i);
}
int launch_synth_uniform(float2* out, unsigned long long seed, long long start, long long count, cudaStream_t s) {
    if (count <= 0) return 0;
    synth_uniform_kernel<<<stream_grid(count, 256, 8), 256, 0, s>>>(out, seed, start, count);
    QDSP_LAUNCH_OK();
    return 0;
}
__device__ __forceinline__ double frac_ratio(long long n, long long f, long long fs) {
    long long r = ((n % fs) * (f % fs)) % fs;  // |n%fs * f%fs| < 2^63 for fs < 2^31
    if (r < 0) r += fs;
    return (double)r / (double)fs;
}
__global__ void __launch_bounds__(256) synth_fm_kernel(float2* __restrict__ out, long long start, long long count,
                                                      long long fs, long long fc, long long fm, double beta,
                                                      double amp, float noise_amp, unsigned long long noise_seed) {
    const double two_pi = 6.283185307179586476925286766559;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
        const long long n = start + i;
        const double ph = two_pi * frac_ratio(n, fc, fs) + beta * sin(two_pi * frac_ratio(n, fm, fs));
        double sn, cs;
        sincos(ph, &sn, &cs);
        float2 v = make_float2((float)(amp * cs), (float)(amp * sn));
        if (noise_amp != 0.0f) {
            const float2 u = uniform_sample(noise_seed, n);
            v.x = __fadd_rn(v.x, __fmul_rn(noise_amp, u.x));
            v.y = __fadd_rn(v.y, __fmul_rn(noise_amp, u.y));
        }
        out[i] = v;
    }
}
int launch_synth_fm(float2* out, long long start, long long count, long long fs, long long fc, long long fm,
                    double dev, double amp, double noise_amp, unsigned long long noise_seed, cudaStream_t s) {
    if (count <= 0) return 0;
    synth_fm_kernel<<<stream_grid(count, 256, 8), 256, 0, s>>>(out, start, count, fs, fc, fm, dev / (double)fm, amp,
                                                               (float)noise_amp, noise_seed);
    QDSP_LAUNCH_OK();
    return 0;
}

// BASELINE config 4 wideband input (qdsp_b200/synth.py cfg4_input): sum of nch FM carriers at (k-(nch-1)/2)*spacing,
// fm = 300+10k Hz, dev 5 kHz, A = amp, + noise_amp * U(seed, n). Phases from exact integer fractions, evaluated in float32
// (a synthetic test signal: the oracle consumes the very samples generated here, copied back).
__global__ void __launch_bounds__(256) synth_comb_kernel(float2* __restrict__ out, long long start, long long count,
                                                        long long fs, int nch, long long spacing, float dev, float amp,
                                                        float noise_amp, unsigned long long noise_seed) {
    const float two_pi = 6.283185307179586f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
        const long long n = start + i;
        float ar = 0.f, ai = 0.f;
        for (int k = 0; k < nch; k++) {
            // fc = (2k - (nch-1)) * spacing / 2 may be a half-integer multiple of spacing: work in units of fs * 2
            const long long fc2 = (2ll * k - (nch - 1)) * spacing;          // 2 * fc
            long long r = (((n % (2 * fs)) * (fc2 % (2 * fs))) % (2 * fs));
            if (r < 0) r += 2 * fs;
            const float pc = (float)((double)r / (double)(2 * fs));
            const long long fm = 300 + 10 * k;
            const float pm = (float)frac_ratio(n, fm, fs);
            float sm, cm_;
            sincosf(two_pi * pm, &sm, &cm_);
            float sn, cs;
            sincosf(two_pi * pc + (dev / (float)fm) * sm, &sn, &cs);
            ar += cs;
            ai += sn;
        }
        float2 v = make_float2(amp * ar, amp * ai);
        if (noise_amp != 0.0f) {
            const float2 u = uniform_sample(noise_seed, n);
            v.x += noise_amp * u.x;
            v.y += noise_amp * u.y;
        }
        out[i] = v;
    }
}
int launch_synth_comb(float2* out, long long start, long long count, long long fs, int nch, long long spacing, double dev,
                      double amp, double noise_amp, unsigned long long noise_seed, cudaStream_t s) {
    if (count <= 0) return 0;
    synth_comb_kernel<<<stream_grid(count, 256, 8), 256, 0, s>>>(out, start, count, fs, nch, spacing, (float)dev, (float)amp,
                                                                 (float)noise_amp, noise_seed);
    QDSP_LAUNCH_OK();
    return 0;
}
// BASELINE config 5 inputs (synth.py qpsk_cf32): QPSK symbols held for sps samples, rotated by freq_off rad/sample,
// slow amplitude modulation (depth, period), + sigma * U(seed, n).
__global__ void __launch_bounds__(256) synth_qpsk_kernel(float2* __restrict__ out, long long start, long long count,
                                                        unsigned long long seed, int sps, double freq_off, float sigma,
                                                        float am_depth, long long am_period) {
    const double two_pi = 6.283185307179586476925286766559;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
        const long long n = start + i;
        const uint64_t sym = splitmix64(((seed + 77ull) << 40) ^ (uint64_t)(n / sps));
        const int bits = (int)(sym & 3ull);
        double ph = 0.78539816339744830962 + 1.57079632679489661923 * bits + fmod(freq_off * (double)n, two_pi);
        double sn, cs;
        sincos(ph, &sn, &cs);
        double a = 1.0;
        if (am_depth != 0.0f) a += (double)am_depth * sin(two_pi * (double)(n % am_period) / (double)am_period);
        const float2 u = uniform_sample(seed, n);
        out[i] = make_float2((float)(a * cs) + sigma * u.x, (float)(a * sn) + sigma * u.y);
    }
}
int launch_synth_qpsk(float2* out, long long start, long long count, unsigned long long seed, int sps, double freq_off,
                      double sigma, double am_depth, long long am_period, cudaStream_t s) {
    if (count <= 0) return 0;
    synth_qpsk_kernel<<<stream_grid(count, 256, 8), 256, 0, s>>>(out, start, count, seed, sps, freq_off, (float)sigma,
                                                                 (float)am_depth, am_period > 0 ? am_period : 1);
    QDSP_LAUNCH_OK();
    return 0;
}

// ---- FP32 roofline denominator: FMA-saturation probe --------------------------------------------------
template <int PACKED>
__global__ void __launch_bounds__(512) fp32_peak_kernel(float* sink, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 1e-6f + i, i * 0.5f);
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (PACKED) {
                acc[i] = __ffma2_rn(acc[i], aa, bb);
            } else {
                acc[i].x = fmaf(acc[i].x, aa.x, bb.x);
                acc[i].y = fmaf(acc[i].y, aa.y, bb.y);
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) r += acc[i].x + acc[i].y;
    if (r == 123.456f) sink[0] = r;
}
double run_fp32_peak(int packed, int iters) {
    float* sink = nullptr;
    if (cudaMalloc(&sink, 64) != cudaSuccess) return -1.0;
    const int ctas = sm_count() * 4, threads = 512;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        if (packed) fp32_peak_kernel<1><<<ctas, threads>>>(sink, iters, 0.999f, 0.001f);
        else fp32_peak_kernel<0><<<ctas, threads>>>(sink, iters, 0.999f, 0.001f);
        count_launch();
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 32.0 * (double)iters * (double)ctas * threads;  // 32 FMA / thread / iter
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

}  // namespace qdsp
