// qdsp_b200/csrc/k_generic.cu — shape-agnostic polyphase kernels (any interp/decim/tap count, float or
// cf32 elements, optional NCO prologue and FM-demod epilogue). These are the always-correct variants
// the dispatcher falls back to when no specialised sm_100a kernel (k_decim.cu, k_fir.cu) covers the
// shape; they also serve as the on-device cross-check of the specialised kernels.
//
// Math (reference src/dsp/resampling.h:99-132 and filter.h:51-74, restated):
//   block b, output k:  i = k*D, p = i % I, n = i / I
//   y_b[k] = sum_{t<TPP} phases[p][t] * x[S_b + n + t - TPP + lead]
//   phases[p][t] = taps[t*I + (I-1-p)] (0 beyond tapCount)      -- buildTapPhases, resampling.h:137-166
//   lead = 0 for PolyphaseResampler, 1 for FIR (filter.h:65 reads &buffer[i+1]).
#include "internal.cuh"
#include "kernels.cuh"

namespace qdsp {

template <typename T, bool ROTATE>
__device__ __forceinline__ T generic_output(const VStream<T>& xs, const float* __restrict__ phases, int TPP,
                                            int interp, int decim, int lead, long long in_start, int k,
                                            uint64_t nco_phase0, uint64_t nco_step) {
    const long long i = (long long)k * decim;
    const int p = (int)(i % interp);
    const long long n = i / interp;
    const float* __restrict__ h = phases + (size_t)p * TPP;
    const long long base = in_start + n - TPP + lead;
    T acc0 = Elem<T>::zero(), acc1 = Elem<T>::zero();
    int t = 0;
    if (base >= 0 && !ROTATE) {
        const T* __restrict__ x = xs.in + base;
        for (; t + 1 < TPP; t += 2) {
            acc0 = Elem<T>::mac(x[t], h[t], acc0);
            acc1 = Elem<T>::mac(x[t + 1], h[t + 1], acc1);
        }
        if (t < TPP) acc0 = Elem<T>::mac(x[t], h[t], acc0);
    } else {
        for (; t < TPP; t++) {
            T v = xs.at(base + t);
            if constexpr (ROTATE) {
                // closed-form NCO: absolute sample index relative to this call = base + t
                const float2 ph = phasor_from_turns(nco_phase0 + nco_step * (uint64_t)(base + t));
                v = cmul_exact(v, ph);
            }
            // same even/odd accumulator split as the fast path: results do not depend on the path
            if (t & 1) acc1 = Elem<T>::mac(v, h[t], acc1);
            else acc0 = Elem<T>::mac(v, h[t], acc0);
        }
    }
    return Elem<T>::add(acc0, acc1);
}

// One thread per output; grid = (tiles, blocks of the partition).
template <typename T>
__global__ void __launch_bounds__(256) generic_resamp_kernel(VStream<T> xs, const float* __restrict__ phases, int TPP,
                                                            int lead, PartitionDev part, T* __restrict__ out) {
    const BlkInfo bi = part.get(blockIdx.y);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bi.out_count) return;
    out[bi.out_start + k] =
        generic_output<T, false>(xs, phases, TPP, part.interp, part.decim, lead, bi.in_start, k, 0, 0);
}

// Schedule export: the (phase, index) pair each output uses, computed with the kernels' own integer
// arithmetic (parity of indices vs resampling.h:121-123).
__global__ void schedule_kernel(PartitionDev part, int* __restrict__ phase, long long* __restrict__ index) {
    const BlkInfo bi = part.get(blockIdx.y);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bi.out_count) return;
    const long long i = (long long)k * part.decim;
    phase[bi.out_start + k] = (int)(i % part.interp);
    index[bi.out_start + k] = i / part.interp;
}

// Fused xlate -> resample -> FM demod, generic shape, nch channels (blockIdx.z). 255 outputs + 1
// leading "previous" output per CTA; the previous output of a block's first tile is the last output
// of the previous block (its window sits at that block's own offset), and the very first output of
// the call differentiates against the carried demod phase (reference demodulator.h:88-93,104).
__global__ void __launch_bounds__(256)
generic_vfofm_kernel(VStream<float2> xs, const float* __restrict__ phases, int TPP, PartitionDev part,
                     const NcoDev* __restrict__ nco, long long abs0, float phasor_speed,
                     const float* __restrict__ demod_in, float* __restrict__ demod_out, float* __restrict__ audio,
                     float2* __restrict__ iq, long long out_stride) {
    __shared__ float s_ang[257];
    const int b = blockIdx.y;
    const int ch = blockIdx.z;
    const BlkInfo bi = part.get(b);
    const int k0 = blockIdx.x * 255;
    if (k0 >= bi.out_count) return;
    const uint64_t nco_step = nco[ch].step;
    const uint64_t nco_phase0 = nco[ch].init + nco_step * (uint64_t)abs0;
    const int k = k0 - 1 + (int)threadIdx.x;  // thread 0 computes the previous output
    float ang = 0.0f;
    bool have = false;
    float2 y = make_float2(0.f, 0.f);
    if (k >= 0 && k < bi.out_count) {
        y = generic_output<float2, true>(xs, phases, TPP, part.interp, part.decim, 0, bi.in_start, k, nco_phase0,
                                         nco_step);
        have = true;
    } else if (k < 0) {
        // previous output lives in an earlier block (skip empty ones) or in the carried state
        int pb = b - 1;
        BlkInfo pbi{};
        while (pb >= 0) {
            pbi = part.get(pb);
            if (pbi.out_count > 0) break;
            pb--;
        }
        if (pb >= 0) {
            y = generic_output<float2, true>(xs, phases, TPP, part.interp, part.decim, 0, pbi.in_start,
                                             pbi.out_count - 1, nco_phase0, nco_step);
            ang = fast_arctan2_ref(y.y, y.x);
        } else {
            ang = demod_in[ch];
        }
    }
    if (have) {
        ang = fast_arctan2_ref(y.y, y.x);
        if (threadIdx.x > 0 && iq) iq[ch * out_stride + bi.out_start + k] = y;
    }
    s_ang[threadIdx.x] = ang;
    __syncthreads();
    if (threadIdx.x > 0 && have) {
        audio[ch * out_stride + bi.out_start + k] = fm_step_ref(ang, s_ang[threadIdx.x - 1], phasor_speed);
        // the last output of the whole call leaves its phase behind for the next call
        if (bi.out_start + k == part.total_out - 1) demod_out[ch] = ang;
    }
}

// ---- host launchers --------------------------------------------------------------------------
template <typename T>
int launch_generic_resamp(const T* hist, int H, const T* in, const float* phases_dev, int TPP, int lead,
                          const Partition& part, T* out, cudaStream_t s) {
    if (part.view.nblocks == 0 || part.max_out == 0) return 0;
    VStream<T> xs{hist, in, H};
    dim3 grid((part.max_out + 255) / 256, part.view.nblocks);
    generic_resamp_kernel<T><<<grid, 256, 0, s>>>(xs, phases_dev, TPP, lead, part.view, out);
    QDSP_LAUNCH_OK();
    return 0;
}
template int launch_generic_resamp<float>(const float*, int, const float*, const float*, int, int, const Partition&,
                                          float*, cudaStream_t);
template int launch_generic_resamp<float2>(const float2*, int, const float2*, const float*, int, int,
                                           const Partition&, float2*, cudaStream_t);

int launch_schedule(const Partition& part, int* phase_dev, long long* index_dev, cudaStream_t s) {
    if (part.view.nblocks == 0 || part.max_out == 0) return 0;
    dim3 grid((part.max_out + 255) / 256, part.view.nblocks);
    schedule_kernel<<<grid, 256, 0, s>>>(part.view, phase_dev, index_dev);
    QDSP_LAUNCH_OK();
    return 0;
}

int launch_generic_vfofm(const float2* hist, int H, const float2* in, const float* phases_dev, int TPP,
                         const Partition& part, const NcoDev* nco, long long abs0, int nch, float phasor_speed,
                         const float* demod_in, float* demod_out, float* audio, float2* iq, long long out_stride,
                         cudaStream_t s) {
    if (part.view.nblocks == 0 || part.max_out == 0) return 0;
    VStream<float2> xs{hist, in, H};
    dim3 grid((part.max_out + 254) / 255, part.view.nblocks, nch);
    generic_vfofm_kernel<<<grid, 256, 0, s>>>(xs, phases_dev, TPP, part.view, nco, abs0, phasor_speed, demod_in,
                                              demod_out, audio, iq, out_stride);
    QDSP_LAUNCH_OK();
    return 0;
}

}  // namespace qdsp
