// qdsp_b200/csrc/k_chanfft.cu — FFT polyphase channelizer for sm_100a: the ALGORITHMIC alternative to the
// direct-form channelizer (k_chan.cu) for the case BASELINE config 4 actually is: M channels spaced fs / M apart, decimation a
// multiple of M. The reference computes every channel as its own VFO (routing.h:47-57 -> vfo.h:19-36 -> resampling.h:99-132)
// + FloatFMDemod (demodulator.h:81-99): M x 10 241 MACs per output row. With theta_k = beta - 2 pi k / M + delta_k (beta =
// channel 0's NCO step, delta_k = what float32 rounding of the reference's phaseDelta leaves, |delta_k| < 2e-7 rad/sample):
//
//   y_k[m] = sum_t h[t] x[N] e^{j theta_k N},  N = N_s(m) + t
//          = e^{-j 2 pi k (N_s mod M) / M} e^{j delta_k N_c} * sum_{r<M} e^{-j 2 pi k r / M} ( u_m[r] + j delta_k v_m[r] ) + O((delta T / 2)^2)
//   u_m[r] = sum_j h[r + M j] x'[N_s + r + M j],   v_m[r] = sum_j (r + M j - t_c) h[r + M j] x'[...],   x'[N] = x[N] e^{j beta N}
//
// i.e. ONE pass of M column filters (J = 41 taps each) over the pre-rotated stream and two M-point FFTs per output row give
// all M channels: 2 x 10 241 MACs + 2 FFTs instead of M x 10 241 MACs (about 100x fewer flops for M = 256). The first-order
// term in delta_k is what keeps parity with the reference's float32 NCO frequencies (delta T / 2 ~ 1e-3 rad would not pass
// the 1e-4 audio bar; the second-order residue is ~5e-7).
//
// Kernel A (chanfft_poly_kernel): thread = column r, walks down the rows (M samples each) of a segment with its J taps in
// registers; 9 outputs in flight (static slot rotation: 9 group bodies of 5 rows); per sample and output 2 FFMA + 2 FADD
// (running sum R and sum of running sums W, from which sum_j j a_j = J U - W: no second tap bank). Rows arrive by TMA bulk
// copy (5 contiguous rows per stage). Writes U, W per (row, column).
// Kernel B (chanfft_fft_kernel): 16 output rows per CTA: v from (U, W), two 256-point shared-memory FFTs per row (one half of
// the CTA each), per-channel phase (64-bit turn arithmetic), first-order correction, fast_arctan2 + FM step, transposed store.
#include <math.h>
#include <new>
#include <vector>
#include "decim_common.cuh"

namespace qdsp {

constexpr int kCfM = 256;      // channels = FFT size
constexpr int kCfDR = 5;       // rows (of M samples) per output: decimation = 5 * 256 = 1280
constexpr int kCfJ = 41;       // taps per column: 41 * 256 >= 10 241 + 1
constexpr int kCfSlots = 9;    // outputs in flight: ceil(41 / 5)
constexpr int kCfStages = 3;
constexpr int kCfRB = 16;      // output rows per CTA of the FFT kernel

struct ChanFftPlan {
    int T = 0, D = 0, nch = 0;
    float* gcol_dev = nullptr;          // [2 pads][J * M]: g[t'] = h[t' - pad]
    long long* dturn_dev = nullptr;     // [M] delta_k in turns * 2^64 (signed)
    float* delta_dev = nullptr;         // [M] delta_k in rad / sample
    float2* uw = nullptr;               // [2][rows][M] scratch: U then W
    size_t uw_cap = 0;
    double max_delta = 0.0;
};

struct ChanFftArgs {
    const float2* hist;
    const float2* in;
    int H;
    long long n_in;
    PartitionDev part;
    int T, pad, seg_rows;
    long long abs0;
    uint64_t beta_step, beta_ph0;
    const float* gcol;                  // [J * M] for this pad
    float2* U;
    float2* W;
};

template <int S>
__device__ __forceinline__ void chanfft_group(const float2* __restrict__ tile, int r, const float (&gt)[kCfJ], float2 (&R)[kCfSlots],
                                              float2 (&Wc)[kCfSlots], float2& p, const float2 w256, float2& u_fin, float2& w_fin) {
    R[S] = make_float2(0.f, 0.f);          // a new output starts in this group
    Wc[S] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ir = 0; ir < kCfDR; ir++) {
        const float2 x = tile[ir * kCfM + r];
        const float2 z = cmul(x, p);
        p = cmul(p, w256);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int slot = (S - i + kCfSlots) % kCfSlots;
            const float g = gt[5 * i + ir];
            R[slot].x = fmaf(g, z.x, R[slot].x);
            R[slot].y = fmaf(g, z.y, R[slot].y);
            Wc[slot].x += R[slot].x;
            Wc[slot].y += R[slot].y;
        }
        if (ir == 0) {                     // tap row 40: the output that started 8 groups ago completes
            const int slot = (S + 1) % kCfSlots;
            const float g = gt[40];
            R[slot].x = fmaf(g, z.x, R[slot].x);
            R[slot].y = fmaf(g, z.y, R[slot].y);
            Wc[slot].x += R[slot].x;
            Wc[slot].y += R[slot].y;
            u_fin = R[slot];
            w_fin = Wc[slot];
        }
    }
}

__global__ void __launch_bounds__(kCfM, 2) chanfft_poly_kernel(const ChanFftArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t STAGE_BYTES = kCfDR * kCfM * 8u;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + kCfStages * STAGE_BYTES);
    const int r = threadIdx.x;
    const int seg = blockIdx.x, b = blockIdx.y;
    const BlkInfo bi = a.part.get(b);
    const int k0 = seg * a.seg_rows;
    if (k0 >= bi.out_count) return;
    const int nout = bi.out_count - k0 < a.seg_rows ? bi.out_count - k0 : a.seg_rows;
    const int ngroups = nout + 8;
    // group gi of the segment = rows 5 (k0 + gi) .. +4; row rho starts at sample row0 + 256 rho (relative to this call's input)
    const long long row0 = bi.in_start - a.T - a.pad + (long long)k0 * (kCfDR * kCfM);
    float gt[kCfJ];
#pragma unroll
    for (int j = 0; j < kCfJ; j++) gt[j] = a.gcol[j * kCfM + r];
    if (r == 0) {
        for (int s = 0; s < kCfStages; s++) mbar_init(&mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int gi) {     // called by every thread (uniform)
        if (gi >= ngroups) return;
        const int slot = gi % kCfStages;
        float2* dst = reinterpret_cast<float2*>(smem_raw + slot * STAGE_BYTES);
        const long long s0 = row0 + (long long)gi * (kCfDR * kCfM);
        if (s0 >= 0 && s0 + kCfDR * kCfM <= a.n_in) {
            if (r == 0) {
                mbar_arrive_expect_tx(&mbar[slot], STAGE_BYTES);
                tma_bulk_g2s(dst, a.in + s0, STAGE_BYTES, &mbar[slot]);
            }
        } else {
            VStream<float2> xs{a.hist, a.in, a.H};
#pragma unroll
            for (int ir = 0; ir < kCfDR; ir++) {
                const long long idx = s0 + ir * kCfM + r;
                dst[ir * kCfM + r] = idx < a.n_in ? xs.at(idx) : make_float2(0.f, 0.f);
            }
            __syncthreads();
            if (r == 0) mbar_arrive(&mbar[slot]);
        }
    };
    for (int gi = 0; gi < kCfStages; gi++) issue(gi);

    float2 R[kCfSlots], Wc[kCfSlots];
#pragma unroll
    for (int s = 0; s < kCfSlots; s++) R[s] = Wc[s] = make_float2(0.f, 0.f);
    const float2 w256 = phasor_from_turns(a.beta_step * (uint64_t)kCfM);
    float2 p = make_float2(1.f, 0.f);
    float2* Uo = a.U + (size_t)(bi.out_start + k0) * kCfM + r;
    float2* Wo = a.W + (size_t)(bi.out_start + k0) * kCfM + r;

#pragma unroll 1
    for (int gi = 0; gi < ngroups; gi++) {
        const int slot = gi % kCfStages;
        if ((gi & 3) == 0) {     // exact phasor re-seed every 4 groups (20 rows): x'[N] = x[N] e^{j beta N}, N absolute
            const long long N = a.abs0 + row0 + (long long)gi * (kCfDR * kCfM) + r;
            p = phasor_from_turns(a.beta_ph0 + a.beta_step * (uint64_t)N);
        }
        mbar_wait(&mbar[slot], (uint32_t)((gi / kCfStages) & 1));
        const float2* tile = reinterpret_cast<const float2*>(smem_raw + slot * STAGE_BYTES);
        float2 uf = make_float2(0.f, 0.f), wf = uf;
        switch (gi % kCfSlots) {
            case 0: chanfft_group<0>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            case 1: chanfft_group<1>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            case 2: chanfft_group<2>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            case 3: chanfft_group<3>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            case 4: chanfft_group<4>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            case 5: chanfft_group<5>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            case 6: chanfft_group<6>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            case 7: chanfft_group<7>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
            default: chanfft_group<8>(tile, r, gt, R, Wc, p, w256, uf, wf); break;
        }
        const int m = gi - 8;          // the output completed by this group's first row
        if (m >= 0 && m < nout) {
            Uo[(size_t)m * kCfM] = uf;
            Wo[(size_t)m * kCfM] = wf;
        }
        __syncthreads();               // every thread is done with the slot
        issue(gi + kCfStages);
    }
}

struct ChanFftBArgs {
    const float2* U;
    const float2* W;
    long long total_out;
    long long abs0;                 // absolute index of this call's first input sample
    int T, pad, D;
    const long long* dturn;         // [M]
    const float* delta;             // [M]
    const float* demod_in;
    float* demod_out;
    float* audio;
    long long out_stride;
    float phasor_speed;
};

// 256-point radix-2 FFT of two arrays at once: threads 0..127 butterfly array A, 128..255 array B (in place, shared memory)
__device__ __forceinline__ void fft256_pair(float2* A, float2* B, const float2* tw, int t) {
    float2* X = t < 128 ? A : B;
    const int q = t & 127;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const int half = 1 << s;
        const int grp = q >> s, pos = q & (half - 1);
        const int i0 = (grp << (s + 1)) + pos, i1 = i0 + half;
        const float2 w = tw[pos << (7 - s)];
        const float2 x0 = X[i0], x1 = X[i1];
        const float2 tt = make_float2(x1.x * w.x - x1.y * w.y, x1.x * w.y + x1.y * w.x);
        X[i0] = make_float2(x0.x + tt.x, x0.y + tt.y);
        X[i1] = make_float2(x0.x - tt.x, x0.y - tt.y);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kCfM) chanfft_fft_kernel(const ChanFftBArgs a) {
    __shared__ float2 sA[kCfM], sB[kCfM], tw[kCfM / 2];
    __shared__ float s_audio[kCfRB][kCfM + 1];
    const int t = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * kCfRB;
    if (m0 >= a.total_out) return;
    const int nrows = a.total_out - m0 < kCfRB ? (int)(a.total_out - m0) : kCfRB;
    if (t < kCfM / 2) {
        float sn, cs;
        sincospif(-2.0f * (float)t / (float)kCfM, &sn, &cs);    // e^{-j 2 pi t / M}
        tw[t] = make_float2(cs, sn);
    }
    const int rev = (int)(__brev((unsigned)t) >> 24);            // bit reversal of the 8-bit index
    const long long dturn = a.dturn[t];
    const float delta = a.delta[t];
    const float tc = 0.5f * (float)a.T;
    float prev = 0.f;
    // row m0 - 1 first (its angle is the predecessor of row m0's), unless m0 is the stream position the carried state describes
    const int first = m0 > 0 ? -1 : 0;
    if (m0 == 0) prev = a.demod_in[t];
    __syncthreads();
    for (int rr = first; rr < nrows; rr++) {
        const long long m = m0 + rr;
        const float2 u = a.U[(size_t)m * kCfM + t], w = a.W[(size_t)m * kCfM + t];
        // v[r] = sum_j (r + M j - pad - tc) a_j = (r - pad - tc) U + M (J U - W)
        const float c1 = (float)t - (float)a.pad - tc + (float)(kCfM * kCfJ);
        const float2 v = make_float2(fmaf(c1, u.x, -(float)kCfM * w.x), fmaf(c1, u.y, -(float)kCfM * w.y));
        sA[rev] = u;                                             // decimation in time: bit-reversed load, natural-order output
        sB[rev] = v;
        __syncthreads();
        fft256_pair(sA, sB, tw, t);
        const float2 Uf = sA[t], Vf = sB[t];
        const float2 Y = make_float2(fmaf(-delta, Vf.y, Uf.x), fmaf(delta, Vf.x, Uf.y));    // Uf + j delta Vf
        // channel phase: delta_k N_c - 2 pi k ((N_s - pad) mod M) / M, in 64-bit turns
        const long long Ns = a.abs0 + m * (long long)a.D - a.T;
        const unsigned long long al = (unsigned long long)((Ns - a.pad) & (kCfM - 1));
        // delta_k * N_c with N_c = N_s + tc (tc may be a half-integer): modular 64-bit turn arithmetic, exact
        const unsigned long long turns = (unsigned long long)dturn * (unsigned long long)Ns +
                                         (unsigned long long)llrint((double)dturn * (double)tc) - (((unsigned long long)t * al) << 56);
        const float2 C = phasor_from_turns(turns);
        const float2 y = cmul(Y, C);
        const float ang = fast_arctan2_ref(y.y, y.x);
        if (rr >= 0) {
            s_audio[rr][t] = fm_step_ref(ang, prev, a.phasor_speed);
            if (m == a.total_out - 1) a.demod_out[t] = ang;
        }
        prev = ang;
        __syncthreads();
    }
    // transposed store: consecutive threads write consecutive rows of one channel
    for (int idx = t; idx < kCfRB * kCfM; idx += kCfM) {
        const int k = idx / kCfRB, mr = idx - k * kCfRB;
        if (mr < nrows) a.audio[(size_t)k * a.out_stride + m0 + mr] = s_audio[mr][k];
    }
}

// ---- host side -------------------------------------------------------------------------------------
ChanFftPlan* chanfft_plan_create(const float* taps, int T, int interp, int decim, int nch, const uint64_t* nco_steps) {
    static const bool on = getenv("QDSP_CHAN_FFT") ? atoi(getenv("QDSP_CHAN_FFT")) != 0 : true;
    if (!on || interp != 1 || nch != kCfM || decim != kCfDR * kCfM || T + 1 > kCfJ * kCfM) return nullptr;
    // channel k must sit on the bin grid of channel 0: theta_k = theta_0 - 2 pi k / M + delta_k with a tiny delta_k
    std::vector<long long> dturn(kCfM);
    std::vector<float> delta(kCfM);
    double maxd = 0.0;
    for (int k = 0; k < kCfM; k++) {
        const uint64_t ideal = nco_steps[0] - ((uint64_t)k << 56);      // 2^64 / 256 = 2^56 per bin
        const long long d = (long long)(nco_steps[k] - ideal);
        const double rad = (double)d * (6.283185307179586476925286766559 / 18446744073709551616.0);
        if (fabs(rad) > 2e-6) return nullptr;                            // not a uniform comb: direct form
        dturn[k] = d;
        delta[k] = (float)rad;
        if (fabs(rad) > maxd) maxd = fabs(rad);
    }
    ChanFftPlan* p = new (std::nothrow) ChanFftPlan();
    if (!p) return nullptr;
    p->T = T;
    p->D = decim;
    p->nch = nch;
    p->max_delta = maxd;
    std::vector<float> g((size_t)2 * kCfJ * kCfM, 0.0f);
    for (int pad = 0; pad < 2; pad++)
        for (int tp = 0; tp < kCfJ * kCfM; tp++) {
            const int t = tp - pad;
            g[(size_t)pad * kCfJ * kCfM + tp] = (t >= 0 && t < T) ? taps[t] : 0.0f;
        }
    if (cudaMalloc(&p->gcol_dev, g.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(p->gcol_dev, g.data(), g.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&p->dturn_dev, kCfM * sizeof(long long)) != cudaSuccess ||
        cudaMemcpy(p->dturn_dev, dturn.data(), kCfM * sizeof(long long), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&p->delta_dev, kCfM * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(p->delta_dev, delta.data(), kCfM * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_last_error("chanfft_plan_create: device allocation failed");
        chanfft_plan_destroy(p);
        return nullptr;
    }
    return p;
}
void chanfft_plan_destroy(ChanFftPlan* p) {
    if (!p) return;
    if (p->gcol_dev) cudaFree(p->gcol_dev);
    if (p->dturn_dev) cudaFree(p->dturn_dev);
    if (p->delta_dev) cudaFree(p->delta_dev);
    if (p->uw) cudaFree(p->uw);
    delete p;
}
// usable for this batch? uniform run() partition on the decimation grid, 16-byte aligned input
bool chanfft_usable(const ChanFftPlan* p, const Partition& part, const void* in) {
    if (!p || part.view.table != nullptr || part.total_out <= 0) return false;
    if (part.view.nblocks > 1 && (part.view.block_size % p->D) != 0) return false;
    return (reinterpret_cast<uintptr_t>(in) & 15) == 0;
}

int launch_chanfft(ChanFftPlan* plan, const float2* hist, int H, const float2* in, const Partition& part, uint64_t beta_step,
                   uint64_t beta_ph0, long long abs0, float phasor_speed, const float* demod_in, float* demod_out, float* audio,
                   long long out_stride, cudaStream_t s) {
    const long long rows = part.total_out;
    const size_t need = (size_t)2 * rows * kCfM;
    if (need > plan->uw_cap) {
        if (plan->uw) cudaFree(plan->uw);
        plan->uw = nullptr;
        plan->uw_cap = 0;
        QDSP_CUDA_OK(cudaMalloc(&plan->uw, need * sizeof(float2)));
        plan->uw_cap = need;
    }
    ChanFftArgs a{};
    a.hist = hist;
    a.in = in;
    a.H = H;
    a.n_in = part.view.total;
    a.part = part.view;
    a.T = plan->T;
    a.pad = plan->T & 1;                 // blocks start on multiples of D (even): the window start parity is T's
    static const int seg_env = getenv("QDSP_CHANFFT_SEG") ? atoi(getenv("QDSP_CHANFFT_SEG")) : 0;
    a.seg_rows = seg_env > 0 ? seg_env : 64;
    a.abs0 = abs0;
    a.beta_step = beta_step;
    a.beta_ph0 = beta_ph0;
    a.gcol = plan->gcol_dev + (size_t)a.pad * kCfJ * kCfM;
    a.U = plan->uw;
    a.W = plan->uw + (size_t)rows * kCfM;
    const size_t smem = kCfStages * kCfDR * kCfM * 8 + kCfStages * 8 + 16;
    static bool attr = false;
    if (!attr) {
        QDSP_CUDA_OK(cudaFuncSetAttribute(chanfft_poly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid((part.max_out + a.seg_rows - 1) / a.seg_rows, part.view.nblocks);
    chanfft_poly_kernel<<<grid, kCfM, smem, s>>>(a);
    QDSP_LAUNCH_OK();
    ChanFftBArgs bb{};
    bb.U = a.U;
    bb.W = a.W;
    bb.total_out = rows;
    bb.abs0 = abs0;
    bb.T = plan->T;
    bb.pad = a.pad;
    bb.D = plan->D;
    bb.dturn = plan->dturn_dev;
    bb.delta = plan->delta_dev;
    bb.demod_in = demod_in;
    bb.demod_out = demod_out;
    bb.audio = audio;
    bb.out_stride = out_stride;
    bb.phasor_speed = phasor_speed;
    chanfft_fft_kernel<<<(unsigned)((rows + kCfRB - 1) / kCfRB), kCfM, 0, s>>>(bb);
    QDSP_LAUNCH_OK();
    return 0;
}

}  // namespace qdsp
