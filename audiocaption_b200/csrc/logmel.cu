// Fused STFT -> |X|^2 -> mel filterbank -> 10*log10 kernel (fp32).
//
// Replaces torchaudio MelSpectrogram + AmplitudeToDB as called by the reference at
// captioning/models/hf_wrapper.py:269-279,292-293 and captioning/models/cnn_encoder.py:338-350,418-419.
//
// One CTA = 32 consecutive frames of one clip.  The waveform segment the 32 frames cover is
// staged once in shared memory (reflect padding resolved while loading, coalesced reads); each
// warp then transforms 4 frames.  The N/2-point complex FFT of the even/odd packed, windowed frame lives in REGISTERS:
// lane l holds z[l + 32 j] (j < R = N/64), runs an R-point FFT over j on its own registers, applies the W^(l q) twiddles,
// and the remaining 32-point FFT over the LANES is five radix-2 stages of warp shuffles; the real-FFT unpacking pairs
// bin k with bin N/2 - k through two more shuffles.  Shared memory only sees the 257 (513) power values for the banded
// mel accumulation (only the non-zero band of each filter, detected from the module's own `fb` buffer at create time)
// and the dB tile.  (The first version ran radix-4 Stockham passes through shared memory: 5 round trips per frame with
// up to 8-way bank conflicts, 0.19 ms per 64 clips; this one 0.1x ms.)  The 64x32 result tile is written back with frames
// contiguous (128 B rows) in the [B, n_mels, T] layout the reference produces.
//
// Algorithmic HBM traffic: 4*n_samples (read) + 4*n_mels*T (write) bytes per clip.
#include <math.h>

#include <algorithm>

#include "common.cuh"

struct ac_frontend {
    int n_fft, hop, n_freqs, n_mels;
    float* window_dev;     // [n_fft]
    float2* twiddle_dev;   // [n_fft/2 + 1]  e^{-2 pi i k / n_fft}
    float* fbw_dev;        // band-compact filter weights
    int* band_dev;         // [n_mels][3] = {first bin, n bins, offset into fbw}
    int fbw_len;           // floats in fbw_dev
};

namespace ac {

constexpr int kFramesPerCta = 32;
constexpr int kMelThreads = 256;
constexpr int kMelWarps = kMelThreads / 32;
constexpr int kMaxMels = 64;

__host__ __device__ constexpr int bitrev_c(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// In-register radix-2 DIF FFT of size R (8 or 16): natural order in, position p holds output bitrev(p).
template <int R>
__device__ __forceinline__ void fft_regs(float (&re)[R], float (&im)[R]) {
    // e^{-2 pi i t / 16}, t = 0..7
    constexpr float C16[8] = {1.f, 0.9238795325112867f, 0.7071067811865476f, 0.3826834323650898f,
                              0.f, -0.3826834323650898f, -0.7071067811865476f, -0.9238795325112867f};
    constexpr float S16[8] = {0.f, -0.3826834323650898f, -0.7071067811865476f, -0.9238795325112867f,
                              -1.f, -0.9238795325112867f, -0.7071067811865476f, -0.3826834323650898f};
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int b = 0; b < R; b += 2 * half) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const int t = i * (16 / (2 * half));
                const float ar = re[b + i], ai = im[b + i], cr = re[b + i + half], ci = im[b + i + half];
                re[b + i] = ar + cr; im[b + i] = ai + ci;
                const float dr = ar - cr, di = ai - ci;
                if (t == 0) { re[b + i + half] = dr; im[b + i + half] = di; }
                else if (t == 4) { re[b + i + half] = di; im[b + i + half] = -dr; }          // * (-i)
                else {
                    re[b + i + half] = dr * C16[t] - di * S16[t];
                    im[b + i + half] = dr * S16[t] + di * C16[t];
                }
            }
        }
    }
}

template <int NFFT>
__global__ void __launch_bounds__(kMelThreads, NFFT == 512 ? 4 : 2)
logmel_kernel(const float* __restrict__ wav, int n_samples, int n_frames, int hop, int n_mels,
              const float* __restrict__ window, const float2* __restrict__ twiddle,
              const float* __restrict__ fbw, int fbw_len, const int* __restrict__ band,
              float* __restrict__ out, float* __restrict__ gmax) {
    constexpr int M = NFFT / 2;           // complex FFT size
    constexpr int R = M / 32;             // points per lane
    constexpr int LOGR = R == 8 ? 3 : 4;
    static_assert(R == 8 || R == 16, "n_fft must be 512 or 1024");
    constexpr int PSTRIDE = M + M / 32 + 8;       // power buffer per warp: bin k lives at k + (k >> 5) (conflict-free)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg_len = (kFramesPerCta - 1) * hop + NFFT;
    float* s_seg = reinterpret_cast<float*>(smem_raw);                       // [seg_len] (+pad to 4)
    float2* s_tw = reinterpret_cast<float2*>(s_seg + ((seg_len + 3) & ~3));  // [M+1]  e^{-2 pi i k / NFFT}
    float* s_pow = reinterpret_cast<float*>(s_tw + (M + 2));                 // [warps][PSTRIDE]
    float* s_win = s_pow + kMelWarps * PSTRIDE;                              // [NFFT]
    float* s_tile = s_win + NFFT;                                            // [kMaxMels][33]
    float* s_fbw = s_tile + kMaxMels * 33;                                   // [fbw_len] band-compact filter weights
    __shared__ float s_red[kMelWarps];

    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kFramesPerCta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* w = wav + (size_t)b * n_samples;

    // ---- stage the segment (reflect padding as torch.stft(center=True, pad_mode="reflect"))
    const int seg_start = t0 * hop - NFFT / 2;
    for (int i = tid; i < seg_len; i += kMelThreads) {
        int idx = seg_start + i;
        if (idx < 0) idx = -idx;
        if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
        idx = min(max(idx, 0), n_samples - 1);   // only reachable for frames past n_frames
        s_seg[i] = __ldg(w + idx);
    }
    for (int i = tid; i <= M; i += kMelThreads) s_tw[i] = twiddle[i];
    for (int i = tid; i < NFFT; i += kMelThreads) s_win[i] = window[i];
    for (int i = tid; i < fbw_len; i += kMelThreads) s_fbw[i] = fbw[i];
    __syncthreads();

    // twid(m) = e^{-2 pi i m / NFFT}, 0 <= m <= NFFT (the table covers m <= M)
    auto twid = [&](int m) {
        const float2 t = s_tw[m > M ? m - M : m];
        return m > M ? make_float2(-t.x, -t.y) : t;
    };
    // per-lane constants: lane-FFT twiddles W_{2d}^(lane mod d), the output index m = bitrev5(lane) of the lane FFT,
    // and the lane that holds bin group (32 - m) mod 32 (partner of this lane's q = 0 bin in the real-FFT unpacking)
    float2 tw_d[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int d = 16 >> s;
        tw_d[s] = twid((lane & (d - 1)) * (NFFT / (2 * d)));
    }
    const int m_out = (int)(__brev((unsigned)lane) >> 27);
    const int lane_q0 = (int)(__brev((unsigned)((32 - m_out) & 31)) >> 27);

    float* pw = s_pow + warp * PSTRIDE;
    float local_max = -INFINITY;
    // this lane's two mel filters (lane, lane + 32): first bin, length, offset into the compact weights
    const bool has0 = lane < n_mels, has1 = lane + 32 < n_mels;
    const int lo0 = has0 ? band[3 * lane] : 0, nb0 = has0 ? band[3 * lane + 1] : 0, of0 = has0 ? band[3 * lane + 2] : 0;
    const int lo1 = has1 ? band[3 * (lane + 32)] : 0, nb1 = has1 ? band[3 * (lane + 32) + 1] : 0, of1 = has1 ? band[3 * (lane + 32) + 2] : 0;

    for (int f = warp; f < kFramesPerCta; f += kMelWarps) {
        const int t = t0 + f;
        if (t >= n_frames) break;   // warp-uniform
        const float* x = s_seg + f * hop;
        // windowed frame packed as z[n] = x[2n] + i x[2n+1]; this lane: n = lane + 32 j
        float re[R], im[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int n = lane + 32 * j;
            const float2 v = *reinterpret_cast<const float2*>(x + 2 * n);
            const float2 wv = *reinterpret_cast<const float2*>(s_win + 2 * n);
            re[j] = v.x * wv.x; im[j] = v.y * wv.y;
        }
        fft_regs<R>(re, im);                                  // position p: Y_lane[q], q = bitrev(p)
#pragma unroll
        for (int p = 1; p < R; ++p) {                         // * W_M^(lane q)
            const int q = bitrev_c(p, LOGR);
            const float2 tw = twid(2 * lane * q);
            const float a = re[p], c = im[p];
            re[p] = a * tw.x - c * tw.y; im[p] = a * tw.y + c * tw.x;
        }
        // 32-point DIF FFT over the lanes, every register position at once.  Lower lane of a pair: u + v; upper lane:
        // (u - v) W.  Both as (other + sgn * mine) * W' with W' = 1 on the lower lanes: no branch, no wasted half.
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int d = 16 >> s;
            const bool upper = (lane & d) != 0;
            const float sgn = upper ? -1.0f : 1.0f;
            const float wx = upper ? tw_d[s].x : 1.0f, wy = upper ? tw_d[s].y : 0.0f;
#pragma unroll
            for (int p = 0; p < R; ++p) {
                const float a = fmaf(sgn, re[p], __shfl_xor_sync(0xffffffffu, re[p], d));
                const float c = fmaf(sgn, im[p], __shfl_xor_sync(0xffffffffu, im[p], d));
                re[p] = a * wx - c * wy; im[p] = a * wy + c * wx;
            }
        }
        // position p of this lane now holds Z[k], k = R * m_out + bitrev(p).  Real-FFT unpacking against Z[M - k]:
        // q >= 1: lane ^ 31 holds group 31 - m_out, position bitrev(R - q);  q = 0: lane_q0, position 0.
        const float z0r = re[0], z0i = im[0];
#pragma unroll
        for (int p = 0; p < R; ++p) {
            const int q = bitrev_c(p, LOGR);
            const int pp = q == 0 ? 0 : bitrev_c(R - q, LOGR);
            const int src = q == 0 ? lane_q0 : (lane ^ 31);
            const float br = __shfl_sync(0xffffffffu, re[pp], src);
            const float bi = __shfl_sync(0xffffffffu, im[pp], src);
            const int k = R * m_out + q;
            const float ex = 0.5f * (re[p] + br), ey = 0.5f * (im[p] - bi);
            const float ox = 0.5f * (im[p] + bi), oy = -0.5f * (re[p] - br);
            const float2 tw = s_tw[k];
            const float xr = ex + (ox * tw.x - oy * tw.y);
            const float xi = ey + (ox * tw.y + oy * tw.x);
            pw[k + (k >> 5)] = xr * xr + xi * xi;
        }
        if (m_out == 0) pw[M + (M >> 5)] = (z0r - z0i) * (z0r - z0i);     // Nyquist bin
        __syncwarp();
        // banded mel filters + dB: both filters of the lane advance in one loop (two independent chains)
        {
            float a0 = 0.0f, a1 = 0.0f;
            const int nmax = max(nb0, nb1);
            for (int j = 0; j < nmax; ++j) {
                const int k0 = lo0 + j, k1 = lo1 + j;
                if (j < nb0) a0 = fmaf(pw[k0 + (k0 >> 5)], s_fbw[of0 + j], a0);
                if (j < nb1) a1 = fmaf(pw[k1 + (k1 >> 5)], s_fbw[of1 + j], a1);
            }
            if (has0) {
                const float db = 10.0f * log10f(fmaxf(a0, 1e-10f));
                s_tile[lane * 33 + f] = db;
                local_max = fmaxf(local_max, db);
            }
            if (has1) {
                const float db = 10.0f * log10f(fmaxf(a1, 1e-10f));
                s_tile[(lane + 32) * 33 + f] = db;
                local_max = fmaxf(local_max, db);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // coalesced store: frames contiguous
    const int nf = min(kFramesPerCta, n_frames - t0);
    float* o = out + (size_t)b * n_mels * n_frames + t0;
    for (int m = warp; m < n_mels; m += kMelWarps)
        if (lane < nf) o[(size_t)m * n_frames + lane] = s_tile[m * 33 + lane];

    if (gmax != nullptr) {
        local_max = warp_max(local_max);
        if (lane == 0) s_red[warp] = local_max;
        __syncthreads();
        if (warp == 0) {
            float v = lane < kMelWarps ? s_red[lane] : -INFINITY;
            v = warp_max(v);
            if (lane == 0 && v > -INFINITY) atomic_max_float(gmax, v);
        }
    }
}

template <int NFFT>
static size_t logmel_smem_bytes(int hop, int fbw_len) {
    constexpr int M = NFFT / 2;
    int seg_len = (kFramesPerCta - 1) * hop + NFFT;
    size_t fl = ((seg_len + 3) & ~3) + 2 * (M + 2) + kMelWarps * (M + M / 32 + 8) + NFFT + kMaxMels * 33 + fbw_len;
    return fl * sizeof(float);
}

__global__ void fill_kernel(float* p, float v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void db_clamp_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ gmax, float top_db) {
    const float floor_v = *gmax - top_db;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] = fmaxf(x[i], floor_v);
}

}  // namespace ac

extern "C" {

int ac_frontend_create(const float* window_host, int n_fft, int hop, const float* fb_host,
                       int n_freqs, int n_mels, ac_frontend_t** out) {
    AC_REQUIRE(out && window_host && fb_host, "ac_frontend_create: null argument");
    AC_REQUIRE(n_fft == 512 || n_fft == 1024, "ac_frontend_create: n_fft must be 512 or 1024 (got %d)", n_fft);
    AC_REQUIRE(n_freqs == n_fft / 2 + 1, "ac_frontend_create: n_freqs %d != n_fft/2+1", n_freqs);
    AC_REQUIRE(n_mels >= 1 && n_mels <= ac::kMaxMels, "ac_frontend_create: n_mels %d not in [1,64]", n_mels);
    AC_REQUIRE(hop >= 1 && hop <= n_fft && hop % 2 == 0, "ac_frontend_create: bad hop %d", hop);
    ac_frontend_t* fe = new ac_frontend_t();
    fe->n_fft = n_fft; fe->hop = hop; fe->n_freqs = n_freqs; fe->n_mels = n_mels;
    // non-zero band of every mel filter, taken from the module's own filterbank buffer
    std::vector<int> band(3 * n_mels);
    std::vector<float> fbw;
    for (int m = 0; m < n_mels; ++m) {
        int lo = n_freqs, hi = -1;
        for (int k = 0; k < n_freqs; ++k)
            if (fb_host[(size_t)k * n_mels + m] != 0.0f) { lo = k < lo ? k : lo; hi = k; }
        int n = hi >= lo ? hi - lo + 1 : 0;
        band[3 * m] = n ? lo : 0; band[3 * m + 1] = n; band[3 * m + 2] = (int)fbw.size();
        for (int j = 0; j < n; ++j) fbw.push_back(fb_host[(size_t)(lo + j) * n_mels + m]);
    }
    if (fbw.empty()) fbw.push_back(0.0f);
    std::vector<float2> tw(n_fft / 2 + 1);
    for (int k = 0; k <= n_fft / 2; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)n_fft;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    AC_CUDA(cudaMalloc(&fe->window_dev, n_fft * sizeof(float)));
    AC_CUDA(cudaMalloc(&fe->twiddle_dev, tw.size() * sizeof(float2)));
    AC_CUDA(cudaMalloc(&fe->fbw_dev, fbw.size() * sizeof(float)));
    AC_CUDA(cudaMalloc(&fe->band_dev, band.size() * sizeof(int)));
    AC_CUDA(cudaMemcpy(fe->window_dev, window_host, n_fft * sizeof(float), cudaMemcpyHostToDevice));
    AC_CUDA(cudaMemcpy(fe->twiddle_dev, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    AC_CUDA(cudaMemcpy(fe->fbw_dev, fbw.data(), fbw.size() * sizeof(float), cudaMemcpyHostToDevice));
    AC_CUDA(cudaMemcpy(fe->band_dev, band.data(), band.size() * sizeof(int), cudaMemcpyHostToDevice));
    fe->fbw_len = (int)fbw.size();
    AC_CUDA(cudaFuncSetAttribute(ac::logmel_kernel<512>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    AC_CUDA(cudaFuncSetAttribute(ac::logmel_kernel<1024>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    AC_CUDA(cudaFuncSetAttribute(ac::logmel_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ac::logmel_smem_bytes<512>(512, fe->fbw_len)));
    AC_CUDA(cudaFuncSetAttribute(ac::logmel_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ac::logmel_smem_bytes<1024>(1024, fe->fbw_len)));
    *out = fe;
    return AC_OK;
}

void ac_frontend_destroy(ac_frontend_t* fe) {
    if (!fe) return;
    cudaFree(fe->window_dev); cudaFree(fe->twiddle_dev); cudaFree(fe->fbw_dev); cudaFree(fe->band_dev);
    delete fe;
}

int ac_frontend_num_frames(const ac_frontend_t* fe, int n_samples) { return 1 + n_samples / fe->hop; }

int ac_logmel_fwd(const ac_frontend_t* fe, const float* wav_dev, int batch, int n_samples,
                  float* lms_dev, float* gmax_dev, void* stream) {
    AC_REQUIRE(fe, "ac_logmel_fwd: null front-end");
    AC_REQUIRE(batch >= 0 && batch <= 65535, "ac_logmel_fwd: batch %d out of range", batch);
    AC_REQUIRE(n_samples > fe->n_fft / 2, "ac_logmel_fwd: n_samples %d too short for reflect padding of %d",
               n_samples, fe->n_fft / 2);
    if (batch == 0) return AC_OK;
    AC_REQUIRE(wav_dev && lms_dev, "ac_logmel_fwd: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 1 + n_samples / fe->hop;
    if (gmax_dev) {
        ac::fill_kernel<<<1, 32, 0, st>>>(gmax_dev, -INFINITY, 1);
        AC_LAUNCHED("fill_kernel");
    }
    dim3 grid(ac::cdiv(T, ac::kFramesPerCta), batch);
    AC_TIMED("logmel", st);
    if (fe->n_fft == 512) {
        size_t sm = ac::logmel_smem_bytes<512>(fe->hop, fe->fbw_len);
        ac::logmel_kernel<512><<<grid, ac::kMelThreads, sm, st>>>(
            wav_dev, n_samples, T, fe->hop, fe->n_mels, fe->window_dev, fe->twiddle_dev, fe->fbw_dev,
            fe->fbw_len, fe->band_dev, lms_dev, gmax_dev);
    } else {
        size_t sm = ac::logmel_smem_bytes<1024>(fe->hop, fe->fbw_len);
        ac::logmel_kernel<1024><<<grid, ac::kMelThreads, sm, st>>>(
            wav_dev, n_samples, T, fe->hop, fe->n_mels, fe->window_dev, fe->twiddle_dev, fe->fbw_dev,
            fe->fbw_len, fe->band_dev, lms_dev, gmax_dev);
    }
    AC_LAUNCHED("logmel_kernel");
    return AC_OK;
}

int ac_db_clamp(float* x_dev, int64_t n, const float* gmax_dev, float top_db, void* stream) {
    AC_REQUIRE(x_dev && gmax_dev, "ac_db_clamp: null argument");
    if (n == 0) return AC_OK;
    int blocks = (int)std::min<int64_t>(ac::cdiv64(n, 256), ac::kNumSMs * 16);
    ac::db_clamp_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, n, gmax_dev, top_db);
    AC_LAUNCHED("db_clamp_kernel");
    return AC_OK;
}

}  // extern "C"
