// Fused STFT -> |X|^2 -> mel filterbank -> 10*log10 kernel (fp32).
//
// Replaces torchaudio MelSpectrogram + AmplitudeToDB as called by the reference at
// captioning/models/hf_wrapper.py:269-279,292-293 and captioning/models/cnn_encoder.py:338-350,418-419.
//
// One CTA = 32 consecutive frames of one clip.  The waveform segment the 32 frames cover is
// staged once in shared memory (reflect padding resolved while loading, coalesced reads); each
// warp then transforms 4 frames: window multiply, N/2-point complex FFT of the even/odd packed
// frame (radix-4 Stockham passes between two private shared-memory buffers), real-FFT unpacking to the power spectrum, banded
// mel accumulation (only the non-zero band of each filter, detected from the module's own
// `fb` buffer at create time), dB.  The 64x32 result tile is written back with frames
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
};

namespace ac {

constexpr int kFramesPerCta = 32;
constexpr int kMelThreads = 256;
constexpr int kMelWarps = kMelThreads / 32;
constexpr int kMaxMels = 64;

template <int NFFT>
__global__ void __launch_bounds__(kMelThreads)
logmel_kernel(const float* __restrict__ wav, int n_samples, int n_frames, int hop, int n_mels,
              const float* __restrict__ window, const float2* __restrict__ twiddle,
              const float* __restrict__ fbw, const int* __restrict__ band,
              float* __restrict__ out, float* __restrict__ gmax) {
    constexpr int M = NFFT / 2;           // complex FFT size
    constexpr int NF = NFFT / 2 + 1;      // one-sided bins
    constexpr int PSTRIDE = NF + 7;       // power buffer stride per warp

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg_len = (kFramesPerCta - 1) * hop + NFFT;
    float* s_seg = reinterpret_cast<float*>(smem_raw);                       // [seg_len] (+pad to 4)
    float2* s_tw = reinterpret_cast<float2*>(s_seg + ((seg_len + 3) & ~3));  // [M+1]
    float2* s_fft = s_tw + (M + 2);                                          // [warps][2][M]
    float* s_pow = reinterpret_cast<float*>(s_fft + 2 * kMelWarps * M);      // [warps][PSTRIDE]
    float* s_win = s_pow + kMelWarps * PSTRIDE;                              // [NFFT]
    float* s_tile = s_win + NFFT;                                            // [kMaxMels][33]
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
    __syncthreads();

    float2* za = s_fft + warp * 2 * M;       // ping-pong buffers of the Stockham FFT
    float2* zb = za + M;
    float* pw = s_pow + warp * PSTRIDE;
    float local_max = -INFINITY;

    for (int f = warp; f < kFramesPerCta; f += kMelWarps) {
        const int t = t0 + f;
        if (t >= n_frames) break;   // warp-uniform
        const float* x = s_seg + f * hop;
        // windowed frame packed as z[n] = x[2n] + i x[2n+1] (natural order)
        for (int n = lane; n < M; n += 32) {
            float2 v = *reinterpret_cast<const float2*>(x + 2 * n);
            float2 wv = *reinterpret_cast<const float2*>(s_win + 2 * n);
            za[n] = make_float2(v.x * wv.x, v.y * wv.y);
        }
        __syncwarp();
        // Stockham autosort FFT (no bit reversal; ping-pong between the warp's two buffers): radix-4 passes in
        // registers -- 4 passes for M = 256 instead of 8 radix-2 round trips through shared memory -- plus one
        // radix-2 pass when M is not a power of 4.  twid(m) = e^{-2 pi i m / NFFT}; the table covers m <= M.
        auto twid = [&](int m) {
            float2 t = s_tw[m > M ? m - M : m];
            return m > M ? make_float2(-t.x, -t.y) : t;
        };
        auto cmul = [](float2 p, float2 q) { return make_float2(p.x * q.x - p.y * q.y, p.x * q.y + p.y * q.x); };
        float2* in = za; float2* out = zb;
        int Ns = 1;
        for (; Ns * 4 <= M; Ns *= 4) {
            const int tstep = NFFT / (4 * Ns);
            for (int j = lane; j < M / 4; j += 32) {
                const int k = j & (Ns - 1);
                float2 v0 = in[j], v1 = in[j + M / 4], v2 = in[j + M / 2], v3 = in[j + 3 * M / 4];
                if (Ns > 1) {
                    v1 = cmul(v1, twid(k * tstep));
                    v2 = cmul(v2, twid(2 * k * tstep));
                    v3 = cmul(v3, twid(3 * k * tstep));
                }
                // forward DFT-4: multiplication by -i is (a, b) -> (b, -a)
                const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
                const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
                const int j0 = ((j - k) << 2) + k;
                out[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
                out[j0 + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);
                out[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
                out[j0 + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);
            }
            __syncwarp();
            float2* tmp = in; in = out; out = tmp;
        }
        if (Ns < M) {   // M = 2 * 4^p: final radix-2 pass
            const int tstep = NFFT / (2 * Ns);
            for (int j = lane; j < M / 2; j += 32) {
                const int k = j & (Ns - 1);
                const float2 v0 = in[j], v1 = cmul(in[j + M / 2], twid(k * tstep));
                const int j0 = ((j - k) << 1) + k;
                out[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
                out[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
            }
            __syncwarp();
            float2* tmp = in; in = out; out = tmp;
        }
        const float2* z = in;
        // unpack the real FFT and take |X|^2
        for (int k = lane; k <= M; k += 32) {
            float2 A = z[k & (M - 1)];
            float2 Bm = z[(M - k) & (M - 1)];
            float2 E = make_float2(0.5f * (A.x + Bm.x), 0.5f * (A.y - Bm.y));
            float2 O = make_float2(0.5f * (A.y + Bm.y), -0.5f * (A.x - Bm.x));
            float2 tw = s_tw[k];
            float xr = E.x + (O.x * tw.x - O.y * tw.y);
            float xi = E.y + (O.x * tw.y + O.y * tw.x);
            pw[k] = xr * xr + xi * xi;
        }
        __syncwarp();
        // banded mel filters + dB
        for (int m = lane; m < n_mels; m += 32) {
            int lo = band[3 * m], n = band[3 * m + 1], off = band[3 * m + 2];
            float acc = 0.0f;
            for (int j = 0; j < n; ++j) acc = fmaf(pw[lo + j], __ldg(fbw + off + j), acc);
            float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
            s_tile[m * 33 + f] = db;
            local_max = fmaxf(local_max, db);
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
static size_t logmel_smem_bytes(int hop) {
    constexpr int M = NFFT / 2;
    constexpr int NF = NFFT / 2 + 1;
    int seg_len = (kFramesPerCta - 1) * hop + NFFT;
    size_t fl = ((seg_len + 3) & ~3) + 2 * (M + 2) + 4 * kMelWarps * M + kMelWarps * (NF + 7) + NFFT +
                kMaxMels * 33;
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
    AC_CUDA(cudaFuncSetAttribute(ac::logmel_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ac::logmel_smem_bytes<512>(512)));
    AC_CUDA(cudaFuncSetAttribute(ac::logmel_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ac::logmel_smem_bytes<1024>(1024)));
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
        size_t sm = ac::logmel_smem_bytes<512>(fe->hop);
        ac::logmel_kernel<512><<<grid, ac::kMelThreads, sm, st>>>(
            wav_dev, n_samples, T, fe->hop, fe->n_mels, fe->window_dev, fe->twiddle_dev, fe->fbw_dev,
            fe->band_dev, lms_dev, gmax_dev);
    } else {
        size_t sm = ac::logmel_smem_bytes<1024>(fe->hop);
        ac::logmel_kernel<1024><<<grid, ac::kMelThreads, sm, st>>>(
            wav_dev, n_samples, T, fe->hop, fe->n_mels, fe->window_dev, fe->twiddle_dev, fe->fbw_dev,
            fe->band_dev, lms_dev, gmax_dev);
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
