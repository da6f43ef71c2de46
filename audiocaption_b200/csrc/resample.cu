// Band-limited sinc resampling of a batch of waveforms by a rational factor (sm_100a).
//
// Replaces torchaudio.functional.resample(waveform, orig_freq, new_freq) (resampling_method "sinc_interp_hann",
// lowpass_filter_width 6, rolloff 0.99) as called by the reference at demo.py:36, python_scripts/inference/inference.py:37
// and captioning/datasets/caption_dataset.py:110-120 -- the 32 kHz -> 16 kHz step in front of the EfficientNet-B2
// captioner and the 44.1 kHz -> 32 kHz step of the Clotho pipeline.
//
// With orig = orig_freq / gcd and new = new_freq / gcd, output sample n = k * new + j is the dot product of polyphase
// filter j (taps = 2 * width + orig coefficients) with the zero-padded input starting at k * orig - width.  The
// coefficient table [new][taps] is built on the host by the module that owns it (the same closed form torchaudio uses)
// and passed in; the kernel is a plain FIR: HBM-bound (reads each input sample once through L1/L2, writes each output
// once), one thread per output sample, coefficients through the read-only cache (32 kHz -> 16 kHz: 1 phase x 28 taps).
#include "common.cuh"

namespace ac {

__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, const float* __restrict__ coef, float* __restrict__ y, int n_in, int n_out,
                int orig, int nw, int taps, int width) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_out) return;
    const int k = n / nw, j = n - k * nw;
    const float* xb = x + (size_t)b * n_in;
    const float* c = coef + (size_t)j * taps;
    const int i0 = k * orig - width;
    float acc = 0.0f;
    for (int i = 0; i < taps; ++i) {
        const int p = i0 + i;
        if (p >= 0 && p < n_in) acc = fmaf(__ldg(c + i), __ldg(xb + p), acc);
    }
    y[(size_t)b * n_out + n] = acc;
}

}  // namespace ac

extern "C" {

int ac_resample_out_len(int n_in, int orig, int nw) {
    return (int)(((int64_t)nw * n_in + orig - 1) / orig);       // ceil(new * length / orig)
}

int ac_resample(const float* wav_dev, int batch, int n_in, const float* coef_dev, int orig, int nw, int taps, int width,
                float* out_dev, void* stream) {
    using namespace ac;
    AC_REQUIRE(batch >= 0 && n_in >= 1 && orig >= 1 && nw >= 1 && taps == 2 * width + orig && width >= 1,
               "ac_resample: bad arguments (orig %d new %d taps %d width %d)", orig, nw, taps, width);
    if (batch == 0) return AC_OK;
    AC_REQUIRE(wav_dev && coef_dev && out_dev && batch <= 65535, "ac_resample: null argument or batch > 65535");
    const int n_out = ac_resample_out_len(n_in, orig, nw);
    cudaStream_t st = (cudaStream_t)stream;
    AC_TIMED("resample", st);
    resample_kernel<<<dim3(cdiv(n_out, 256), batch), 256, 0, st>>>(wav_dev, coef_dev, out_dev, n_in, n_out, orig, nw, taps, width);
    AC_LAUNCHED("resample_kernel");
    return AC_OK;
}

}  // extern "C"
