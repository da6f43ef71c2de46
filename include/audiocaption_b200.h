/*
 * audiocaption_b200 -- C ABI of the B200 (sm_100a) hot path of wsntxxn/AudioCaption.
 *
 * The reference is pure Python: it has no FFI of its own.  The boundary this library sits
 * behind is the reference's nn.Module call convention; every entry point
 * below names the reference code it replaces (paths relative to the reference root).
 * Host-side mirrors of those modules (audiocaption_b200/captioning/...) bind these symbols
 * through ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; *_dev pointers are CUDA device pointers owned by the CALLER
 *     (PyTorch), *_host pointers are host memory; `stream` is a cudaStream_t passed as void*.
 *   - every call is asynchronous on `stream`; nothing synchronises unless stated.
 *   - the library owns only the opaque handles it creates (packed, BN-folded weight copies),
 *     freed with the matching *_destroy.
 *   - return 0 on success, negative on error; ac_last_error() gives the message
 *     (thread-local).
 *   - all arithmetic is fp32 ("dtype": "f32"); token ids / lengths are int64 as in the
 *     reference.
 */
#ifndef AUDIOCAPTION_B200_H
#define AUDIOCAPTION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AC_OK 0
#define AC_ERR_ARG (-1)
#define AC_ERR_CUDA (-2)
#define AC_ERR_WORKSPACE (-3)

typedef struct ac_frontend ac_frontend_t;
typedef struct ac_effb2 ac_effb2_t;
typedef struct ac_trm ac_trm_t;
typedef struct ac_cnn14 ac_cnn14_t;
typedef struct ac_bigru ac_bigru_t;
typedef struct ac_bah ac_bah_t;
typedef struct ac_sed ac_sed_t;
typedef struct ac_trm_train ac_trm_train_t;
typedef struct ac_bigru_train ac_bigru_train_t;

int ac_version(void);
const char* ac_last_error(void);
/* number of kernels launched by this library since process start (bench.py "gpu_launches") */
int64_t ac_launch_count(void);
/* Optional per-kernel CUDA-event timing for bench.py's roofline leg: when enabled every launch is
 * bracketed by an event pair on its stream; ac_timing_report synchronises the device and writes
 * "kernel_name launches total_ms" lines into buf, then clears the record. */
void ac_timing_enable(int on);
int ac_timing_report(char* buf, int buf_len);

/* ------------------------------------------------------------------ log-mel front-end
 * Replaces torchaudio MelSpectrogram + AmplitudeToDB as called at
 *   captioning/models/hf_wrapper.py:269-279,292-293   (EffB2: n_fft 512, hop 160, HTK)
 *   captioning/models/cnn_encoder.py:338-350,418-419  (Cnn14: n_fft 1024, hop 320, Slaney)
 * window_host[n_fft] and fb_host[n_freqs*n_mels] (row-major [n_freqs, n_mels]) are the
 * module's own state_dict buffers `spectrogram.window` and `mel_scale.fb`.
 * n_fft must be 512 or 1024 and win_length == n_fft; n_mels <= 64. */
int ac_frontend_create(const float* window_host, int n_fft, int hop, const float* fb_host,
                       int n_freqs, int n_mels, ac_frontend_t** out);
void ac_frontend_destroy(ac_frontend_t* fe);
int ac_frontend_num_frames(const ac_frontend_t* fe, int n_samples);
/* wav_dev [batch, n_samples] -> lms_dev [batch, n_mels, n_frames] = 10*log10(max(mel, 1e-10)).
 * gmax_dev (nullable, 1 float): receives the maximum over the whole batch (the batch-global
 * reference of AmplitudeToDB(top_db) for 3-D input); the clamp itself is applied by
 * ac_db_clamp or fused into ac_effb2_fwd's first layer. */
int ac_logmel_fwd(const ac_frontend_t* fe, const float* wav_dev, int batch, int n_samples,
                  float* lms_dev, float* gmax_dev, void* stream);
/* x = max(x, *gmax - top_db) in place (AmplitudeToDB(top_db=...)). */
int ac_db_clamp(float* x_dev, int64_t n, const float* gmax_dev, float top_db, void* stream);

/* ------------------------------------------------------------------ resampling (input pipeline)
 * Replaces torchaudio.functional.resample (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) as called at demo.py:36,
 * python_scripts/inference/inference.py:37 and captioning/datasets/caption_dataset.py:110-120.
 * orig / nw = orig_freq / new_freq divided by their gcd; coef_dev [nw][taps] (taps = 2 * width + orig) is the polyphase
 * kernel of torchaudio's `_get_sinc_resample_kernel`; wav_dev [batch, n_in] -> out_dev [batch, ac_resample_out_len]. */
int ac_resample_out_len(int n_in, int orig, int nw);
int ac_resample(const float* wav_dev, int batch, int n_in, const float* coef_dev, int orig, int nw, int taps, int width,
                float* out_dev, void* stream);

/* ------------------------------------------------------------------ pointwise GEMM (diagnostic)
 * The kernel behind every 1x1 convolution / Linear of the path (efficientnet_pytorch
 * `_expand_conv/_project_conv/_conv_head` + BN + swish, hf_wrapper.py:1045-1052 `attn_proj`):
 *   C[m,n] = act((sum_k A[m,k] * ascale[m / rows_per_group, k] * W[n,k]) * cscale[n] + cbias[n]) + R[m,n]
 * A [M,K], W [N,K], C/R [M,N] row-major fp32; ascale/cscale/cbias/R nullable; act 0 none, 1 swish, 2 relu.
 * path 0 = plain fp32 SIMT kernel, 1 = tcgen05 tensor-core kernel (TMA-fed, 3xTF32 split, fp32-level
 * accuracy; K % 8 == 0).  Exposed so that tests can check the kernel in isolation; synchronises. */
int ac_gemm(const float* A_dev, const float* W_dev, float* C_dev, int M, int N, int K,
            const float* ascale_dev, int rows_per_group, const float* cscale_dev, const float* cbias_dev,
            const float* R_dev, int act, int path, void* stream);

/* Diagnostic: pipeline trace of CTA 0 of the tensor-core GEMM (clock64 stamps, 8 event kinds x 256 events:
 * 0 TMA issue, 1 chunk landed, 2 transform done, 3 MMA start, 4 MMAs + commits issued, 5 accumulator full, 6 epilogue done,
 * 7 MMAs issued (before the commits)).
 * on != 0 enables tracing for subsequent launches; out_host (nullable, 2048 int64) receives and clears the trace. */
int ac_gemm_trace(int on, long long* out_host);

/* Diagnostic: one depthwise layer of an MBConv block on caller buffers (efficientnet_pytorch `_depthwise_conv`
 * with static-same padding (pad_lo before, pad_hi after, both axes) + `_bn1` folded to scale/bias + swish).
 * in [B,Hi,Wi,C] NHWC, w [k*k][C], out [B,Ho,Wo,C]; partial [B][ac_dwconv_partial_rows(...)][C] receives per-tile
 * channel sums of `out` (the SE squeeze numerator).  k in {3,5}, s in {1,2}, C % 4 == 0. */
int ac_dwconv(const float* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev, float* out_dev,
              float* partial_dev, int B, int Hi, int Wi, int C, int k, int s, int pad_lo, int pad_hi, void* stream);
int ac_dwconv_partial_rows(int Ho, int Wo, int C, int k, int s);

/* Diagnostic: one 3x3 / stride 1 / pad 1 convolution + per-channel scale and bias + activation, the body of
 * captioning/models/cnn_encoder.py:32-75 `ConvBlock` (conv -> BatchNorm(eval) -> ReLU), as an implicit GEMM on the
 * tcgen05 pipeline (4-D TMA boxes shifted per tap; 3xTF32 split).  in [B,H,W,Cin] NHWC, w [Cout,Cin,3,3] (PyTorch
 * layout), scale/bias [Cout] nullable, out [B,H,W,Cout] NHWC; Cin % 32 == 0, Cout % 32 == 0, W <= 128; act 0 | 2.
 * Synchronises. */
int ac_conv3x3(const float* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev, float* out_dev,
               int B, int H, int W, int Cin, int Cout, int act, void* stream);

/* ------------------------------------------------------------------ EfficientNet-B2 encoder
 * Replaces hf_wrapper.py:218-241 `_EffiNet.forward` (efficientnet_pytorch 0.7.1
 * `extract_features` + mean over frequency), eval mode.
 * tensors_dev: the `backbone.eff_net` state_dict tensors in state_dict order WITHOUT the
 * `num_batches_tracked` entries; numels[i] is checked against the built-in B2 plan. */
int ac_effb2_create(const float* const* tensors_dev, const int64_t* numels, int n_tensors,
                    void* stream, ac_effb2_t** out);
void ac_effb2_destroy(ac_effb2_t* net);
int ac_effb2_num_tensors(void);
int ac_effb2_out_frames(int n_frames);
int ac_effb2_out_dim(void);
size_t ac_effb2_workspace_bytes(int batch, int n_mels, int n_frames);
/* lms_dev [batch, n_mels, n_frames] (dB, NOT yet clamped when gmax_dev != NULL: the
 * top_db clamp max(x, *gmax - top_db) is applied while the stem convolution loads its
 * input) -> attn_emb_dev [batch, out_frames, 1408]. */
int ac_effb2_fwd(const ac_effb2_t* net, const float* lms_dev, const float* gmax_dev, float top_db,
                 int batch, int n_mels, int n_frames, float* attn_emb_dev,
                 void* workspace_dev, size_t workspace_bytes, void* stream);
/* Introspection of the built-in block plan (tests compare it with the oracle's):
 * fills out[9] = {cin, cout, expand, k, stride, pad_lo, pad_hi, n_squeeze, has_skip}. */
int ac_effb2_block_info(int block, int* out9);

/* ------------------------------------------------------------------ Cnn14 (PANNs) encoder
 * Replaces captioning/models/cnn_encoder.py:414-464 `Cnn14Encoder.forward` after the log-mel front-end (HF copy
 * hf_wrapper.py:1259-1304), eval mode: bn0 over mel -> 6 x `ConvBlock` (cnn_encoder.py:32-75) with 2x2 average
 * pooling (block 6: none) -> mean over mel -> attn_emb; fc_emb = relu(fc1(max_with_lens + mean_with_lens))
 * (captioning/utils/model_util.py:41-84).
 * tensors_dev (ac_cnn14_num_tensors() = 66 entries, the module's state_dict order without the
 * `melspec_extractor.*` buffers and the `num_batches_tracked` entries): bn0.{weight,bias,running_mean,running_var};
 * for block 1..6: conv1.weight [Cout,Cin,3,3], conv2.weight, bn1.{weight,bias,running_mean,running_var}, bn2.{...};
 * fc1.weight [2048,2048], fc1.bias. */
int ac_cnn14_num_tensors(void);
int ac_cnn14_out_dim(void);
int ac_cnn14_out_frames(int n_frames);
size_t ac_cnn14_workspace_bytes(int batch, int n_mels, int n_frames);
int ac_cnn14_create(const float* const* tensors_dev, const int64_t* numels, int n_tensors, void* stream,
                    ac_cnn14_t** out);
void ac_cnn14_destroy(ac_cnn14_t* net);
/* lms_dev [batch, 64, n_frames] (dB, the reference's AmplitudeToDB output), lens_dev [batch] int64 (valid output
 * frames per clip, `feat_length`) -> attn_emb_dev [batch, out_frames, 2048], fc_emb_dev [batch, 2048]. */
int ac_cnn14_fwd(const ac_cnn14_t* net, const float* lms_dev, int batch, int n_mels, int n_frames,
                 const int64_t* lens_dev, float* attn_emb_dev, float* fc_emb_dev,
                 void* workspace_dev, size_t workspace_bytes, void* stream);

/* Precision of the 3x3 convolutions: tf32_passes = 3 (default) issues every product as three TF32 MMAs (fp32-level accuracy:
 * the mode of every fp32 parity claim), 1 uses plain TF32 operands (10-bit mantissa, fp32 accumulation) -- the tensor-core
 * mode for the configurations BASELINE.json states in bf16 (training, temporal captioner). */
int ac_cnn14_set_precision(ac_cnn14_t* net, int tf32_passes);
/* Caps the persistent grid of the bf16-mode convolutions at n CTAs (0 = one per SM), so that a frozen encoder running on a
 * second stream (the training step's look-ahead, audiocaption_b200/train_step.py) leaves SMs to the trainable chain. */
int ac_cnn14_set_sm_limit(ac_cnn14_t* net, int n);
int ac_sed_set_precision(ac_sed_t* net, int tf32_passes);
/* ac_conv3x3 with the precision switch (diagnostic). */
int ac_conv3x3_p(const float* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev, float* out_dev,
                 int B, int H, int W, int Cin, int Cout, int act, int tf32_passes, void* stream);
/* The bf16 form of the same convolution (diagnostic; the "bf16" precision mode of ac_cnn14_set_precision, value 16, runs it
 * on every layer): in_dev / out_dev are NHWC bf16 [B,H,W,Cin] / [B,H,W,Cout], w_dev the fp32 Conv2d weight, bias fp32,
 * Cin % 64 == 0.  Replaces the same ConvBlock arithmetic (captioning/models/cnn_encoder.py:32-75) at bf16 operand precision
 * with fp32 accumulation. */
int ac_conv3x3_bf16(const void* in_dev, const float* w_dev, const float* scale_dev, const float* bias_dev, void* out_dev,
                    int B, int H, int W, int Cin, int Cout, int act, void* stream);

/* Train-mode forward of the frozen encoder (BatchNorm folded = eval, `freeze_cnn_bn`), with the functional dropouts of
 * captioning/models/cnn_encoder.py:432-456 active: p_conv after each ConvBlock (reference 0.2), p_fc around fc1 (0.5).
 * Masks are a function of (seed, site, element index).  p_conv = p_fc = 0 is ac_cnn14_fwd. */
int ac_cnn14_fwd_train(const ac_cnn14_t* net, const float* lms_dev, int batch, int n_mels, int n_frames,
                       const int64_t* lens_dev, float p_conv, float p_fc, uint64_t seed, float* attn_emb_dev,
                       float* fc_emb_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ sound-event tagger of the temporal captioner
 * Replaces captioning/models/hf_wrapper.py:1791-1859 `Cnn8rnnSedModel.forward_prob` (bn0, 4 ConvBlocks with 'avg+max'
 * pooling (2,2)(2,2)(1,2)(1,2), mean over mel, fc1 + ReLU, bidirectional GRU(512 -> 256), fc_audioset, sigmoid,
 * clamp(1e-7, 1)) and the double threshold of :123-162 (`double_threshold(x, high, low)`), eval mode.
 * tensors_dev (ac_sed_num_tensors() = 56, state_dict order without `num_batches_tracked`): bn0 x4; blocks 1..4:
 * conv1.weight, conv2.weight, bn1 x4, bn2 x4; fc1.weight [512,512], fc1.bias; rnn.{weight_ih,weight_hh,bias_ih,bias_hh}_l0
 * and the `_reverse` four; fc_audioset.weight [classes,512], fc_audioset.bias. */
int ac_sed_num_tensors(void);
int ac_sed_segments(int n_frames);   /* n_frames / 4 */
int ac_sed_create(const float* const* tensors_dev, const int64_t* numels, int n_tensors, int classes, void* stream,
                  ac_sed_t** out);
void ac_sed_destroy(ac_sed_t* net);
size_t ac_sed_workspace_bytes(const ac_sed_t* net, int batch, int n_mels, int n_frames);
/* lms_dev [batch, 64, n_frames] -> prob_dev (nullable) [batch, segments, classes] = `segmentwise_output`;
 * labels_dev [batch, segments, classes] uint8 = the double-thresholded decisions at segment resolution (the
 * reference's frame-level matrix is this one repeated 4x along time and padded with its last row).
 * runs_dev (nullable, 16-byte aligned) [max_runs][4] int32 + n_runs_dev (1 int32): every kept run as (clip, class, first
 * segment, one past the last segment), unordered; *n_runs_dev counts all runs even beyond max_runs (overflow check). */
int ac_sed_fwd(const ac_sed_t* net, const float* lms_dev, int batch, int n_mels, int n_frames, float high, float low,
               float* prob_dev, unsigned char* labels_dev, int* runs_dev, int max_runs, int* n_runs_dev,
               void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ bidirectional GRU encoder
 * Replaces captioning/models/rnn_encoder.py:34-49 `RnnEncoder.forward` = pack_wrapper(nn.GRU(batch_first=True,
 * bidirectional=True, num_layers=L), x, lens) (captioning/utils/model_util.py:10-27; HF copy hf_wrapper.py:1307-1347),
 * eval mode: every clip runs over its own t < lens[b] (the reverse direction starts at lens[b]-1), outputs at
 * t >= lens[b] are zero.  hidden must be 256.
 * tensors_dev: nn.GRU state_dict order, 8 per layer: weight_ih_l{k} [768, D], weight_hh_l{k} [768, 256],
 * bias_ih_l{k}, bias_hh_l{k}, then the same four with the `_reverse` suffix. */
int ac_bigru_create(const float* const* tensors_dev, const int64_t* numels, int n_tensors, int input_dim, int hidden,
                    int num_layers, void* stream, ac_bigru_t** out);
void ac_bigru_destroy(ac_bigru_t* net);
int ac_bigru_out_dim(const ac_bigru_t* net);
size_t ac_bigru_workspace_bytes(const ac_bigru_t* net, int batch, int T);
/* x_dev [batch, T_in, input_dim], lens_dev [batch] int64 -> out_dev [batch, T_out, 512] with T_out <= T_in
 * (pad_packed_sequence returns max(lens) frames: pass T_out = max(lens)). */
int ac_bigru_fwd(const ac_bigru_t* net, const float* x_dev, const int64_t* lens_dev, int batch, int T_in, int T_out,
                 float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* masked mean over time, hf_wrapper.py:330-352 `mean_with_lens`:
 * x_dev [batch, T, D], lens_dev [batch] int64 -> out_dev [batch, D] */
int ac_masked_mean(const float* x_dev, const int64_t* lens_dev, int batch, int T, int D,
                   float* out_dev, void* stream);

/* ------------------------------------------------------------------ Transformer caption decoder
 * Replaces hf_wrapper.py:976-1068 / captioning/models/transformer_decoder.py:11-103
 * (TransformerDecoder.forward) and the decode loops of captioning/models/base.py:152-218
 * (stepwise greedy) and :254-361 (beam search), with a KV cache (eval mode, no dropout).
 * tensors_dev order: word_embedding.weight, pos_encoder.pe, then per layer
 *   self_attn.in_proj_weight, self_attn.in_proj_bias, self_attn.out_proj.weight, .bias,
 *   multihead_attn.in_proj_weight, .in_proj_bias, .out_proj.weight, .bias,
 *   linear1.weight, .bias, linear2.weight, .bias, norm1.weight, .bias, norm2.*, norm3.*,
 * then classifier.weight, attn_proj.0.weight, attn_proj.0.bias, attn_proj.3.weight, .bias. */
int ac_trm_create(const float* const* tensors_dev, const int64_t* numels, int n_tensors,
                  int d_model, int nhead, int nlayers, int dim_ff, int vocab, int attn_emb_dim,
                  int pe_len, void* stream, ac_trm_t** out);
void ac_trm_destroy(ac_trm_t* dec);
int ac_trm_num_tensors(int nlayers);
size_t ac_trm_workspace_bytes(const ac_trm_t* dec, int rows, int t_mem, int max_len);
/* Greedy decode of `batch` clips, all max_len steps on the device (rows that emitted <end>
 * keep emitting <end>, which equals the reference's early-stopped output).
 * attn_emb_dev [batch, t_mem, attn_emb_dim]; attn_emb_len_dev [batch] int64.
 * seq_dev [batch, max_len] int64; logprob_dev [batch, max_len] (nullable);
 * logit_dev [batch, max_len, vocab] (nullable); embed_dev [batch, max_len, d_model] (nullable). */
int ac_trm_greedy(const ac_trm_t* dec, const float* attn_emb_dev, const int64_t* attn_emb_len_dev,
                  int batch, int t_mem, int max_len, int start_idx, int end_idx, int pad_idx,
                  int64_t* seq_dev, float* logprob_dev, float* logit_dev, float* embed_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);
/* Re-read the source tensors (same order / sizes as at creation) into an existing handle: asynchronous, no allocation.
 * The training loop calls it once per optimizer step. */
int ac_trm_update(ac_trm_t* dec, const float* const* tensors_dev, int n_tensors, void* stream);
/* The sampling half of scheduled-sampling training (captioning/models/transformer_model.py:34-57 under
 * captioning/models/base.py:152-170 with mode == "train"): the model's own token row, step by step with the KV cache.
 * forced_dev [batch, max_len] int64: >= 0 = emit (and feed back) that token at the step -- the step's coin chose the
 * ground-truth prefix, whose arg-max the dense pass already produced; < 0 = take this decode's arg-max.  No early stop,
 * no <end> forcing (train mode).  seq_dev [batch, max_len] int64, logprob_dev nullable. */
int ac_trm_sample_forced(const ac_trm_t* dec, const float* attn_emb_dev, const int64_t* attn_emb_len_dev, int batch, int t_mem,
                         int max_len, int start_idx, int end_idx, int pad_idx, const int64_t* forced_dev, int64_t* seq_dev,
                         float* logprob_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* Beam search, one independent search per clip with the reference's bookkeeping
 * (double log-softmax with temperature, -1000 penalty, `== beam_size` stop rule,
 * score/(t+1) ranking).  seq_dev [batch, max_len] int64 = best beam per clip. */
int ac_trm_beam(const ac_trm_t* dec, const float* attn_emb_dev, const int64_t* attn_emb_len_dev,
                int batch, int t_mem, int max_len, int beam_size, float temp,
                int start_idx, int end_idx, int pad_idx, int64_t* seq_dev,
                void* workspace_dev, size_t workspace_bytes, void* stream);

/* Diagnostic: clock64 stamps of the phases of decode step 5 (CTA 0) of subsequent ac_trm_greedy launches
 * (0 step start; per layer l: 1+12l .. 12+12l around the six GEMVs and their cluster barriers; 40 classifier start,
 * 41 classifier done, 42 arg-max done).  out_host (nullable): 64 int64. */
int ac_trm_trace(int on, long long* out_host);

/* ------------------------------------------------------------------ temporal Bahdanau-attention GRU decoder
 * Replaces captioning/models/hf_wrapper.py:1377-1414 (`Seq2SeqAttention`), :1444-1554 (`BahAttnCatFcDecoder`,
 * `TemporalBahAttnDecoder.forward`) and the decode glue :1557-1788 (`Seq2SeqAttnModel`, `TemporalSeq2SeqAttnModel`)
 * under `CaptionModel.stepwise_forward` / `beam_search` (captioning/models/base.py:152-218, 254-361), eval mode.
 * All widths (embedding, GRU hidden, attention, memory, fc) are 512 as in `Cnn14RnnTempAttnGruConfig`.
 * tensors_dev (15, the decoder's state_dict order): word_embedding.weight [V,512], classifier.weight [V,512],
 * classifier.bias [V], model.weight_ih_l0 [1536,1536], model.weight_hh_l0 [1536,512], model.bias_ih_l0,
 * model.bias_hh_l0, attn.v [512], attn.h2attn.weight [512,1024], attn.h2attn.bias, fc_proj.weight [512,512],
 * fc_proj.bias, ctx_proj.weight [512,512], ctx_proj.bias, temporal_embedding.weight [4,512]. */
int ac_bah_num_tensors(void);
int ac_bah_create(const float* const* tensors_dev, const int64_t* numels, int n_tensors, int vocab, void* stream,
                  ac_bah_t** out);
void ac_bah_destroy(ac_bah_t* dec);
/* rows = batch for greedy, batch * beam for beam search */
size_t ac_bah_workspace_bytes(const ac_bah_t* dec, int rows, int T);
/* fc_emb_dev [batch,512], attn_emb_dev [batch,T,512], lens_dev / tags_dev [batch] int64 (tags in 0..3) ->
 * seq_dev [batch,max_len] int64 (rows that emitted <end> stay <end>), logprob_dev [batch,max_len] (nullable),
 * logit_dev [batch,max_len,V] (nullable; when given, finished rows keep being evaluated as the reference does). */
int ac_bah_greedy(const ac_bah_t* dec, const float* fc_emb_dev, const float* attn_emb_dev, const int64_t* lens_dev,
                  const int64_t* tags_dev, int batch, int T, int max_len, int start_idx, int end_idx,
                  int64_t* seq_dev, float* logprob_dev, float* logit_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);
/* The same decode with the decoder's step-level inputs / outputs exposed (`TemporalBahAttnDecoder.forward`,
 * hf_wrapper.py:1513-1554, is ONE step: call with max_len = 1): state_in_dev (nullable) [batch,512] = the GRU state to
 * start from (`state`; default zeros), first_word_dev (nullable) [batch] int64 = the word fed at step 0 (when NULL step 0
 * feeds temporal_embedding(tag), the reference's `t == 0` branch); state_out_dev (nullable) [batch,512] = state after the
 * last executed step (= `embed`: the 1-layer GRU's output), attn_w_out_dev (nullable) [batch,max_len,T] = `attn_weight`
 * of every step (hf_wrapper.py:1603-1607 stores them as [batch,T,max_len]). */
int ac_bah_greedy_ex(const ac_bah_t* dec, const float* fc_emb_dev, const float* attn_emb_dev, const int64_t* lens_dev,
                     const int64_t* tags_dev, int batch, int T, int max_len, int start_idx, int end_idx,
                     const float* state_in_dev, const int64_t* first_word_dev, int64_t* seq_dev, float* logprob_dev,
                     float* logit_dev, float* state_out_dev, float* attn_w_out_dev, void* workspace_dev,
                     size_t workspace_bytes, void* stream);
/* per-clip beam search with the reference's bookkeeping (beam 1..5); the GRU state follows `prev_words_beam`. */
int ac_bah_beam(const ac_bah_t* dec, const float* fc_emb_dev, const float* attn_emb_dev, const int64_t* lens_dev,
                const int64_t* tags_dev, int batch, int T, int max_len, int beam, float temp, int start_idx,
                int end_idx, int64_t* seq_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ================================================================== training step
 * The training entry points read LIVE parameter storage (the module's own tensors) and write gradients into caller-owned
 * buffers; nothing is copied at create time.  Call *_refresh once after every optimizer step (it re-packs the weights for
 * the tensor-core GEMMs), then forward, then backward.  Activations needed by the backward pass live in the caller's
 * workspace between the two calls.  Dropout masks are functions of (seed, site, element index): pass the same p_drop /
 * seed to the forward and backward calls of one step; p_drop = 0 gives the deterministic (eval-mode) arithmetic. */

/* ------------------------------------------------------------------ bi-GRU encoder, training
 * Replaces captioning/models/rnn_encoder.py:34-49 `RnnEncoder.forward` in train mode (nn.GRU inter-layer dropout) and its
 * autograd backward.  params_dev / grads_dev: nn.GRU order as for ac_bigru_create (a NULL grads entry = frozen).
 * need_input_grad != 0 also prepares the gradient w.r.t. x (not needed when the CNN below is frozen). */
int ac_bigru_train_create(const float* const* params_dev, float* const* grads_dev, const int64_t* numels, int n_tensors,
                          int input_dim, int hidden, int num_layers, int need_input_grad, void* stream,
                          ac_bigru_train_t** out);
void ac_bigru_train_destroy(ac_bigru_train_t* net);
size_t ac_bigru_train_workspace_bytes(const ac_bigru_train_t* net, int batch, int T);
int ac_bigru_train_refresh(ac_bigru_train_t* net, void* stream);
/* x_dev [batch, T, input_dim], lens_dev [batch] int64 (1..T) -> out_dev [batch, T, 512] (zero past each length). */
int ac_bigru_train_fwd(ac_bigru_train_t* net, const float* x_dev, const int64_t* lens_dev, int batch, int T, float p_drop,
                       uint64_t seed, float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* dout_dev [batch, T, 512] -> every non-NULL grads_dev tensor is overwritten; dx_dev (nullable) [batch, T, input_dim]. */
int ac_bigru_train_bwd(ac_bigru_train_t* net, const float* x_dev, const int64_t* lens_dev, const float* dout_dev, int batch,
                       int T, float p_drop, uint64_t seed, float* dx_dev, void* workspace_dev, size_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------------------------ Transformer decoder, dense full-prefix forward + backward
 * Replaces captioning/models/transformer_decoder.py:80-103 `TransformerDecoder.forward` as the training loop calls it
 * (captioning/models/transformer_model.py:20-57 under captioning/models/base.py:131-170) and its autograd backward.
 * With p_drop = 0 it is also the eval-mode `forward(input_dict)` on a full prefix.
 * params_dev / grads_dev: the tensor order of ac_trm_create.  Token rows: word_dev [n_seq_max, L] int64 and key_pad_dev
 * [n_seq_max, L] uint8 (`cap_padding_mask`); row s attends to the memory of clip s % B (scheduled sampling runs the
 * ground-truth captions as rows [0, B) and the model's own samples as rows [B, 2B)).  n_seq_max, L, B, T fix the workspace
 * layout and must be the same in every call of one step. */
int ac_trm_train_num_tensors(int nlayers);
int ac_trm_train_create(const float* const* params_dev, float* const* grads_dev, const int64_t* numels, int n_tensors,
                        int d_model, int nhead, int nlayers, int dim_ff, int vocab, int attn_emb_dim, int pe_len, void* stream,
                        ac_trm_train_t** out);
void ac_trm_train_destroy(ac_trm_train_t* dec);
size_t ac_trm_train_workspace_bytes(const ac_trm_train_t* dec, int n_seq_max, int L, int B, int T);
int ac_trm_train_vocab_padded(const ac_trm_train_t* dec);      /* row stride of the logits: vocab rounded up to 8 */
int ac_trm_train_refresh(ac_trm_train_t* dec, void* stream);
/* attn_emb_dev [B, T, attn_emb_dim] -> projected memory + per-layer cross-attention keys / values (in the workspace) */
int ac_trm_train_memory_fwd(ac_trm_train_t* dec, const float* attn_emb_dev, int B, int T, int n_seq_max, int L, float p_drop,
                            uint64_t seed, void* workspace_dev, size_t workspace_bytes, void* stream);
/* token rows [seq0, seq0 + n_seq) through the decoder layers; attn_len_dev [B] int64 = valid memory frames */
int ac_trm_train_seq_fwd(ac_trm_train_t* dec, const int64_t* word_dev, const unsigned char* key_pad_dev, int seq0, int n_seq,
                         int n_seq_max, int L, const int64_t* attn_len_dev, int B, int T, float p_drop, uint64_t seed,
                         void* workspace_dev, size_t workspace_bytes, void* stream);
/* logits_dev [n_rows, vocab_padded] = classifier(hidden[rows_dev[i]]) (rows_dev int32, NULL = identity); embed_dev (nullable)
 * [n_rows, d_model]; keep != 0 records the selection for ac_trm_train_bwd. */
int ac_trm_train_logits(ac_trm_train_t* dec, const int* rows_dev, int n_rows, int n_seq_max, int L, int B, int T,
                        float* logits_dev, float* embed_dev, int keep, void* workspace_dev, size_t workspace_bytes, void* stream);
/* dlogits_dev [n_rows, vocab_padded] -> all parameter gradients (overwritten) and dattn_emb_dev (nullable) [B, T, attn_emb_dim] */
int ac_trm_train_bwd(ac_trm_train_t* dec, const float* dlogits_dev, const int* rows_dev, int n_rows, const int64_t* word_dev,
                     const unsigned char* key_pad_dev, int n_seq, int n_seq_max, int L, const float* attn_emb_dev,
                     const int64_t* attn_len_dev, int B, int T, float p_drop, uint64_t seed, float* dattn_emb_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* SpecAugment on the dB log-mel (cnn_encoder.py:352-353,424-425: torchlibrosa `SpecAugmentation(time_drop_width=64,
 * time_stripes_num=2, freq_drop_width=8, freq_stripes_num=2)`, training only): zeroes, per clip, n_stripes frame ranges and
 * n_stripes mel ranges.  stripes_dev [batch][2 * n_stripes][2] int32 (begin, width): frame ranges first, then mel ranges
 * (drawn by the caller).  lms_dev [batch, n_mels, n_frames] is modified in place. */
int ac_specaug_apply(float* lms_dev, int batch, int n_mels, int n_frames, const int* stripes_dev, int n_stripes, void* stream);

/* ------------------------------------------------------------------ loss and optimizer
 * captioning/losses/loss.py:51-74 `LabelSmoothingLoss.forward` (reduction "mean") fused with its gradient:
 * logit_dev [B, L, ld_logit >= V], tgt_dev [B, *] int64 with row stride ld_tgt (tgt = cap[:, 1:] is a strided view),
 * tgt_len_dev [B] int64 -> loss_dev[0]; dlogit_dev (nullable, same layout as logit) = grad_scale * dloss/dlogit
 * (columns V..ld_logit-1 are zeroed).  workspace: B * L floats. */
int ac_ls_ce_fwd_bwd(const float* logit_dev, int ld_logit, const int64_t* tgt_dev, int ld_tgt, const int64_t* tgt_len_dev, int B,
                     int L, int V, float smoothing, float grad_scale, float* loss_dev, float* dlogit_dev, void* workspace_dev,
                     size_t workspace_bytes, void* stream);
/* `sample_next_word(method="greedy")` (captioning/models/base.py:214-218) over M logit rows of stride ld_logit:
 * idx_dev[m] = first arg-max over the V valid columns, logprob_dev[m] (nullable) = its log-softmax value. */
int ac_argmax_rows(const float* logit_dev, int ld_logit, int M, int V, int64_t* idx_dev, float* logprob_dev, void* stream);
/* python_scripts/train_eval/run.py:123-127: skip when the loss is NaN, else clip_grad_norm_(max_norm) over the flat
 * gradient, then torch.optim.Adam.step (L2 weight decay, bias correction from the device-side step counter).
 * grad_scale multiplies the gradient first (1 / world_size after a sum all-reduce).  norm_out_dev (nullable) receives the
 * total gradient norm before clipping. */
size_t ac_clip_adam_workspace_bytes(void);
int ac_clip_adam(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, float lr,
                 float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                 const float* loss_dev, int* step_dev, float* norm_out_dev, void* workspace_dev, size_t workspace_bytes,
                 void* stream);

/* ---- SM partition for the training step's look-ahead (csrc/sm_partition.cu) ------------------------------------------
 * The reference's loop (python_scripts/train_eval/run.py:77-148) runs the frozen CNN and the trainable part of a step back
 * to back; here the frozen encoder of batch i+1 runs beside the trainable part of batch i on a DISJOINT set of SMs (CUDA
 * green contexts).  Set 0 gets about `sms_first` SMs (8-SM granularity), set 1 the rest; one stream per set. */
typedef struct ac_sm_partition ac_sm_partition_t;
int ac_sm_partition_create(int sms_first, ac_sm_partition_t** out);
void* ac_sm_partition_stream(const ac_sm_partition_t* p, int which);   /* cudaStream_t of set 0 / 1 */
int ac_sm_partition_sms(const ac_sm_partition_t* p, int which);        /* SMs in set 0 / 1 */
void ac_sm_partition_destroy(ac_sm_partition_t* p);

#ifdef __cplusplus
}
#endif
#endif
