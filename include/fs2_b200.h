/*
 * fs2_b200.h -- C ABI of the B200-native FastSpeech2-align inference forward.
 *
 * This is the drop-in boundary for ONE hot path of SMART-TTS/SMART-NAR_Fast_TTS:
 * `FastSpeech2Align.forward` with `mel_lens=None`
 * (reference: model/fastspeech2_align.py:30-100).  The reference has no FFI of its
 * own (it is pure PyTorch), so the entry points below are what a ctypes binding in
 * the reference's `model/fastspeech2_align.py` would call.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  All tensor pointers are DEVICE
 *    pointers on the handle's device unless a parameter says "host".
 *  - tensors are dense row-major fp32 unless stated; ids / lengths are int64 like the
 *    reference's `texts` / `src_lens` (utils/tools.py:56-63); masks are uint8 (1 = padded),
 *    the layout of a torch.bool tensor.
 *  - every function returns FS2_OK (0) or a negative error code, never throws, never
 *    exits.  `fs2_last_error` gives the message for the last failure on that handle.
 *  - all work is enqueued on the `stream` argument (a cudaStream_t / CUstream passed as
 *    void*; NULL = legacy default stream).  The only blocking call is the 8-byte D2H of
 *    {T_max, frame count} at the end of `fs2_forward_stage1` (`fs2_forward_stage1_async` leaves even
 *    that to the caller).
 *  - a handle is bound to one device and is not thread-safe; separate handles are
 *    independent.  The library owns packed weights and workspace; the caller owns every
 *    tensor it passes in or receives results in.
 */
#ifndef FS2_B200_H
#define FS2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the only exported symbols of libfs2_b200.so */
#endif

#define FS2_OK 0
#define FS2_ERR_INVALID (-1)      /* bad argument */
#define FS2_ERR_CUDA (-2)         /* CUDA runtime / driver error */
#define FS2_ERR_STATE (-3)        /* call order violated (weights not loaded, stage2 before stage1) */
#define FS2_ERR_UNSUPPORTED (-4)  /* dims outside what the kernels are built for */
#define FS2_ERR_MISSING_WEIGHT (-5)
#define FS2_ERR_NO_DEVICE (-6)    /* no CUDA device: there is NO CPU fallback */

/* arithmetic used by a segment of the path */
#define FS2_PREC_FP32 0 /* fp32 FFMA kernels (fp32-faithful; decides durations and pitch/energy buckets) */
#define FS2_PREC_BF16 1 /* tcgen05 tensor-core kernels: bf16 operands, fp32 accumulate in TMEM */
#define FS2_PREC_BF16X3 2 /* tcgen05, fp32-faithful: operands split into 3 bf16 terms, 6 cross products per MAC block
                             accumulated in fp32 (error ~2^-23 per product); attention stays fp32 FFMA */
#define FS2_PREC_F16X2 3  /* tcgen05, fp32-faithful: operands carried as 2 power-of-two-scaled fp16 terms (22 significant
                             bits), 3 cross products per MAC block -- half the tensor work of BF16X3 at the same error
                             level as an fp32 FFMA GEMM; activations saturate at |x| = 4094; attention runs on the tensor
                             cores too (Q, K, V, P as scaled fp16 hi / lo planes, fp32 softmax) */

typedef struct fs2_handle fs2_handle;

/* Model hyper-parameters: config/LJSpeech/model.yaml:1-25, preprocess.yaml:24-32,
 * PostNet() defaults transformer/Layers.py:112-118, vocab = len(symbols)+1 (Models.py:40). */
typedef struct fs2_dims {
  int32_t vocab;         /* 361 */
  int32_t d_model;       /* 256 (encoder_hidden == decoder_hidden) */
  int32_t n_enc_layers;  /* 4 */
  int32_t n_dec_layers;  /* 4 */
  int32_t n_heads;       /* 2 */
  int32_t d_ffn;         /* 1024 */
  int32_t ffn_k1;        /* 9 */
  int32_t ffn_k2;        /* 1 */
  int32_t vp_filter;     /* 256 */
  int32_t vp_kernel;     /* 3 */
  int32_t n_bins;        /* 256 */
  int32_t n_mel;         /* 80 */
  int32_t pn_dim;        /* 512 */
  int32_t pn_kernel;     /* 5 */
  int32_t pn_layers;     /* 5 */
  int32_t max_seq_len;   /* 1000 */
  int32_t pitch_phoneme_level;  /* 0 = frame_level (LJSpeech), 1 = phoneme_level (modules.py:117-121) */
  int32_t energy_phoneme_level; /* 0 = frame_level, 1 = phoneme_level (modules.py:122-126) */
} fs2_dims;

/* One entry of the reference `state_dict` (key list: SURVEY.md section 8(b)). fp32 data. */
typedef struct fs2_weight_desc {
  const char* name;   /* e.g. "mel_decoder.layer_stack.0.pos_ffn.w_1.weight" */
  const void* data;   /* fp32, contiguous */
  int32_t ndim;
  int64_t shape[4];
  int32_t on_device;  /* 1 = device pointer on the handle's device, 0 = host pointer */
} fs2_weight_desc;

/* ---- lifetime ---------------------------------------------------------------------- */
/* replaces: FastSpeech2Align.__init__ (model/fastspeech2_align.py:16-28) */
int fs2_create(fs2_handle** out, const fs2_dims* dims, int device);
void fs2_destroy(fs2_handle* h);
const char* fs2_last_error(const fs2_handle* h); /* h may be NULL: last creation error */
const char* fs2_version(void);

/* replaces: load_state_dict in utils/model.py:16-22.  Copies + repacks (QKV concat,
 * BatchNorm fold, conv weight -> per-tap K-major bf16 tiles).  `mel_encoder.*` is optional (kept for
 * fs2_op_mel_encoder, packed on its first call); `*.num_batches_tracked` is accepted and ignored.  Missing inference
 * keys -> error. */
int fs2_load_weights(fs2_handle* h, const fs2_weight_desc* descs, int32_t n);

/* Lets `h` run on the packed weights `src` already loaded instead of its own copy (same device, same dims).  The
 * weights are read-only after fs2_load_weights, so the handles of several CUDA streams -- the reference's
 * `for batch in batchs` loop (synthesize.py:59-76) run concurrently -- share ONE copy (~300 MB with every operand
 * format) instead of one per stream.  The block is reference-counted: it is freed when the last handle that uses it
 * is destroyed or loads / shares other weights.  Workspaces stay private to each handle. */
int fs2_share_weights(fs2_handle* h, const fs2_handle* src);

/* encoder_prec covers txt_encoder + all three variance predictors (the discrete
 * decisions: durations, pitch/energy buckets); decoder_prec covers mel_decoder,
 * mel_linear and PostNet.  Defaults: F16X2 / BF16. */
int fs2_set_precision(fs2_handle* h, int32_t encoder_prec, int32_t decoder_prec);

/* Row packing of the internal activation layout.  keep_rows = padded rows kept after each utterance's valid rows:
 * 2 (default) is the minimum that reproduces the reference's padded-grid convolutions exactly (SURVEY.md section 8(a)
 * note 1); any value >= the longest utterance stores the reference's full padded [B, S_max] grid.  Results are
 * identical for every legal value; only the amount of padded work changes. */
int fs2_set_row_packing(fs2_handle* h, int32_t keep_rows);

/* Layout of the mel_post output of fs2_forward_stage2 / fs2_op_mel_postnet.  0 (default): [B, T, n_mel] as the reference
 * returns it.  1: channel-major [B, n_mel, T] -- the tensor `predictions[1].transpose(1, 2)` that the reference hands to
 * the vocoder (utils/tools.py:191, utils/model.py:70-76), written directly by the PostNet's last convolution (its
 * epilogue holds one row per thread, so every channel is one coalesced store) instead of by a later transposing copy.
 * `mel` (the pre-PostNet output) always stays [B, T, n_mel]. */
int fs2_set_mel_post_layout(fs2_handle* h, int32_t channel_major);

/* Which operator expands phoneme rows to frame rows in fs2_forward_stage2.  HARD (default): LengthRegulator
 * (modules.py:195-230), what the reference's VarianceAdaptor instantiates (modules.py:22,129,136).  GAUSSIAN:
 * GaussianUpsampling (modules.py:162-192; the operator the reference's README.md:10 announces but never wires in) applied
 * to the same rounded durations: frame t = sum_i w[i,t] x[i], w = softmax-like Gaussian weights around the phoneme centres,
 * over ALL L phoneme slots of the padded batch (no masking, as in the reference class); mel_lens, masks and T are those of
 * the hard regulator (s_b = sum d_b, integer durations).  Everything downstream is unchanged. */
#define FS2_UPSAMPLER_HARD 0
#define FS2_UPSAMPLER_GAUSSIAN 1
int fs2_set_upsampler(fs2_handle* h, int32_t upsampler);

/* ---- the forward, in two stages because T = max(sum(durations)) is data dependent --- */
/* replaces: fastspeech2_align.py:46-53 (src mask, TxtEncoder) + modules.py:116-135
 * (duration predictor, rounding) + the length bookkeeping of LengthRegulator.LR
 * (modules.py:201-218).
 *   texts[B,L] int64 (0 = PAD), src_lens[B] int64
 *   out: log_d[B,L], d_rounded[B,L] (fp32 integers, bit-exact gate), mel_lens[B] int64,
 *        src_mask[B,L] uint8, *T_max_out (host int) = max_b mel_lens[b]
 *   pitch_ph / energy_ph [B,L]: only written when the corresponding *_phoneme_level dim
 *   is 1 (may be NULL otherwise). */
int fs2_forward_stage1(fs2_handle* h, const int64_t* texts, const int64_t* src_lens, int32_t B, int32_t L,
                       float p_control, float e_control, float d_control, float* log_d, float* d_rounded,
                       int64_t* mel_lens, uint8_t* src_mask, float* pitch_ph, float* energy_ph,
                       int32_t* T_max_out, void* stream);

/* Multi-GPU variant of stage 1 (same work, same outputs, NO host synchronisation): tmax_dev[2] (device int32, caller
 * owned) receives {max_b mel_lens[b], sum_b mel_lens[b]} on `stream`.  The caller reduces tmax_dev[0] across the ranks
 * of its utterance shards on the same stream (all-reduce MAX: the only exchange step of the path), reads the two ints
 * back ONCE and passes them to fs2_forward_stage1_commit before fs2_forward_stage2(h, T, ...).  One synchronisation
 * per forward instead of two. */
int fs2_forward_stage1_async(fs2_handle* h, const int64_t* texts, const int64_t* src_lens, int32_t B, int32_t L,
                             float p_control, float e_control, float d_control, float* log_d, float* d_rounded,
                             int64_t* mel_lens, uint8_t* src_mask, float* pitch_ph, float* energy_ph,
                             int32_t* tmax_dev, void* stream);
/* T_max: the (possibly batch-global) maximum mel length stage 2 will be called with; frames: this rank's sum of
 * mel_lens (sizes tile-shape decisions only; 0 = unknown). */
int fs2_forward_stage1_commit(fs2_handle* h, int32_t T_max, int32_t frames);

/* sum_b mel_lens[b] of the last completed stage 1 on this handle (read back together with T_max; the caller sizes packed
 * result buffers with it, fs2_pack_valid_rows); -1 before the first stage 1. */
int64_t fs2_last_frame_count(const fs2_handle* h);

/* replaces: LengthRegulator expand/pad (modules.py:220-226, utils/tools.py:288-306),
 * frame-level pitch/energy (modules.py:137-149), MelDecoder (Models.py:212-244),
 * mel_linear + PostNet + residual (fastspeech2_align.py:82-85).
 *   T >= the T_max of stage 1 (a larger T = the batch-global maximum when the batch is
 *   sharded across GPUs; outputs are then padded exactly like the unsharded reference).
 *   out: mel[B,T,n_mel], mel_post[B,T,n_mel], pitch[B,T], energy[B,T] (NULL allowed when
 *   phoneme-level), mel_mask[B,T] uint8. */
int fs2_forward_stage2(fs2_handle* h, int32_t T, float p_control, float e_control, float* mel, float* mel_post,
                       float* pitch, float* energy, uint8_t* mel_mask, void* stream);

/* ---- the same two stages through the handle's CUDA-graph cache ------------------------------------------------------
 * replaces: the per-batch launch sequence of the reference's driver loop (synthesize.py:59-76) -- ~30 + ~38 kernel
 * launches per forward -- by two cudaGraphLaunch calls.  Graphs are keyed on
 * (B, L bucket) for stage 1 and (B, L bucket, T bucket) for stage 2 plus every pointer and control value baked into the
 * launches, so the caller passes L_cap / T_cap = the upper bound of the bucket the true L / T falls in (any policy: the
 * library only needs L <= L_cap, T <= T_cap) and REUSES its output buffers from call to call.  The true L / T travel
 * through device memory and every kernel reads them from there, so the results are bit-identical to
 * fs2_forward_stage1 / fs2_forward_stage2 on the exact shapes: a bucket only bounds grid and workspace sizes.
 *   - outputs: caller-owned buffers with room for the bucket ([B, L_cap] / [B, T_cap, n_mel] / ... elements); they are
 *     written DENSELY with the true shape ([B, L] / [B, T, n_mel]: the first B*L / B*T*n_mel elements).
 *   - texts [B, L] / src_lens [B] may change from call to call (they are copied to staging buffers first).
 *   - first call with a key: plain launches (sizes the workspace); second: stream capture + instantiate; then replays.
 *     Growth of the workspace, new weights or a changed setting invalidate the cached graphs (re-captured on demand).
 *   - a bucket never straddles max_seq_len (the positional table differs on either side): L_cap / T_cap are clamped.
 * The only blocking call is still the 8-byte read-back of {T_max, frames} at the end of stage 1. */
int fs2_forward_stage1_graph(fs2_handle* h, const int64_t* texts, const int64_t* src_lens, int32_t B, int32_t L,
                             int32_t L_cap, float p_control, float e_control, float d_control, float* log_d,
                             float* d_rounded, int64_t* mel_lens, uint8_t* src_mask, float* pitch_ph, float* energy_ph,
                             int32_t* T_max_out, void* stream);
int fs2_forward_stage2_graph(fs2_handle* h, int32_t T, int32_t T_cap, float p_control, float e_control, float* mel,
                             float* mel_post, float* pitch, float* energy, uint8_t* mel_mask, void* stream);
/* graphs currently instantiated on this handle, graph launches and captures since creation (any pointer may be NULL) */
int fs2_graph_stats(const fs2_handle* h, int64_t* n_graphs, int64_t* replays, int64_t* captures);

/* ---- stand-alone operators (no weights) -------------------------------------------- */
/* modules.py:132-135: out = clamp(round(exp(log_d) - 1) * d_control, min=0); n elements */
int fs2_round_durations(const float* log_d, int64_t n, float d_control, float* out, void* stream);
/* modules.py:206-218 bookkeeping: cum[B,L] = inclusive cumsum of max(int(d),0); mel_lens[B];
 * *T_max_out (host) = max mel_lens.  Blocks on a 4-byte D2H. */
int fs2_duration_scan(const float* d, int32_t B, int32_t L, int32_t* cum, int64_t* mel_lens, int32_t* T_max_out,
                      void* stream);
/* modules.py:220-226 + tools.py:288-306: out[b,t,:] = x[b,i,:] for cum[b,i-1] <= t < cum[b,i], 0 for t >= mel_len */
int fs2_length_regulate(const float* x, const int32_t* cum, int32_t B, int32_t L, int32_t D, int32_t T, float* out,
                        void* stream);
/* modules.py:166-192 GaussianUpsampling (sigma fixed to 10.0 as in :175).  d[B,L] fp32 durations,
 * T = number of output frames (>= max sum d; frames beyond max-sum are zero like `pad`),
 * T_w = int(max_b sum d) = extent of the weight tensor.  out[B,T,D]; s[B]; w[B,L,T_w] or NULL. */
int fs2_gaussian_upsample(const float* x, const float* d, int32_t B, int32_t L, int32_t D, int32_t T, int32_t T_w,
                          float* out, float* s, float* w, void* stream);
/* utils/tools.py:89-97: mask[b,i] = i >= lens[b] */
int fs2_mask_from_lengths(const int64_t* lens, int32_t B, int32_t max_len, uint8_t* mask, void* stream);

/* ---- hand-off of the results (no handle, no weights) ---------------------------- */
/* replaces: the per-utterance `.item()` + slice + `.cpu()` loop of synth_samples (utils/tools.py:156-171).
 * src [B,S,C] fp32 (channel_major = 1: [B,C,S]); lens[B] int64, clamped to [0,S].  dst receives the valid rows of
 * utterance 0, 1, ... back to back: rows [offsets[b], offsets[b+1]) of a [sum lens, C] matrix (channel_major: utterance
 * b's block, dst + offsets[b]*C, is a [C, lens[b]] matrix).  offsets[B+1] int64 (device, may be NULL) = exclusive prefix
 * sum of the clamped lens.  dst must hold sum(lens)*C floats (B*S*C always suffices).  S == 0 (the degenerate T == 0
 * batch): nothing is copied, offsets are all 0, src / dst may be NULL. */
int fs2_pack_valid_rows(const float* src, const int64_t* lens, int32_t B, int32_t S, int32_t C, int32_t channel_major,
                        int64_t* offsets, float* dst, void* stream);
/* replaces: `(wavs.cpu().numpy() * max_wav_value).astype("int16")` + `wavs[i][:lengths[i]]` (utils/model.py:77-86).
 * wav [B,N] fp32 (the vocoder's output); lens[B] int64 samples kept per row (clamped to [0,N]; NULL = all N).
 * dst int16: the kept samples back to back, offsets[B+1] as above (may be NULL).  Conversion = numpy's on x86-64:
 * fp32 product, truncation toward zero to int32, low 16 bits (NaN / |x| >= 2^31 -> 0). */
int fs2_wav_to_int16(const float* wav, const int64_t* lens, int32_t B, int64_t N, float max_wav_value, int64_t* offsets,
                     int16_t* dst, void* stream);

/* ---- per-operator entry points on a loaded handle (unit parity tests) --------------- */
/* Models.py:10-30: table[n_pos, d_model] as the reference builds it (float64 -> float32) */
int fs2_op_sinusoid_table(fs2_handle* h, int32_t n_pos, float* out, void* stream);
/* Models.py:82-91: out[B,L,D] = src_word_emb[texts] + PE[:L] */
int fs2_op_embed_pe(fs2_handle* h, const int64_t* texts, int32_t B, int32_t L, float* out, void* stream);
/* Layers.py:39-48 x (layer_end-layer_begin): stack 0 = txt_encoder, 1 = mel_decoder. x,out [B,S,D] */
int fs2_op_fft_stack(fs2_handle* h, int32_t stack, int32_t layer_begin, int32_t layer_end, int32_t prec,
                     const float* x, const int64_t* lens, int32_t B, int32_t S, float* out, void* stream);
/* modules.py:278-286: which 0 = duration, 1 = pitch, 2 = energy. x[B,S,D] -> out[B,S] */
int fs2_op_variance_predictor(fs2_handle* h, int32_t which, const float* x, const int64_t* lens, int32_t B,
                              int32_t S, float* out, void* stream);
/* modules.py:80-100 inference branch: pred <- pred*control (in place); idx = bucketize(pred, bins);
 * x[B,S,D] += embedding[idx]; idx_out[B,S] int32 optional. which 1 = pitch, 2 = energy */
int fs2_op_variance_embed(fs2_handle* h, int32_t which, float* pred, float control, float* x, int32_t B, int32_t S,
                          int32_t* idx_out, void* stream);
/* fastspeech2_align.py:83-85: mel = mel_linear(dec); mel_post = PostNet(mel) + mel. dec[B,T,D] */
int fs2_op_mel_postnet(fs2_handle* h, int32_t prec, const float* dec, int32_t B, int32_t T, float* mel,
                       float* mel_post, void* stream);
/* Training-side aligner: transformer/Models.py:140-173 MelEncoder.forward in eval mode --
 * Prenet (Layers.py:15-28) on the mels with frame 0 replaced by zeros, + positional table, n_dec_layers x FFTBlock2
 * (Layers.py:51-70: cross-attention with queries = mel frames and keys = values = src_seq, then the conv FFN).  This is the
 * forward only (what `fastspeech2_align.py:56` calls; the reference's own training branch stops right after it on an
 * undefined `_calculate_duration`).
 *   src_seq [B,L,d_model] (TxtEncoder output), mels [B,T,n_mel], src_lens / mel_lens [B] int64 (valid rows; the masks).
 *   out [B,T,d_model] (rows >= mel_lens[b] are zero); attn (may be NULL): [n_dec_layers, B, n_heads, T, L] softmax
 *   probabilities, the `dec_crs_attn_list` the reference returns (per layer a [B, H, T, L] tensor, SubLayers.py:47-48).
 *   prec selects the GEMM arithmetic (FS2_PREC_*); the attention itself runs in fp32.
 * Needs the `mel_encoder.*` keys in the loaded state_dict (FS2_ERR_MISSING_WEIGHT otherwise). */
int fs2_op_mel_encoder(fs2_handle* h, int32_t prec, const float* src_seq, const float* mels, const int64_t* src_lens,
                       const int64_t* mel_lens, int32_t B, int32_t L, int32_t T, float* out, float* attn, void* stream);
/* Raw conv-as-GEMM check: out[R,N] = act(sum_t A[r+t-pad,:] . W[:, :, t]^T + bias), rows laid out as
 * B utterances of S rows (zero padded outside [0,S)).  W is torch Conv1d layout [N,K,taps].
 * act: 0 none, 1 relu, 2 tanh.  prec selects the SIMT fp32 kernel or a tcgen05 mode (FS2_PREC_*). */
int fs2_op_conv_gemm(int32_t prec, const float* A, const float* W, const float* bias, int32_t B, int32_t S,
                     int32_t K, int32_t N, int32_t taps, int32_t act, float* out, void* stream);
/* Modules.py:14-25 on packed heads: q,k,v,out [B,S,H*dk]; keys >= lens[b] masked; rows >= lens[b] zero */
int fs2_op_attention(int32_t prec, const float* q, const float* k, const float* v, const int64_t* lens, int32_t B,
                     int32_t S, int32_t H, int32_t dk, float* out, void* stream);

/* ---- tracing (new; the reference has none) ------------------------ */
/* When enabled every kernel class of the forward ("dec.ffn_w1", "enc.attn", "postnet.2", ...) is
 * bracketed by CUDA events on the launching stream.  fs2_profile_read synchronises the device and
 * returns accumulated device milliseconds and launch counts per class since the last reset.
 * Events perturb back-to-back launches (they break the programmatic overlap of consecutive kernels: ~9 % on the sum at
 * batch 256): time whole steps with tracing OFF.  on = 2 brackets only whole segments ("enc.fft_stack", "dec.fft_stack",
 * "mel_postnet": all layers of a stack between ONE pair of events), which leaves the launches inside a segment untouched. */
typedef struct fs2_profile_entry {
  char name[48];
  int64_t launches;
  double ms;
} fs2_profile_entry;
int fs2_profile_enable(fs2_handle* h, int32_t on);
int fs2_profile_reset(fs2_handle* h);
int fs2_profile_read(fs2_handle* h, fs2_profile_entry* out, int32_t max_entries, int32_t* n_out);

/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t fs2_launch_count(const fs2_handle* h);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FS2_B200_H */
