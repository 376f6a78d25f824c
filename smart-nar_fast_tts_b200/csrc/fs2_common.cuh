// fs2_common.cuh -- shared definitions for the sm_100a FastSpeech2-align forward kernels.
//
// Row layout ("ragged grid") used by every internal activation tensor
// ------------------------------------------------------------------
// A reference tensor [B, S, C] is stored as flat rows of C channels.  Utterance b owns rows
// [off[b], off[b+1]): first ext[b] "grid" rows (position p = 0..ext[b]-1), then >= FS2_HALO rows that
// are always ZERO (off[b] is a multiple of FS2_ROW_ALIGN).  Because every Conv1d on the path has padding <= 4 (FFN k=9), a row-shifted
// read A[r + t - pad] that leaves the utterance's grid rows lands in a zero halo row (or outside
// the buffer, which TMA / the loaders treat as zero) -- exactly the zero padding Conv1d applies
// at the edge of the reference's padded [B,S] grid (SubLayers.py:73-85, modules.py:254-272,
// Layers.py:120-167).  So a convolution over the whole batch is ONE implicit GEMM over flat rows,
// tiles straddle utterances, and no per-utterance tile quantisation is paid.
//
// ext[b] = min(lens[b] + halo_keep, S):
//   * halo_keep >= S ("uniform"): ext = S, the reference's padded grid itself (PostNet, whose
//     padded rows are returned to the caller, and the raw conv test entry point);
//   * halo_keep = 2 ("packed"): only the rows that can influence a valid output are stored.
//     FFT blocks are exact on valid rows alone (keys masked, rows zeroed after each sub-layer);
//     the two-layer k=3 variance predictors leak 2 padded rows into the last valid outputs
//     (SURVEY.md section 8(a) note 1), hence the 2 kept rows; everything past them is masked in
//     the reference as well.  Padding of a batch then costs 6 rows per utterance instead of
//     (S_max - len) rows.
// The layout lives in device memory (lens are device data; nothing is copied back to size it).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

#define FS2_HALO 4
#define FS2_ROW_ALIGN 8   // utterances start at flat rows that are multiples of 8: V^T (bf16, row index innermost) is
                          // loaded by TMA at column off[b] + j*128, and TMA needs 16-byte aligned global addresses
// rows reserved per utterance of a layout with S grid rows (upper bound of ext + halo, aligned)
#define FS2_ROWS_PER_UTT(S, halo_rows) ((halo_rows) > 0 ? (((S) + (halo_rows) + FS2_ROW_ALIGN - 1) / FS2_ROW_ALIGN) * FS2_ROW_ALIGN : (S))

typedef __nv_bfloat16 bf16;

// f16x2 operand mode (planes == 2): an fp32 value x is carried as two fp16 terms of x * 2^k, hi = fp16(x 2^k) and
// lo = fp16(x 2^k - hi): 22 significant bits, 3 cross products (hi*hi, hi*lo, lo*hi) per MAC block instead of the six
// of the bf16x3 split.  The power-of-two pre-scale keeps the lo term in fp16's NORMAL range for every value that
// matters (|x 2^k| >= 2^-3); it is undone exactly in the GEMM epilogue (ConvGemmArgs::acc_scale).  Activations use
// the fixed k = 4 (|x| up to 4094 before saturation); each weight tensor picks its own k at load time.
// The 16-bit patterns live in the same bf16-typed plane buffers as the other modes.
#define FS2_F16X2_ACT_SCALE 16.0f

enum Fs2Epi : int {
  EPI_BIAS = 0,         // out = acc + bias
  EPI_RELU = 1,         // out = relu(acc + bias)
  EPI_TANH = 2,         // out = tanh(acc + bias)
  EPI_RES_LN = 3,       // out = LN(acc + bias + residual) * g + b          (N == 256 only)
  EPI_RELU_LN = 4,      // out = LN(relu(acc + bias)) * g + b               (N == 256 only)
  EPI_RELU_LN_DOT = 5,  // scalar = dot(LN(relu(acc + bias))*g + b, dot_w) + dot_b  (N == 256 only)
  EPI_RES = 6,          // out = acc + bias + residual
  EPI_QKV = 7,          // tcgen05 path only: Q,K bf16 row-major + V transposed per (b,h)
};

enum Fs2Mask : int {
  MASK_GRID = 0,  // rows p >= S (halo) -> 0, everything else kept (padded-grid semantics: predictors, PostNet)
  MASK_LEN = 1,   // rows p >= lens[b] -> 0 (FFT blocks: masked_fill after each sub-layer, Layers.py:43-46)
};

struct RowLayout {
  int B;                   // utterances
  int S;                   // rows per utterance of the user tensor [B,S,*]: user row = b*S + p
  int R_cap;               // rows allocated = B * FS2_ROWS_PER_UTT(S, halo) (host-known upper bound of off[B])
  const int* off;          // [B+1] device: first row of utterance b; off[B] = rows in use
  const int* ext;          // [B] device: grid rows of utterance b
  const int* lens;         // [B] device: valid rows (p < lens[b]); may be null where no length mask applies
  const unsigned* rowmap;  // [R_cap] device: (b << 16) | p for grid rows, FS2_ROW_NONE for halo / unused rows
  int rows_hint;           // host-side estimate of off[B] (rows in use) for tile-shape decisions; 0 = unknown (use R_cap)
  const int* S_dev;        // non-null (CUDA-graph replays, fs2_forward_stage*_graph): the TRUE S of this forward lives in
                           // device memory and `S` above is only its bucket's upper bound, used by the host for grid and
                           // buffer sizes.  Device code must read the row count through lay_S(), never through `S`.
};
#define FS2_ROW_NONE 0xFFFFFFFFu
#define FS2_MAX_ROWS_PER_UTT 65535

struct ConvGemmArgs {
  // A operand: activations in grid layout, lda == K
  const float* A;
  const bf16* Ab;   // tcgen05 path: [planes][R][K]
  int K;
  // tcgen05 path: 1 = plain bf16 operands; 3 = bf16x3 split operands (hi, mid, lo planes of A and W): the six
  // significant cross products are accumulated in fp32, which reproduces an fp32 GEMM to ~2^-23 per product;
  // 2 = f16x2 split operands (scaled fp16 hi / lo planes, 3 cross products, see FS2_F16X2_ACT_SCALE)
  int planes;
  int out_planes;   // planes written to out_b (1; 3 / 2 = split the fp32 result for a following bf16x3 / f16x2 GEMM)
  // W operand, per tap
  const float* Wf;  // [taps][K][N]   (SIMT fp32 kernel; n contiguous)
  const bf16* Wb;   // [planes][taps][N][K]   (tcgen05 kernel; K-major B operand)
  const bf16* Wh;   // [2][taps][N][K] fp16 bit patterns of W * w_scale, hi / lo terms (f16x2 mode)
  float acc_scale;  // multiplies the accumulator before the bias: 1, or 1 / (FS2_F16X2_ACT_SCALE * w_scale) in f16x2 mode
  float acc_scale_f16x2;  // the f16x2 value for this weight (run_gemm copies it into acc_scale when that mode is used)
  const float* bias;
  int N, taps;
  // rows
  RowLayout lay;    // layout of A (and of out / out_b / residual unless dst_off is set)
  const int* dst_off;  // non-null: out / out_b live in ANOTHER ragged layout whose utterance b starts at row dst_off[b]
                       // (device): the row of (b, p) is dst_off[b] + p; only grid rows of `lay` are written
  int dst_R_cap;       // rows allocated in that layout (plane stride of out_b)
  int epi, mask_mode;
  const float* residual;  // grid layout, ld == N
  const float* ln_g;
  const float* ln_b;
  const float* dot_w;
  float dot_b;
  // outputs (any may be null)
  float* out;      // grid layout [R, ldo]
  int ldo;
  bf16* out_b;     // grid layout [out_planes][R, ldob] bf16 shadow (tcgen05 path)
  int ldob;
  float* out_user; // dense user layout: row (b,p), p < S, at (b*S + p)*ldu
  int ldu;
  int out_user_B;  // > 0: only utterances b < out_user_B exist in out_user (the layout carries extra pseudo utterances)
  int user_cm;     // 1: out_user is channel-major [B, N, S] (element (b,p,c) at (b*N + c)*S + p): the layout the vocoder
                   // takes (utils/tools.py:191 `predictions[1].transpose(1, 2)`); EPI_BIAS / RELU / TANH / RES only
  // EPI_QKV extras (tcgen05 path)
  bf16* q_b;   // [R, 256]
  bf16* k_b;   // [R, 256]
  bf16* vt_b;  // [H*dk, Rv]  V transposed: row h*dk + d, column = flat row index r
  int Rv;      // R_cap rounded up to 8 (16-byte row pitch for TMA)
};

// Loads of data written by OTHER kernels of the forward (activations, layout tables, lengths): plain coherent
// ld.global issued from volatile asm.  They must NOT be __ldg / ld.global.nc, and must not go through a
// `const T* __restrict__` kernel parameter (nvcc turns those into ld.global.nc as well): ptxas schedules non-coherent
// loads ABOVE griddepcontrol.wait (seen in SASS: LDG.E.CONSTANT issued before ACQBULK), i.e. before the producing
// kernel has finished under programmatic dependent launch -- stale layout tables were read that way (profiles/r1g
// notes).  Volatile asm keeps program order against the (volatile) griddepcontrol.wait.  Weights / biases, which no
// kernel of the forward writes, may keep __ldg.
__device__ __forceinline__ int ld_act(const int* p) {
  int v; asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ unsigned ld_act(const unsigned* p) {
  unsigned v; asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ float ld_act(const float* p) {
  float v; asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
__device__ __forceinline__ long long ld_act(const long long* p) {
  long long v; asm volatile("ld.global.b64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ float4 ld_act(const float4* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// rows per utterance of the user tensors, as device code must see it (see RowLayout::S_dev)
__device__ __forceinline__ int lay_S(const RowLayout& lay) { return lay.S_dev ? ld_act(lay.S_dev) : lay.S; }
// a scalar shape parameter (L, max_len, ...) that a graph replay reads from device memory instead of its baked-in bound
__device__ __forceinline__ int shape_or(const int* dev, int host_value) { return dev ? ld_act(dev) : host_value; }

// (b, p) of a flat row
struct RowPos {
  int b, p;
  bool in_grid;
};
__device__ __forceinline__ RowPos row_pos(const RowLayout& lay, int r, int R) {
  RowPos rp;
  const unsigned code = (r < R) ? ld_act(lay.rowmap + r) : FS2_ROW_NONE;
  rp.in_grid = code != FS2_ROW_NONE;
  rp.b = rp.in_grid ? (int)(code >> 16) : 0;
  rp.p = rp.in_grid ? (int)(code & 0xFFFFu) : 0;
  return rp;
}

// fp32 -> (hi, lo) fp16 bit patterns of y * FS2_F16X2_ACT_SCALE, saturating (never inf)
__device__ __forceinline__ uint16_t f32_to_f16_sat_bits(float v) {
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return h;
}
__device__ __forceinline__ void split2h_scaled(float ys, uint16_t& hi, uint16_t& lo) {   // ys already scaled
  hi = f32_to_f16_sat_bits(ys);
  lo = f32_to_f16_sat_bits(ys - __half2float(__ushort_as_half(hi)));
}
// two values -> packed hi word and packed lo word
__device__ __forceinline__ void split2h_pair(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  uint16_t ha, la, hb, lb;
  split2h_scaled(a * FS2_F16X2_ACT_SCALE, ha, la);
  split2h_scaled(b * FS2_F16X2_ACT_SCALE, hb, lb);
  hi2 = (uint32_t)ha | ((uint32_t)hb << 16);
  lo2 = (uint32_t)la | ((uint32_t)lb << 16);
}

// operand-plane stores shared by the row operators (fs2_rowops.cu, fs2_gaussian.cu)
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// y -> (hi, mid, lo) bf16 terms, residuals exact in fp32
__device__ __forceinline__ void split3f(float y, float& hi, float& mid, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(y));
  const float r1 = y - hi;
  mid = __bfloat162float(__float2bfloat16_rn(r1));
  lo = r1 - mid;
}
// write 4 consecutive values into `planes` operand planes (1: bf16 rounded; 3: bf16 hi/mid/lo split; 2: scaled fp16 hi/lo)
__device__ __forceinline__ void store_planes4(bf16* dst, size_t plane_elems, int planes, float4 a) {
  if (planes == 1) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack2(a.x, a.y), pack2(a.z, a.w));
    return;
  }
  if (planes == 2) {
    uint32_t h0, l0, h1, l1;
    split2h_pair(a.x, a.y, h0, l0);
    split2h_pair(a.z, a.w, h1, l1);
    *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(dst + plane_elems) = make_uint2(l0, l1);
    return;
  }
  float h[4], m[4], l[4];
  split3f(a.x, h[0], m[0], l[0]); split3f(a.y, h[1], m[1], l[1]);
  split3f(a.z, h[2], m[2], l[2]); split3f(a.w, h[3], m[3], l[3]);
  *reinterpret_cast<uint2*>(dst) = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
  *reinterpret_cast<uint2*>(dst + plane_elems) = make_uint2(pack2(m[0], m[1]), pack2(m[2], m[3]));
  *reinterpret_cast<uint2*>(dst + 2 * plane_elems) = make_uint2(pack2(l[0], l[1]), pack2(l[2], l[3]));
}

#define FS2_CUDA_CHECK(expr)                                  \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return fs2_fail_cuda(_e, #expr);   \
  } while (0)

// launch wrappers implemented in the .cu files ------------------------------------------------
int fs2_fail_cuda(cudaError_t e, const char* what);  // records message, returns FS2_ERR_CUDA

cudaError_t simt_conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t st);
cudaError_t simt_attention_launch(const float* qkv, int ldqkv, int q_off, int k_off, int v_off, const RowLayout& lay,
                                  int H, int dk, float* out, int ldo, cudaStream_t st);

// cross-attention of the training-side aligner (queries in layq, keys / values in layk; K at column k_off, V at v_off of kv);
// attn (optional): [B, H, layq.S, layk.S] probabilities, the alignment MelEncoder returns (Models.py:167-171)
cudaError_t simt_cross_attention_launch(const float* q, int ldq, const float* kv, int ldkv, int k_off, int v_off,
                                        const RowLayout& layq, const RowLayout& layk, int H, int dk, float* out, int ldo,
                                        float* attn, cudaStream_t st);

// tcgen05 path
int tc_conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t st);  // returns FS2_* code
bool tc_conv_gemm_staged_supported(const ConvGemmArgs& a);        // fs2_tc_gemm_staged.cu: TMA-staged epilogue variant
int tc_conv_gemm_staged_launch(const ConvGemmArgs& a, cudaStream_t st);
int tc_attention_launch(const bf16* q, const bf16* k, const bf16* vt, const RowLayout& lay, int Rv, int H, int planes,
                        bf16* out_b, cudaStream_t st);

// row operators (fs2_rowops.cu)
// off/ext/rowmap of `lay` from lens32 (device; null = every utterance has S rows): ext = min(lens + halo_keep, S),
// each utterance followed by halo_rows zero rows (FS2_HALO for GEMM operands, 0 for dense user tensors).
// extra_ext > 0 appends one pseudo utterance (index B) with extra_ext grid rows: off has B + 2 entries, ext B + 1.
// S_dev (optional, device): the true S when `S` is only an upper bound (graph replays)
cudaError_t rowops_build_layout(const int* lens32, int B, int S, int halo_keep, int halo_rows, int* off, int* ext,
                                unsigned* rowmap, int R_cap, cudaStream_t st, int extra_ext = 0, const int* S_dev = nullptr);
cudaError_t rowops_embed_pe(const int64_t* texts, const float* emb, const float* pe, int vocab, const RowLayout& lay,
                            int D, float* out_grid, bf16* out_b, int out_planes, float* out_user, cudaStream_t st);
// zero2 (optional): two ints cleared by the same launch (the forward's {T_max, frames} accumulators)
cudaError_t rowops_lens_to_i32(const int64_t* lens, int B, int cap, int* out, cudaStream_t st, int* zero2 = nullptr,
                               const int* cap_dev = nullptr);
// zero0 / zero1 (optional): fp32 tensors of the mask's shape cleared by the same launch
cudaError_t rowops_mask(const int64_t* lens64, const int* lens32, int B, int max_len, uint8_t* mask, cudaStream_t st,
                        float* zero0 = nullptr, float* zero1 = nullptr, const int* max_len_dev = nullptr);
cudaError_t rowops_round_durations(const float* log_d, int64_t n, float d_control, float* out, cudaStream_t st);
cudaError_t rowops_duration_scan(const float* d, int B, int L, int* cum, int64_t* mel_lens, int* mel_lens32,
                                 int* tmax_dev, cudaStream_t st);
// fused: d_rounded = clamp(round(exp(log_d) - 1) * d_control, 0) (modules.py:132-135), then the scan of d_rounded
cudaError_t rowops_round_scan(const float* log_d, float d_control, float* d_rounded, int B, int L, int* cum,
                              int64_t* mel_lens, int* mel_lens32, int* tmax_dev, cudaStream_t st, const int* L_dev = nullptr);
// x rows of utterance b start at src_off[b] (device) or b*src_stride when src_off is null; out rows follow `lay`
cudaError_t rowops_length_regulate(const float* x, const int* src_off, int src_stride, const int* cum, int L, int D,
                                   const RowLayout& lay, float* out, bf16* out_b, int out_planes, cudaStream_t st,
                                   const int* L_dev = nullptr);
cudaError_t rowops_variance_embed(float* pred, float control, const float* bins, int n_bins, const float* emb,
                                  const float* pe, float* x, bf16* xb, int xb_planes, const RowLayout& lay, int D,
                                  int* idx_out, cudaStream_t st);
// mel_linear rows that the (packed) source layout `src` does not carry, in the destination layout `dst` (PostNet grid,
// possibly with one pseudo utterance at index src.B): grid rows p in [src.ext[b], dst.ext[b]) <- bias row (mel_linear of
// a zero decoder row), the rows after them up to the next utterance <- 0, every grid row of the pseudo utterance <- bias;
// and the user tensor [B,S,N] rows p in [src.ext[b], S) <- bias
cudaError_t rowops_fill_padded_rows(const float* bias, int N, const RowLayout& src, const RowLayout& dst, float* out_grid,
                                    bf16* out_b, int out_planes, float* out_user, cudaStream_t st);
// PostNet rows farther than H from the last valid frame depend only on the bias row and on their distance to the end of
// the grid: copy them from the pseudo utterance (index B of `pn`, min(S, 2H+1) all-bias rows) into out_user [B,S,N]
cudaError_t rowops_postnet_far_rows(const float* post_grid, int N, const RowLayout& pn, int B, int H, float* out_user,
                                    int user_cm, cudaStream_t st);
cudaError_t rowops_gaussian_upsample(const float* x, const float* d, int B, int L, int D, int T, int T_w, float* out,
                                     float* s, float* w, cudaStream_t st);   // fs2_gaussian.cu
// the forward's soft length regulator: rowops_length_regulate's arguments + src_rows[b] = rows of utterance b that exist
// in x (the phonemes beyond them are the masked, all-zero ones); integer durations given by their scan `cum`
cudaError_t rowops_gaussian_regulate(const float* x, const int* src_off, const int* src_rows, const int* cum, int L, int D,
                                     const RowLayout& lay, float* out, bf16* out_b, int out_planes, cudaStream_t st,
                                     const int* L_dev = nullptr);
cudaError_t rowops_to_grid(const float* x_user, const RowLayout& lay, int C, float* out, int ldo, int col_off,
                           bf16* out_b, cudaStream_t st);
cudaError_t rowops_zero_first_rows(const RowLayout& lay, int C, float* x, cudaStream_t st);
cudaError_t rowops_from_grid(const float* x_grid, const RowLayout& lay, int C, float* out_user, cudaStream_t st);
// dst_b: [3][taps][n_total][K] -- plane 0 doubles as the plain bf16 weight, planes 1..2 are the split residuals
cudaError_t rowops_pack_weight(const float* src, int N, int K, int taps, const float* scale, float* dst_f,
                               bf16* dst_b, int n_total, int n_off, cudaStream_t st);
// fp32 -> operand planes (3: bf16x3, 2: f16x2, 1: bf16): dst[p][i], plane stride = plane_elems
cudaError_t rowops_split(const float* src, int64_t n, int planes, bf16* dst, int64_t plane_elems, cudaStream_t st);
// f16x2 weight planes from the packed fp32 weight wf [taps][K][N]: returns the power-of-two scale chosen for the
// tensor (max |w| * scale in [8192, 16384)) through *w_scale_out (host); dst [2][taps][N][K]
cudaError_t rowops_pack_weight_f16x2(const float* wf, int N, int K, int taps, bf16* dst, float* w_scale_out, cudaStream_t st);
cudaError_t rowops_bn_fold(const float* conv_bias, const float* g, const float* b, const float* mean,
                           const float* var, int n, float eps, float* scale_out, float* bias_out, cudaStream_t st);
cudaError_t rowops_f32_to_bf16(const float* src, int64_t n, bf16* dst, cudaStream_t st);
cudaError_t rowops_bf16_to_f32(const bf16* src, int64_t n, float* dst, cudaStream_t st);
cudaError_t rowops_unsplit2(const bf16* src, int64_t n, int64_t plane_elems, float* dst, cudaStream_t st);  // f16x2 planes -> fp32
cudaError_t rowops_transpose_v(const bf16* v, int R, int Rv, int D, bf16* vt, cudaStream_t st);
cudaError_t rowops_add_pe(float* x, const float* pe, const RowLayout& lay, int D, cudaStream_t st);
cudaError_t rowops_fill_zero(void* p, size_t bytes, cudaStream_t st);
// fs2_handoff.cu: valid rows of a padded [B,S,C] (channel_major: [B,C,S]) tensor back to back + offsets[B+1];
// int16 = numpy astype of wav * max_wav_value for the first lens[b] samples of each row, back to back
cudaError_t handoff_pack_valid_rows(const float* src, const int64_t* lens, int B, int S, int C, int channel_major,
                                    int64_t* offsets, float* dst, cudaStream_t st);
cudaError_t handoff_wav_to_int16(const float* wav, const int64_t* lens, int B, int64_t N, float max_wav_value,
                                 int64_t* offsets, int16_t* dst, cudaStream_t st);

extern std::atomic<long long> g_fs2_launches;  // kernels launched by the process (every launcher increments it; handles on
                                                // several host threads launch concurrently)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the library is launched with the programmatic-stream-serialization
// attribute: kernel N+1 may be scheduled (and run its private prologue: barrier init, TMEM allocation, tensor-map
// prefetch, parameter staging) while kernel N drains.  Rule that makes this safe without per-buffer reasoning:
// EVERY CTA of EVERY kernel executes griddep_wait() before its first access to global memory that any other kernel
// writes (and before it exits), so a kernel's completion implies the completion of all its predecessors and all
// read-after-write / write-after-read hazards are ordered exactly as with plain stream order.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define FS2_PDL_PROLOGUE() do { griddep_launch_dependents(); griddep_wait(); } while (0)

extern int g_fs2_pdl;  // 1: launch with the programmatic-stream-serialization attribute (default; FS2_NO_PDL=1 clears it)
// 1: the next launch is issued WITHOUT the attribute (fully stream-ordered) and clears the flag.  Set at every C-ABI entry
// point and after every non-kernel stream operation the library enqueues: programmatic overlap is only relied upon
// between two kernels of this library, never against a caller's memcpy / foreign kernel / memset that precedes them.
extern thread_local int g_fs2_plain_next;   // per host thread: handles on different threads are independent
// > 0: this host thread launches everything without the attribute (weight loading).  Per thread, like the flag above:
// toggling the process-wide g_fs2_pdl from two threads loading weights at once could leave it cleared for good.
extern thread_local int g_fs2_pdl_off;
template <typename F>
inline cudaError_t fs2_launch_cfg(dim3 grid, dim3 block, size_t smem, cudaStream_t st, F&& f, int cluster_x = 1) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_fs2_pdl && !g_fs2_plain_next && !g_fs2_pdl_off) ? 1 : 0;
  g_fs2_plain_next = 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (cluster_x > 1) {   // thread-block cluster along x
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cluster_x; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  return f(&cfg);
}
// FS2_LAUNCH(kernel, grid, block, smem_bytes, stream, args...) -> cudaError_t
#define FS2_LAUNCH(kernel, grid, block, smem, st, ...) \
  fs2_launch_cfg(grid, block, smem, st, [&](const cudaLaunchConfig_t* _c) { return cudaLaunchKernelEx(_c, kernel, __VA_ARGS__); })
#define FS2_LAUNCH_CLUSTER(cluster_x, kernel, grid, block, smem, st, ...) \
  fs2_launch_cfg(grid, block, smem, st, [&](const cudaLaunchConfig_t* _c) { return cudaLaunchKernelEx(_c, kernel, __VA_ARGS__); }, cluster_x)
