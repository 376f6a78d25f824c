// fs2_common.cuh -- shared definitions for the sm_100a FastSpeech2-align forward kernels.
//
// Row layout ("grid") used by every internal activation tensor
// ------------------------------------------------------------
// A reference tensor [B, S, C] is stored as R = B*SA rows of C channels, SA = S + FS2_HALO.
// Row r = b*SA + p.  Rows with p >= S (the halo) are always ZERO.  Because every Conv1d on
// the path has padding <= 4 (FFN k=9), a row-shifted read A[r + t - pad] that leaves
// [0,S) of its utterance lands in a halo row (or outside the buffer, which loaders treat
// as zero), which is exactly the zero padding Conv1d applies at the edge of the padded
// [B,S] grid in the reference (SubLayers.py:73-85, modules.py:254-272, Layers.py:120-167).
// So a convolution over the whole batch is ONE implicit GEMM over flat rows, tiles may
// straddle utterances, and no per-utterance tile quantisation is paid.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define FS2_HALO 4

typedef __nv_bfloat16 bf16;

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

struct ConvGemmArgs {
  // A operand: activations in grid layout, lda == K
  const float* A;
  const bf16* Ab;   // tcgen05 path: [planes][R][K]
  int K;
  // tcgen05 path: 1 = plain bf16 operands; 3 = bf16x3 split operands (hi, mid, lo planes of A and W): the six
  // significant cross products are accumulated in fp32, which reproduces an fp32 GEMM to ~2^-23 per product
  int planes;
  int out_planes;   // planes written to out_b (1, or 3 = split the fp32 result for a following bf16x3 GEMM)
  // W operand, per tap
  const float* Wf;  // [taps][K][N]   (SIMT fp32 kernel; n contiguous)
  const bf16* Wb;   // [planes][taps][N][K]   (tcgen05 kernel; K-major B operand)
  const float* bias;
  int N, taps;
  // grid
  int B, S, SA;
  const int* lens;  // [B] int32 (MASK_LEN, EPI_RELU_LN_DOT)
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
  // EPI_QKV extras (tcgen05 path)
  bf16* q_b;   // [R, 256]
  bf16* k_b;   // [R, 256]
  bf16* vt_b;  // [B*H*dk, SAv]  V transposed: row (b*H + h)*dk + d, column p
  int SAv;
};

#define FS2_CUDA_CHECK(expr)                                  \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return fs2_fail_cuda(_e, #expr);   \
  } while (0)

// launch wrappers implemented in the .cu files ------------------------------------------------
int fs2_fail_cuda(cudaError_t e, const char* what);  // records message, returns FS2_ERR_CUDA

cudaError_t simt_conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t st);
cudaError_t simt_attention_launch(const float* qkv, int ldqkv, int q_off, int k_off, int v_off, const int* lens, int B,
                                  int S, int SA, int H, int dk, float* out, int ldo, cudaStream_t st);

// tcgen05 path
int tc_conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t st);  // returns FS2_* code
int tc_attention_launch(const bf16* q, const bf16* k, const bf16* vt, const int* lens, int B, int S, int SA, int SAv,
                        int H, bf16* out_b, cudaStream_t st);

// row operators (fs2_rowops.cu)
cudaError_t rowops_embed_pe(const int64_t* texts, const float* emb, const float* pe, int vocab, int B, int L, int SA,
                            int D, float* out_grid, float* out_user, cudaStream_t st);
cudaError_t rowops_lens_to_i32(const int64_t* lens, int B, int cap, int* out, cudaStream_t st);
cudaError_t rowops_mask(const int64_t* lens64, const int* lens32, int B, int max_len, uint8_t* mask, cudaStream_t st);
cudaError_t rowops_round_durations(const float* log_d, int64_t n, float d_control, float* out, cudaStream_t st);
cudaError_t rowops_duration_scan(const float* d, int B, int L, int* cum, int64_t* mel_lens, int* mel_lens32,
                                 int* tmax_dev, cudaStream_t st);
cudaError_t rowops_length_regulate(const float* x, int x_row_stride_utt, const int* cum, int B, int L, int D, int T,
                                   int out_SA, float* out, cudaStream_t st);
cudaError_t rowops_variance_embed(float* pred, float control, const float* bins, int n_bins, const float* emb,
                                  const float* pe, float* x, bf16* xb, int xb_planes, int B, int S, int SA, int D,
                                  int* idx_out, cudaStream_t st);
cudaError_t rowops_gaussian_upsample(const float* x, const float* d, int B, int L, int D, int T, int T_w, float* out,
                                     float* s, float* w, cudaStream_t st);
cudaError_t rowops_to_grid(const float* x_user, int B, int S, int SA, int C, float* out, int ldo, int col_off,
                           bf16* out_b, cudaStream_t st);
cudaError_t rowops_from_grid(const float* x_grid, int B, int S, int SA, int C, float* out_user, cudaStream_t st);
// dst_b: [3][taps][n_total][K] -- plane 0 doubles as the plain bf16 weight, planes 1..2 are the split residuals
cudaError_t rowops_pack_weight(const float* src, int N, int K, int taps, const float* scale, float* dst_f,
                               bf16* dst_b, int n_total, int n_off, cudaStream_t st);
// fp32 -> bf16x3 planes: dst[p][i], plane stride = plane_elems
cudaError_t rowops_split3(const float* src, int64_t n, bf16* dst, int64_t plane_elems, cudaStream_t st);
cudaError_t rowops_bn_fold(const float* conv_bias, const float* g, const float* b, const float* mean,
                           const float* var, int n, float eps, float* scale_out, float* bias_out, cudaStream_t st);
cudaError_t rowops_f32_to_bf16(const float* src, int64_t n, bf16* dst, cudaStream_t st);
cudaError_t rowops_bf16_to_f32(const bf16* src, int64_t n, float* dst, cudaStream_t st);
cudaError_t rowops_transpose_v(const bf16* v, int B, int SA, int SAv, int D, bf16* vt, cudaStream_t st);
cudaError_t rowops_add_pe(float* x, bf16* xb, const float* pe, int B, int S, int SA, int D, cudaStream_t st);
cudaError_t rowops_fill_zero(void* p, size_t bytes, cudaStream_t st);

extern long long g_fs2_launches;  // kernels launched (incremented by every launcher)
