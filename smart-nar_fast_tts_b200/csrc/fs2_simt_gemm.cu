// fs2_simt_gemm.cu -- fp32 FFMA implicit-GEMM Conv1d / Linear with fused epilogues (sm_100a).
//
// This is the fp32-faithful arithmetic of the path: it computes every Linear / Conv1d of
//   MultiHeadAttention (transformer/SubLayers.py:39-41,56), PositionwiseFeedForward (:89-93),
//   VariancePredictor (model/modules.py:278-286), mel_linear (fastspeech2_align.py:83) and
//   PostNet (transformer/Layers.py:169-177)
// as  out[r, n] = epi( sum_t sum_k A[r + t - pad, k] * W[t][k][n] + bias[n] )  over the flat
// halo'ed row grid described in fs2_common.cuh, with bias / ReLU / tanh / residual /
// LayerNorm / row-mask / final dot fused into the epilogue (LayerNorm statistics by warp
// shuffle: one warp owns 8 full rows of 256 channels).
//
// It decides the discrete outputs (durations, pitch/energy buckets), so it is plain
// IEEE fp32 multiply-add with fp32 accumulation; the tcgen05 kernels (fs2_tc_gemm.cu)
// carry the bulk FLOPs of the decoder.
#include "fs2_common.cuh"

namespace {

constexpr int BK = 16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int BM, int BN>
__global__ void __launch_bounds__(256) simt_conv_gemm_kernel(const ConvGemmArgs a) {
  FS2_PDL_PROLOGUE();
  constexpr int TX = BN / 8;   // threads along N (each owns 2 x 4 columns)
  constexpr int TY = 256 / TX; // threads along M (each owns 8 rows)
  static_assert(TY * 8 == BM, "tile/thread mismatch");
  constexpr int AS = BM + 4;
  constexpr int A_LD = BM * BK / 4 / 256;  // float4 loads of A per thread
  constexpr int B_LD = BK * BN / 4 / 256;  // float4 loads of B per thread
  __shared__ __align__(16) float As[2][BK][AS];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int R = ld_act(a.lay.off + a.lay.B);   // rows in use (device data); rows beyond are never read as non-zero
  const int r0 = blockIdx.x * BM;
  if (r0 >= R) return;
  const int n0 = blockIdx.y * BN;
  const int pad = (a.taps - 1) / 2;
  const int KB = a.K / BK;
  const int iters = a.taps * KB;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[A_LD], rb[B_LD];

  auto load_global = [&](int it) {
    const int t = it / KB;
    const int k0 = (it - t * KB) * BK;
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 2, kq = idx & 3;
      const int rr = r0 + row + t - pad;
      if (rr >= 0 && rr < R)
        ra[i] = ld_act(reinterpret_cast<const float4*>(a.A + (size_t)rr * a.K + k0 + kq * 4));
      else
        ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * 256;
      const int kk = idx / (BN / 4), nq = idx % (BN / 4);
      const int n = n0 + nq * 4;
      if (n < a.N)
        rb[i] = __ldg(reinterpret_cast<const float4*>(a.Wf + ((size_t)t * a.K + k0 + kk) * a.N + n));
      else
        rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 2, kq = idx & 3;
      As[buf][kq * 4 + 0][row] = ra[i].x;
      As[buf][kq * 4 + 1][row] = ra[i].y;
      As[buf][kq * 4 + 2][row] = ra[i].z;
      As[buf][kq * 4 + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * 256;
      const int kk = idx / (BN / 4), nq = idx % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[i];
    }
  };

  load_global(0);
  store_smem(0);
  __syncthreads();

  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) load_global(it + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (it + 1 < iters) {
      store_smem(buf ^ 1);
      __syncthreads();
    }
  }

  // ---------------------------------------------------------------- epilogue
  const int nA = n0 + tx * 4;           // columns nA..nA+3
  const int nB = n0 + BN / 2 + tx * 4;  // columns nB..nB+3
  float bias[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bias[j] = (nA + j < a.N) ? __ldg(a.bias + nA + j) : 0.f;
    bias[4 + j] = (nB + j < a.N) ? __ldg(a.bias + nB + j) : 0.f;
  }
  const bool ln_mode = (a.epi == EPI_RES_LN || a.epi == EPI_RELU_LN || a.epi == EPI_RELU_LN_DOT);
  float g[8], be[8], dw[8];
  if (ln_mode) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      g[j] = __ldg(a.ln_g + nA + j);
      g[4 + j] = __ldg(a.ln_g + nB + j);
      be[j] = __ldg(a.ln_b + nA + j);
      be[4 + j] = __ldg(a.ln_b + nB + j);
      if (a.epi == EPI_RELU_LN_DOT) {
        dw[j] = __ldg(a.dot_w + nA + j);
        dw[4 + j] = __ldg(a.dot_w + nB + j);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty * 8 + i;
    const bool in_buf = r < R;   // warp-uniform (a warp owns consecutive rows of one ty)
    const RowPos rp = row_pos(a.lay, r, R);
    const int b = rp.b, p = rp.p;
    const bool in_grid = rp.in_grid;
    const bool keep_len = in_grid && (a.lay.lens == nullptr || p < ld_act(a.lay.lens + b));
    const bool keep = (a.mask_mode == MASK_LEN) ? keep_len : in_grid;
    const bool dst_ok = a.dst_off ? in_grid : in_buf;
    const size_t dst_r = a.dst_off ? (size_t)(ld_act(a.dst_off + b) + p) : (size_t)r;

    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[i][j] + bias[j];

    if (a.epi == EPI_RES_LN || a.epi == EPI_RES) {
      if (in_buf) {
        if (nA < a.N) {
          const float4 q = ld_act(reinterpret_cast<const float4*>(a.residual + (size_t)r * a.N + nA));
          v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
        }
        if (nB < a.N) {
          const float4 q = ld_act(reinterpret_cast<const float4*>(a.residual + (size_t)r * a.N + nB));
          v[4] += q.x; v[5] += q.y; v[6] += q.z; v[7] += q.w;
        }
      }
    }
    if (a.epi == EPI_RELU || a.epi == EPI_RELU_LN || a.epi == EPI_RELU_LN_DOT) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (a.epi == EPI_TANH) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tanhf(v[j]);
    }
    if (ln_mode) {
      // LayerNorm over the 256 channels of the row (one warp holds the whole row); two-pass variance
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
      s = warp_sum(s);
      const float mean = s * (1.0f / 256.0f);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float dlt = v[j] - mean; q = fmaf(dlt, dlt, q); }
      q = warp_sum(q);
      const float rstd = 1.0f / sqrtf(q * (1.0f / 256.0f) + 1e-5f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * g[j] + be[j];
      if (a.epi == EPI_RELU_LN_DOT) {
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) d = fmaf(v[j], dw[j], d);
        d = warp_sum(d) + a.dot_b;
        if (tx == 0 && in_grid && a.out_user) a.out_user[(size_t)b * lay_S(a.lay) + p] = keep_len ? d : 0.0f;
        continue;
      }
    }
    if (!keep) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    if (a.out && dst_ok) {
      if (nA < a.N) *reinterpret_cast<float4*>(a.out + dst_r * a.ldo + nA) = make_float4(v[0], v[1], v[2], v[3]);
      if (nB < a.N) *reinterpret_cast<float4*>(a.out + dst_r * a.ldo + nB) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (a.out_user && a.user_cm && in_grid && (a.out_user_B <= 0 || b < a.out_user_B)) {
      const int Su = lay_S(a.lay);
      float* o = a.out_user + (size_t)b * a.N * Su + p;   // channel-major [B, N, S]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (nA < a.N) o[(size_t)(nA + j) * Su] = v[j];
        if (nB < a.N) o[(size_t)(nB + j) * Su] = v[4 + j];
      }
    } else if (a.out_user && in_grid && (a.out_user_B <= 0 || b < a.out_user_B)) {
      float* o = a.out_user + ((size_t)b * lay_S(a.lay) + p) * a.ldu;
      if (nA < a.N) *reinterpret_cast<float4*>(o + nA) = make_float4(v[0], v[1], v[2], v[3]);
      if (nB < a.N) *reinterpret_cast<float4*>(o + nB) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

}  // namespace

cudaError_t simt_conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t st) {
  const int R = a.lay.R_cap;
  if (R <= 0) return cudaSuccess;
  if (!a.lay.off || !a.lay.rowmap) return cudaErrorInvalidValue;
  if (a.K % BK != 0 || a.N % 4 != 0) return cudaErrorInvalidValue;
  const bool ln = (a.epi == EPI_RES_LN || a.epi == EPI_RELU_LN || a.epi == EPI_RELU_LN_DOT);
  if (ln && a.N != 256) return cudaErrorInvalidValue;
  if (ln || a.N % 256 == 0) {
    dim3 grid((R + 63) / 64, (a.N + 255) / 256);
    (void)FS2_LAUNCH((simt_conv_gemm_kernel<64, 256>), grid, 256, 0, st, a);
  } else {
    dim3 grid((R + 127) / 128, (a.N + 127) / 128);
    (void)FS2_LAUNCH((simt_conv_gemm_kernel<128, 128>), grid, 256, 0, st, a);
  }
  ++g_fs2_launches;
  return cudaGetLastError();
}
