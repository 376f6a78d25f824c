// fs2_simt_attn.cu -- fp32 flash-style scaled-dot-product attention (sm_100a, FFMA path).
//
// Restates transformer/Modules.py:14-25 + the head split/merge of SubLayers.py:39-55:
//   out[b, q, h*dk:(h+1)*dk] = softmax_k( Q_h[b,q,:] . K_h[b,k,:] / sqrt(dk)  masked(k >= len_b -> -inf) ) @ V_h[b]
// without materialising the [H*B, S, S] matrix the reference builds (its callers discard it:
// Models.py:97,241 return_attns=False).  Online softmax over 64-key tiles; one CTA = 32
// queries of one (utterance, head).  Query rows >= len_b are written as zeros (the caller
// masks them anyway, Layers.py:43).  Used for the fp32-faithful encoder and for fp32 mode.
#include "fs2_common.cuh"
#include <math.h>

namespace {

constexpr int BQ = 32;
constexpr int BKV = 64;

template <int DK>
__global__ void __launch_bounds__(128) simt_attention_kernel(const float* qkv, int ldqkv, int q_off,
                                                             int k_off, int v_off, const RowLayout lay,
                                                             float* out, int ldo, float temperature) {
  FS2_PDL_PROLOGUE();
  constexpr int QS = DK + 4;     // padded row stride (floats) of Q/K tiles: conflict-free float4 column walks
  constexpr int PS = BKV + 4;
  constexpr int DC = DK / 64;    // float4 output column groups per thread
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                 // [BQ][QS]
  float* Ks = Qs + BQ * QS;         // [BKV][QS]
  float* Vs = Ks + BKV * QS;        // [BKV][DK]
  float* Ps = Vs + BKV * DK;        // [BQ][PS]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // ty in [0,8)
  const int b = blockIdx.z, h = blockIdx.y;
  const int p0 = blockIdx.x * BQ;
  const size_t row0 = (size_t)ld_act(lay.off + b);
  const int SA = ld_act(lay.off + b + 1) - (int)row0;   // this utterance's rows (grid + halo)
  const int len = min(ld_act(lay.lens + b), ld_act(lay.ext + b));
  if (p0 >= SA) return;

  if (p0 >= len) {  // whole tile is padding: zeros
    for (int idx = tid; idx < BQ * (DK / 4); idx += 128) {
      const int q = idx / (DK / 4), c = idx % (DK / 4);
      if (p0 + q < SA)
        *reinterpret_cast<float4*>(out + (row0 + p0 + q) * ldo + h * DK + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }

  // Q tile
  for (int idx = tid; idx < BQ * (DK / 4); idx += 128) {
    const int q = idx / (DK / 4), c = idx % (DK / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p0 + q < len) v = ld_act(reinterpret_cast<const float4*>(qkv + (row0 + p0 + q) * ldqkv + q_off + h * DK + c * 4));
    *reinterpret_cast<float4*>(Qs + q * QS + c * 4) = v;
  }

  float m[4], l[4], o[4][DC * 4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < DC * 4; ++c) o[i][c] = 0.f;
  }

  for (int k0 = 0; k0 < len; k0 += BKV) {
    __syncthreads();  // previous tile fully consumed (and Q tile visible on first pass)
    for (int idx = tid; idx < BKV * (DK / 4); idx += 128) {
      const int kr = idx / (DK / 4), c = idx % (DK / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + kr < len) {
        const float* base = qkv + (row0 + k0 + kr) * ldqkv + h * DK + c * 4;
        kv = ld_act(reinterpret_cast<const float4*>(base + k_off));
        vv = ld_act(reinterpret_cast<const float4*>(base + v_off));
      }
      *reinterpret_cast<float4*>(Ks + kr * QS + c * 4) = kv;
      *reinterpret_cast<float4*>(Vs + kr * DK + c * 4) = vv;
    }
    __syncthreads();

    // scores: queries ty + 8*i, keys tx + 16*j
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < DK; d += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (ty + 8 * i) * QS + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * QS + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }
    // scale, key mask, online softmax (row statistics shared by the 16 tx lanes of a half-warp)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = k0 + tx + 16 * j;
        s[i][j] = (key < len) ? s[i][j] / temperature : -INFINITY;
        tmax = fmaxf(tmax, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
      const float m_new = fmaxf(m[i], tmax);  // finite: every processed tile holds >= 1 valid key
      const float alpha = expf(m[i] - m_new);
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pv = expf(s[i][j] - m_new);
        psum += pv;
        Ps[(ty + 8 * i) * PS + tx + 16 * j] = pv;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
      l[i] = l[i] * alpha + psum;
      m[i] = m_new;
#pragma unroll
      for (int c = 0; c < DC * 4; ++c) o[i][c] *= alpha;
    }
    __syncthreads();

    // O += P @ V : queries ty + 8*i, output columns tx*4 + 64*c .. +3
#pragma unroll 2
    for (int j = 0; j < BKV; j += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (ty + 8 * i) * PS + j);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          const float4 vv = *reinterpret_cast<const float4*>(Vs + (j + jj) * DK + c * 64 + tx * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float pw = jj == 0 ? pv[i].x : jj == 1 ? pv[i].y : jj == 2 ? pv[i].z : pv[i].w;
            o[i][c * 4 + 0] = fmaf(pw, vv.x, o[i][c * 4 + 0]);
            o[i][c * 4 + 1] = fmaf(pw, vv.y, o[i][c * 4 + 1]);
            o[i][c * 4 + 2] = fmaf(pw, vv.z, o[i][c * 4 + 2]);
            o[i][c * 4 + 3] = fmaf(pw, vv.w, o[i][c * 4 + 3]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + 8 * i;
    if (p >= SA) continue;
    const bool valid = p < len;
    const float inv = valid ? 1.0f / l[i] : 0.f;
#pragma unroll
    for (int c = 0; c < DC; ++c) {
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) r = make_float4(o[i][c * 4] * inv, o[i][c * 4 + 1] * inv, o[i][c * 4 + 2] * inv, o[i][c * 4 + 3] * inv);
      *reinterpret_cast<float4*>(out + (row0 + p) * ldo + h * DK + c * 64 + tx * 4) = r;
    }
  }
}

template <int DK>
cudaError_t launch(const float* qkv, int ldqkv, int q_off, int k_off, int v_off, const RowLayout& lay, int H, float* out,
                   int ldo, cudaStream_t st) {
  const size_t smem = sizeof(float) * (BQ * (DK + 4) + BKV * (DK + 4) + BKV * DK + BQ * (BKV + 4));
  static std::atomic<bool> configured{false};   // handles on several host threads may race here: benign, but formally atomic
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(simt_attention_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((FS2_ROWS_PER_UTT(lay.S, FS2_HALO) + BQ - 1) / BQ, H, lay.B);
  (void)FS2_LAUNCH((simt_attention_kernel<DK>), grid, 128, smem, st, qkv, ldqkv, q_off, k_off, v_off, lay, out, ldo,
                                                      (float)sqrt((double)DK));
  ++g_fs2_launches;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Cross-attention of the training-side aligner (transformer/Layers.py:51-70 FFTBlock2 -> SubLayers.py:29-59 with
// q = mel frames, k = v = phonemes; Modules.py:14-25).  Unlike the self-attention above the [B, H, T, L] probability
// matrix IS the product here (MelEncoder returns it as the alignment, Models.py:167-171), so the kernel materialises it:
// one CTA = 16 query rows of one (utterance, head); scores of all L keys live in shared memory (L <= 2048), plain
// two-pass softmax (max, exp, sum, divide -- the reference's order of operations), then P V.  Query rows at padded
// positions are computed like valid ones (the reference masks keys only); masked keys (k >= src_len) get probability 0.
constexpr int XQ = 16;    // query rows per CTA
constexpr int XK = 64;    // keys per staged chunk

template <int DK>
__global__ void __launch_bounds__(128) simt_cross_attention_kernel(const float* q, int ldq, const float* kv, int ldkv,
                                                                   int k_off, int v_off, const RowLayout layq,
                                                                   const RowLayout layk, float temperature, float* out,
                                                                   int ldo, float* attn, int H) {
  FS2_PDL_PROLOGUE();
  constexpr int ST = DK + 4;                 // padded row stride (floats): conflict-free float4 walks
  extern __shared__ __align__(16) float xs[];
  const int b = blockIdx.z, h = blockIdx.y, p0 = blockIdx.x * XQ;
  const int Tq = ld_act(layq.ext + b);       // query rows of the utterance (the padded grid: ext = T)
  if (p0 >= Tq) return;
  const int Lk = ld_act(layk.ext + b);       // key rows stored
  const int klen = min(ld_act(layk.lens + b), Lk);
  const int Lp = (Lk + 3) & ~3;
  float* q_s = xs;                           // [XQ][ST]
  float* kv_s = q_s + XQ * ST;               // [XK][ST]
  float* s_s = kv_s + XK * ST;               // [XQ][Lp]
  const int tid = threadIdx.x;
  const size_t rq0 = (size_t)ld_act(layq.off + b) + p0, rk0 = (size_t)ld_act(layk.off + b);
  const int nq = min(XQ, Tq - p0);

  for (int idx = tid; idx < XQ * (DK / 4); idx += 128) {
    const int r = idx / (DK / 4), c = idx % (DK / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nq) v = ld_act(reinterpret_cast<const float4*>(q + (rq0 + r) * ldq + h * DK + c * 4));
    *reinterpret_cast<float4*>(q_s + r * ST + c * 4) = v;
  }
  // ---- scores: thread = 2 query rows x 4 keys of the chunk
  const int rp = tid >> 4, kg = tid & 15;
  for (int j0 = 0; j0 < Lk; j0 += XK) {
    __syncthreads();
    for (int idx = tid; idx < XK * (DK / 4); idx += 128) {
      const int jj = idx / (DK / 4), c = idx % (DK / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + jj < Lk) v = ld_act(reinterpret_cast<const float4*>(kv + (rk0 + j0 + jj) * ldkv + k_off + h * DK + c * 4));
      *reinterpret_cast<float4*>(kv_s + jj * ST + c * 4) = v;
    }
    __syncthreads();
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
    for (int d = 0; d < DK; d += 4) {
      const float4 qa = *reinterpret_cast<const float4*>(q_s + (2 * rp) * ST + d);
      const float4 qb = *reinterpret_cast<const float4*>(q_s + (2 * rp + 1) * ST + d);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 kk = *reinterpret_cast<const float4*>(kv_s + (kg * 4 + k) * ST + d);
        acc[0][k] = fmaf(qa.x, kk.x, acc[0][k]); acc[0][k] = fmaf(qa.y, kk.y, acc[0][k]);
        acc[0][k] = fmaf(qa.z, kk.z, acc[0][k]); acc[0][k] = fmaf(qa.w, kk.w, acc[0][k]);
        acc[1][k] = fmaf(qb.x, kk.x, acc[1][k]); acc[1][k] = fmaf(qb.y, kk.y, acc[1][k]);
        acc[1][k] = fmaf(qb.z, kk.z, acc[1][k]); acc[1][k] = fmaf(qb.w, kk.w, acc[1][k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + kg * 4 + k;
      if (j < Lk) {
        s_s[(2 * rp) * Lp + j] = j < klen ? acc[0][k] / temperature : -INFINITY;
        s_s[(2 * rp + 1) * Lp + j] = j < klen ? acc[1][k] / temperature : -INFINITY;
      }
    }
  }
  __syncthreads();
  // ---- softmax over the keys: 8 lanes per query row
  {
    const int r = tid >> 3, l8 = tid & 7;
    float* row = s_s + r * Lp;
    float m = -INFINITY;
    for (int j = l8; j < Lk; j += 8) m = fmaxf(m, row[j]);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    float sum = 0.f;
    for (int j = l8; j < Lk; j += 8) { const float e = expf(row[j] - m); row[j] = e; sum += e; }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    float* arow = (attn && r < nq) ? attn + (((size_t)b * H + h) * layq.S + p0 + r) * layk.S : nullptr;
    for (int j = l8; j < Lk; j += 8) {
      const float pr = row[j] / sum;      // all keys masked (src_len == 0): 0 / 0 = NaN, like the reference
      row[j] = pr;
      if (arow) arow[j] = pr;
    }
    if (arow) for (int j = Lk + l8; j < layk.S; j += 8) arow[j] = 0.f;   // key slots the (packed) layout does not store
  }
  // ---- P V: thread = 2 query rows x 8 channels
  const int cg = tid & 15;
  float o[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[0][i] = 0.f; o[1][i] = 0.f; }
  for (int j0 = 0; j0 < klen; j0 += XK) {
    __syncthreads();
    for (int idx = tid; idx < XK * (DK / 4); idx += 128) {
      const int jj = idx / (DK / 4), c = idx % (DK / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + jj < klen) v = ld_act(reinterpret_cast<const float4*>(kv + (rk0 + j0 + jj) * ldkv + v_off + h * DK + c * 4));
      *reinterpret_cast<float4*>(kv_s + jj * ST + c * 4) = v;
    }
    __syncthreads();
    const int nj = min(XK, klen - j0);
    for (int jj = 0; jj < nj; ++jj) {
      const float pa = s_s[(2 * rp) * Lp + j0 + jj], pb = s_s[(2 * rp + 1) * Lp + j0 + jj];
      const float4 va = *reinterpret_cast<const float4*>(kv_s + jj * ST + cg * 8);
      const float4 vb = *reinterpret_cast<const float4*>(kv_s + jj * ST + cg * 8 + 4);
      const float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[0][i] = fmaf(pa, vv[i], o[0][i]); o[1][i] = fmaf(pb, vv[i], o[1][i]); }
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = 2 * rp + rr;
    if (r < nq) {
      float* dst = out + (rq0 + r) * ldo + h * DK + cg * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(o[rr][0], o[rr][1], o[rr][2], o[rr][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[rr][4], o[rr][5], o[rr][6], o[rr][7]);
    }
  }
}

template <int DK>
cudaError_t launch_cross(const float* q, int ldq, const float* kv, int ldkv, int k_off, int v_off, const RowLayout& layq,
                         const RowLayout& layk, int H, float* out, int ldo, float* attn, cudaStream_t st) {
  const int Lp = (layk.S + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)(XQ + XK) * (DK + 4) + (size_t)XQ * Lp);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(simt_cross_attention_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  dim3 grid((layq.S + XQ - 1) / XQ, H, layq.B);
  (void)FS2_LAUNCH((simt_cross_attention_kernel<DK>), grid, 128, smem, st, q, ldq, kv, ldkv, k_off, v_off, layq, layk,
                   (float)sqrt((double)DK), out, ldo, attn, H);
  ++g_fs2_launches;
  return cudaGetLastError();
}

}  // namespace

cudaError_t simt_cross_attention_launch(const float* q, int ldq, const float* kv, int ldkv, int k_off, int v_off,
                                        const RowLayout& layq, const RowLayout& layk, int H, int dk, float* out, int ldo,
                                        float* attn, cudaStream_t st) {
  if (layq.B <= 0 || layq.S <= 0 || layk.S <= 0) return cudaSuccess;
  if (!layq.off || !layq.ext || !layk.off || !layk.ext || !layk.lens || layq.B != layk.B) return cudaErrorInvalidValue;
  if (dk == 128) return launch_cross<128>(q, ldq, kv, ldkv, k_off, v_off, layq, layk, H, out, ldo, attn, st);
  if (dk == 64) return launch_cross<64>(q, ldq, kv, ldkv, k_off, v_off, layq, layk, H, out, ldo, attn, st);
  return cudaErrorInvalidValue;
}

cudaError_t simt_attention_launch(const float* qkv, int ldqkv, int q_off, int k_off, int v_off, const RowLayout& lay,
                                  int H, int dk, float* out, int ldo, cudaStream_t st) {
  if (lay.B <= 0 || lay.S <= 0) return cudaSuccess;
  if (!lay.off || !lay.ext || !lay.lens) return cudaErrorInvalidValue;
  if (dk == 128) return launch<128>(qkv, ldqkv, q_off, k_off, v_off, lay, H, out, ldo, st);
  if (dk == 64) return launch<64>(qkv, ldqkv, q_off, k_off, v_off, lay, H, out, ldo, st);
  return cudaErrorInvalidValue;
}
