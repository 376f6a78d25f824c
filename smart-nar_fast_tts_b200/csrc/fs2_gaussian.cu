// fs2_gaussian.cu -- GaussianUpsampling (reference model/modules.py:162-192) for sm_100a.
//
//   s_b = sum_i d[b,i];  e = cumsum(d);  c_i = e_i - d_i / 2;  t = 0 .. ceil(max_b s_b) - 1
//   w[b,i,t] = exp(-0.01 (t - c_i)^2) / (sum_i exp(-0.01 (t - c_i)^2) + 1e-20);  out[b,t,:] = sum_i w[b,i,t] x[b,i,:]
//
// The op is HBM-bound on its output (1 KB per frame at D = 256; + 4 L bytes per frame when `w` is materialised):
//   * gaussian_centres_kernel, one CTA per utterance: d[b,:] is staged in shared memory with coalesced loads, ONE thread
//     runs the cumulative sum the way torch.cumsum does on the CPU (sequential, double accumulator, every prefix rounded to
//     fp32: a parallel fp32 scan would move the centres of non-integer durations by an ulp of e ~ 1e-4, visible in w),
//     the centres go to a scratch buffer once per utterance
//     instead of once per CTA of the main kernel.
//   * gaussian_upsample_kernel, one CTA of 512 threads per (utterance, tile of 64 frames): exp(-0.01 D^2) is exactly 0 in
//     fp32 for |D| >= 104 (0.01 D^2 > 103.97, below the smallest denormal), so only the phonemes whose centre lies within
//     104 frames of the tile enter the denominators (binary search on the monotone centres; negative durations fall back
//     to the full range; 8 lanes per frame + warp shuffles).  The smallest denominator of the tile then bounds how far a
//     phoneme can sit and still reach a NORMALISED weight of 1e-12 anywhere in the tile (~53 frames when the tile lies
//     inside the utterance, the full 104 where every centre is far and the tiny denominators blow the weights up): only
//     those phonemes (~25, one chunk of 32) have their x rows fetched (cp.async, 16 bytes per lane) and their normalised
//     weights staged in shared memory; a warp additionally skips the phonemes whose weight stays below 1e-12 over its own 8
//     frames.  What is skipped changes a sum by < 1e-12 |x| per phoneme, four orders of magnitude below an fp32 ulp of the
//     result.  Each thread accumulates an 8-frame x 4-channel register tile (3 shared-memory vector loads per 32 FMAs); the
//     weight tensor `w`, when requested, is written exactly (every weight of the fp32 support, zeros elsewhere),
//     coalesced along t.  Output rows leave as 512-byte warp stores (+ operand planes inside the forward).
#include "fs2_common.cuh"
#include <math.h>

#define LAUNCHED_ERR() (++g_fs2_launches, cudaGetLastError())

namespace {

constexpr int GU_TF = 64;        // frames per CTA tile
constexpr int GU_FG = 8;         // frames per thread (register tile rows)
constexpr int GU_NG = GU_TF / GU_FG;   // frame groups per tile
constexpr int GU_CH = 32;        // phonemes per shared-memory chunk
constexpr int GU_THREADS = 64 * GU_NG;  // 64 channel quads x 8 frame groups = 512
constexpr float GU_CUT = 104.f;  // exp(-0.01 * 104^2) == 0 in fp32
constexpr float GU_SKIP = 1e-12f;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// centres[b, :], s[b], mono[b] (1: centres non-decreasing, the banded search is valid)
__global__ void __launch_bounds__(256) gaussian_centres_kernel(const float* d, int L, float* centres, float* s_out,
                                                               int* mono_out) {
  FS2_PDL_PROLOGUE();
  extern __shared__ float d_s[];  // [L]: durations in, centres out
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) d_s[i] = ld_act(d + (size_t)b * L + i);
  __syncthreads();
  if (threadIdx.x == 0) {
    // torch.cumsum on the CPU accumulates fp32 inputs in double (at::acc_type<float, false>) and rounds every prefix to fp32
    double e = 0.0;
    float prev = -INFINITY;
    int mono = 1;
    for (int i = 0; i < L; ++i) {
      const float di = d_s[i];
      e += (double)di;
      const float c = (float)e - 0.5f * di;
      d_s[i] = c;
      if (!(c >= prev)) mono = 0;
      prev = c;
    }
    if (s_out) s_out[b] = (float)e;
    mono_out[b] = mono;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) centres[(size_t)b * L + i] = d_s[i];
}

// One argument block for both uses of the kernel:
//   stand-alone operator: dense x [B, L, ldx] and out [B, T, ldo], centres from gaussian_centres_kernel;
//   inside the forward (fs2_set_upsampler): x = encoder output in the stage-1 ragged layout (utterance b at row
//   src_off[b], rows i >= src_rows[b] are the masked -- zero -- padded phonemes, which still count in the denominators,
//   modules.py has no masking), centres from the integer duration scan `cum` (exact in fp32), out (+ operand planes) in
//   the stage-2 ragged layout: grid rows t < dst_ext[b] carry values, the halo rows up to dst_off[b+1] are zero.
struct GuArgs {
  const float* x; int ldx;
  const int* src_off;      // null: utterance b starts at row b * L
  const int* src_rows;     // null: all L rows exist
  const float* centres;    // [B, L] or null
  const int* cum;          // [B, L] inclusive integer duration scan (used when centres == null)
  const int* mono;         // [B] or null (= monotone)
  int L, Dp, T, T_w;
  const int* L_dev;        // graph replays: the true L / T_w live in device memory (L, T_w above are upper bounds)
  const int* Tw_dev;
  float* out; int ldo;
  const int* dst_off;      // null: dense rows b * T + t
  const int* dst_ext;      // grid rows per utterance (ragged destination)
  bf16* out_b; int out_planes; size_t plane_elems;   // operand planes of the next consumer (forward only)
  float* w_out;
};

__global__ void __launch_bounds__(GU_THREADS, 2) gaussian_upsample_kernel(const GuArgs a) {   // <= 64 registers: 2 CTAs = 32 warps per SM
  FS2_PDL_PROLOGUE();
  extern __shared__ __align__(16) float gu_smem[];
  float* x_s = gu_smem;                        // [GU_CH][256]
  float* w_s = x_s + GU_CH * 256;              // [GU_CH][GU_TF]
  float* den_s = w_s + GU_CH * GU_TF;          // [GU_TF]
  float* c_s = den_s + GU_TF;                  // [L]
  __shared__ unsigned act_s[GU_NG];
  __shared__ float dmin_s[GU_THREADS / 32];

  const int b = blockIdx.y, t0 = blockIdx.x * GU_TF, L = shape_or(a.L_dev, a.L), T_w = shape_or(a.Tw_dev, a.T_w);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cq = tid & 63, fg = tid >> 6;      // channel quad, frame group (a warp lies inside one frame group)
  const int nq = a.Dp >> 2;
  const bool has_ch = cq < nq;
  size_t row0;                                 // destination row of frame 0
  int rows_b, ext_b;                           // rows to write / rows that carry weights
  if (a.dst_off) {
    const int o0 = ld_act(a.dst_off + b);
    row0 = (size_t)o0;
    rows_b = ld_act(a.dst_off + b + 1) - o0;
    ext_b = min(ld_act(a.dst_ext + b), T_w);
  } else {
    row0 = (size_t)b * a.T; rows_b = a.T; ext_b = T_w;
  }
  const int n_t = min(GU_TF, rows_b - t0);     // rows of the tile that exist in `out`
  if (n_t <= 0) return;
  const int n_w = max(0, min(GU_TF, ext_b - t0));   // frames with weights; the rest (`pad` rows / halo rows) are zero
  float* out_t = a.out + (row0 + t0) * a.ldo + cq * 4;
  bf16* outb_t = a.out_b ? a.out_b + (row0 + t0) * a.ldo + cq * 4 : nullptr;

  if (n_w == 0) {                              // pure padding tile
    if (has_ch)
      for (int f = fg; f < n_t; f += GU_NG) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(out_t + (size_t)f * a.ldo) = z;
        if (outb_t && a.out_planes > 0) store_planes4(outb_t + (size_t)f * a.ldo, a.plane_elems, a.out_planes, z);
      }
    return;
  }
  if (a.centres) {
    for (int i = tid; i < L; i += GU_THREADS) c_s[i] = ld_act(a.centres + (size_t)b * L + i);
  } else {       // c_i = e_i - d_i / 2 from the integer scan: bit-identical to the sequential fp32 sum below 2^24 frames
    for (int i = tid; i < L; i += GU_THREADS) {
      const int e = ld_act(a.cum + (size_t)b * L + i), ep = i > 0 ? ld_act(a.cum + (size_t)b * L + i - 1) : 0;
      c_s[i] = (float)e - 0.5f * (float)(e - ep);
    }
  }
  const int is_mono = a.mono ? ld_act(a.mono + b) : 1;
  const int n_src = a.src_rows ? min(ld_act(a.src_rows + b), L) : L;   // x rows beyond are exact zeros
  const size_t src0 = a.src_off ? (size_t)ld_act(a.src_off + b) : (size_t)b * L;
  __syncthreads();

  // phonemes that can reach the tile: c_i in (t0 - 104, t_last + 104)
  int i_lo = 0, i_hi = L;
  if (is_mono) {
    const float lo_v = (float)t0 - GU_CUT, hi_v = (float)(t0 + n_w - 1) + GU_CUT;
    int lo = 0, hi = L;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c_s[mid] > lo_v) hi = mid; else lo = mid + 1; }
    i_lo = lo;
    hi = L;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c_s[mid] >= hi_v) hi = mid; else lo = mid + 1; }
    i_hi = lo;
  }

  // denominators: 8 lanes per frame, lane q takes phonemes i_lo + q, + 8, ...
  float den_mine = INFINITY;
  {
    const int f = tid >> 3, q = tid & 7;
    const float tf = (float)(t0 + f);
    float part = 0.f;
    if (f < n_w)
      for (int i = i_lo + q; i < i_hi; i += 8) { const float dl = tf - c_s[i]; part += expf(-0.01f * (dl * dl)); }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (f < n_w) den_mine = part + 1e-20f;
    if (q == 0) den_s[f] = part + 1e-20f;
  }
  // smallest denominator of the tile -> how far a phoneme can sit and still reach a normalised weight of GU_SKIP somewhere
  // in the tile: exp(-0.01 D^2) / den_min >= GU_SKIP  <=>  D <= sqrt(-100 ln(GU_SKIP den_min)).  ~53 frames at den ~ 1; the
  // full 104 when a frame of the tile is far from every centre (tiny denominators make far phonemes matter).
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den_mine = fminf(den_mine, __shfl_xor_sync(0xffffffffu, den_mine, o));
  if (lane == 0) dmin_s[warp] = den_mine;
  __syncthreads();
  float den_min = dmin_s[0];
#pragma unroll
  for (int k = 1; k < GU_THREADS / 32; ++k) den_min = fminf(den_min, dmin_s[k]);
  int a_lo = i_lo, a_hi = i_hi;     // phonemes whose x rows are fetched and accumulated
  if (is_mono) {
    const float arg = GU_SKIP * den_min;                       // > 0: den_min >= 1e-20
    const float d_cut = fminf(GU_CUT, sqrtf(-100.f * logf(arg)) + 1.f);   // + 1 frame of slack on the analytic bound
    const float lo_v = (float)t0 - d_cut, hi_v = (float)(t0 + n_w - 1) + d_cut;
    int lo = i_lo, hi = i_hi;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c_s[mid] > lo_v) hi = mid; else lo = mid + 1; }
    a_lo = lo;
    hi = i_hi;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c_s[mid] >= hi_v) hi = mid; else lo = mid + 1; }
    a_hi = lo;
  }
  // w rows of the phonemes that are not accumulated: exact weights inside the fp32 support, zeros outside it
  if (a.w_out) {
    const int f = tid & 63;
    if (f < n_w) {
      const float tf = (float)(t0 + f), den = den_s[f];
      for (int i = tid >> 6; i < L; i += GU_NG) {
        if (i >= a_lo && i < a_hi) continue;
        float wv = 0.f;
        if (i >= i_lo && i < i_hi) { const float dl = tf - c_s[i]; wv = expf(-0.01f * (dl * dl)) / den; }
        a.w_out[((size_t)b * L + i) * T_w + t0 + f] = wv;
      }
    }
  }

  float4 acc[GU_FG];
#pragma unroll
  for (int j = 0; j < GU_FG; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int i0 = a_lo; i0 < a_hi; i0 += GU_CH) {
    const int n = min(GU_CH, a_hi - i0);
    const int n_x = max(0, min(n, n_src - i0));   // rows of the chunk that exist in x
    // x rows of the chunk -> shared memory (16 bytes per lane, coalesced 1 KB rows)
    for (int v = tid; v < n_x * nq; v += GU_THREADS) {
      const int ii = v / nq, c4 = v - ii * nq;
      cp_async16(x_s + ii * 256 + c4 * 4, a.x + (src0 + i0 + ii) * a.ldx + c4 * 4);
    }
    cp_async_commit();
    // normalised weights of the chunk, once per (frame, phoneme)
    {
      const int f = tid & 63;
      const float tf = (float)(t0 + f);
      const float den = den_s[f];
      for (int ii = tid >> 6; ii < n; ii += GU_NG) {
        const float dl = tf - c_s[i0 + ii];
        const float wv = f < n_w ? expf(-0.01f * (dl * dl)) / den : 0.f;   // pad rows (t >= T_w) stay zero
        w_s[ii * GU_TF + f] = wv;
        if (a.w_out && f < n_w) a.w_out[((size_t)b * L + i0 + ii) * T_w + t0 + f] = wv;
      }
    }
    cp_async_wait_all();
    __syncthreads();
    // which phonemes of the chunk matter to which frame group (one warp per group, one lane per phoneme)
    if (warp < GU_NG) {
      float m = 0.f;
      if (lane < n_x) {
        const float4* wr = reinterpret_cast<const float4*>(w_s + lane * GU_TF + warp * GU_FG);
#pragma unroll
        for (int j = 0; j < GU_FG / 4; ++j) { const float4 v = wr[j]; m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w))); }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, m >= GU_SKIP);
      if (lane == 0) act_s[warp] = mask;
    }
    __syncthreads();
    unsigned mask = act_s[fg];
    if (has_ch) {
      while (mask) {
        const int ii = __ffs(mask) - 1;
        mask &= mask - 1;
        const float4 xv = *reinterpret_cast<const float4*>(x_s + ii * 256 + cq * 4);
        const float4* wr = reinterpret_cast<const float4*>(w_s + ii * GU_TF + fg * GU_FG);
#pragma unroll
        for (int j4 = 0; j4 < GU_FG / 4; ++j4) {
          const float4 wv = wr[j4];
          const float ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float4& c = acc[j4 * 4 + k];
            c.x = fmaf(ws[k], xv.x, c.x); c.y = fmaf(ws[k], xv.y, c.y);
            c.z = fmaf(ws[k], xv.z, c.z); c.w = fmaf(ws[k], xv.w, c.w);
          }
        }
      }
    }
    __syncthreads();   // x_s / w_s are rewritten by the next chunk
  }
  if (has_ch) {
#pragma unroll
    for (int j = 0; j < GU_FG; ++j) {
      const int f = fg * GU_FG + j;
      if (f < n_t) {
        *reinterpret_cast<float4*>(out_t + (size_t)f * a.ldo) = acc[j];
        if (outb_t && a.out_planes > 0) store_planes4(outb_t + (size_t)f * a.ldo, a.plane_elems, a.out_planes, acc[j]);
      }
    }
  }
}

constexpr size_t GU_SMEM_MAX = 200 * 1024;
inline size_t gu_smem_bytes(int L) { return sizeof(float) * ((size_t)GU_CH * 256 + GU_CH * GU_TF + GU_TF + (size_t)L); }
inline cudaError_t gu_configure() {   // opt in to > 48 KB of dynamic shared memory (per device, idempotent)
  cudaError_t e = cudaFuncSetAttribute(gaussian_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GU_SMEM_MAX);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gaussian_centres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GU_SMEM_MAX);
  return e;
}

}  // namespace

cudaError_t rowops_gaussian_upsample(const float* x, const float* d, int B, int L, int D, int T, int T_w, float* out,
                                     float* s, float* w, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  if (L <= 0 || D <= 0 || D % 4 || B > 65535) return cudaErrorInvalidValue;
  const size_t smem_c = sizeof(float) * (size_t)L;
  if (gu_smem_bytes(L) > GU_SMEM_MAX) return cudaErrorInvalidValue;   // L <= ~40k phonemes per utterance
  cudaError_t e = gu_configure();
  if (e != cudaSuccess) return e;
  // scratch (centres + monotonicity flags) from the stream-ordered allocator: safe with concurrent calls on other streams
  float* centres = nullptr;
  const size_t bytes = sizeof(float) * (size_t)B * L + sizeof(int) * (size_t)B;
  e = cudaMallocAsync(reinterpret_cast<void**>(&centres), bytes, st);
  if (e != cudaSuccess) return e;
  int* mono = reinterpret_cast<int*>(centres + (size_t)B * L);
  g_fs2_plain_next = 1;   // follows a stream operation that is not one of this library's kernels
  (void)FS2_LAUNCH(gaussian_centres_kernel, dim3(B), 256, smem_c, st, d, L, centres, s, mono);
  e = LAUNCHED_ERR();
  if (e == cudaSuccess && T > 0) {
    dim3 grid((T + GU_TF - 1) / GU_TF, B);
    for (int c0 = 0; c0 < D && e == cudaSuccess; c0 += 256) {   // channel slabs of 256 (one slab at the model's D = 256)
      GuArgs a;
      memset(&a, 0, sizeof a);
      a.x = x + c0; a.ldx = D; a.centres = centres; a.mono = mono; a.L = L; a.Dp = D - c0 < 256 ? D - c0 : 256;
      a.T = T; a.T_w = T_w; a.out = out + c0; a.ldo = D; a.w_out = c0 == 0 ? w : nullptr;
      (void)FS2_LAUNCH(gaussian_upsample_kernel, grid, GU_THREADS, gu_smem_bytes(L), st, a);
      e = LAUNCHED_ERR();
    }
  }
  cudaError_t e2 = cudaFreeAsync(centres, st);
  g_fs2_plain_next = 1;
  return e != cudaSuccess ? e : e2;
}

// The forward's soft length regulator (fs2_set_upsampler(h, 1)): same arguments as rowops_length_regulate.  T_w = the
// frame count of the batch's longest utterance (`torch.arange(0, max(s))`, modules.py:174), i.e. the layout's S.
cudaError_t rowops_gaussian_regulate(const float* x, const int* src_off, const int* src_rows, const int* cum, int L, int D,
                                     const RowLayout& lay, float* out, bf16* out_b, int out_planes, cudaStream_t st,
                                     const int* L_dev) {
  if (lay.B <= 0 || lay.R_cap <= 0) return cudaSuccess;
  if (L <= 0 || D != 256 || gu_smem_bytes(L) > GU_SMEM_MAX) return cudaErrorInvalidValue;
  cudaError_t e = gu_configure();
  if (e != cudaSuccess) return e;
  GuArgs a;
  memset(&a, 0, sizeof a);
  a.x = x; a.ldx = D; a.src_off = src_off; a.src_rows = src_rows; a.cum = cum; a.L = L; a.Dp = D;
  a.T = lay.S; a.T_w = lay.S; a.out = out; a.ldo = D; a.dst_off = lay.off; a.dst_ext = lay.ext;
  a.out_b = out_b; a.out_planes = out_planes; a.plane_elems = (size_t)lay.R_cap * D;
  a.L_dev = L_dev; a.Tw_dev = lay.S_dev;
  dim3 grid((FS2_ROWS_PER_UTT(lay.S, FS2_HALO) + GU_TF - 1) / GU_TF, lay.B);   // covers ext + halo rows of the longest utterance
  (void)FS2_LAUNCH(gaussian_upsample_kernel, grid, GU_THREADS, gu_smem_bytes(L), st, a);
  return LAUNCHED_ERR();
}
