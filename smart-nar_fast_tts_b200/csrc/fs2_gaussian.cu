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
//     result.  Each thread accumulates an 8-frame x 4-channel register tile with packed fp32 FMAs (fma.rn.f32x2: a plain
//     3-register FFMA issues every other cycle on this part, FFMA2 retires two per issue; 3 shared-memory vector loads per
//     16 FFMA2); the
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

// centres[b, :], s[b], mono[b] (1: centres non-decreasing, the banded search is valid).  torch.cumsum on the CPU
// accumulates fp32 inputs in DOUBLE (at::acc_type<float, false>) and rounds every prefix to fp32; double sums of fp32
// values are exact while their exponents span < 29 bits (any realistic durations), so a block-parallel double scan gives
// the very same prefixes as the sequential loop: thread t scans its chunk of ceil(L / 256) durations, the chunk totals
// are combined by warp shuffles.
__global__ void __launch_bounds__(256) gaussian_centres_kernel(const float* d, int L, float* centres, float* s_out,
                                                               int* mono_out) {
  FS2_PDL_PROLOGUE();
  extern __shared__ float d_s[];  // [L]: durations in, centres out
  __shared__ double wtot[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < L; i += blockDim.x) d_s[i] = ld_act(d + (size_t)b * L + i);
  __syncthreads();
  const int per = (L + 255) / 256, i0 = tid * per, i1 = min(L, i0 + per);
  double tot = 0.0;
  for (int i = i0; i < i1; ++i) tot += (double)d_s[i];
  double incl = tot;                                   // inclusive scan of the chunk totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wtot[w] = incl;
  __syncthreads();
  double base = incl - tot;
  for (int k = 0; k < w; ++k) base += wtot[k];
  double e = base;
  for (int i = i0; i < i1; ++i) {
    const float di = d_s[i];
    e += (double)di;
    d_s[i] = (float)e - 0.5f * di;
  }
  if (tid == 255 && s_out) s_out[b] = (float)e;        // the last chunk ends at L (empty chunks carry the total along)
  __syncthreads();
  int ok = 1;
  for (int i = tid; i < L; i += blockDim.x) {
    const float c = d_s[i];
    if (!(c >= (i > 0 ? d_s[i - 1] : -INFINITY))) ok = 0;
    centres[(size_t)b * L + i] = c;
  }
  ok = __syncthreads_and(ok);
  if (tid == 0) mono_out[b] = ok;
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
  float* inv_s = w_s + GU_CH * GU_TF;          // [GU_TF]  1 / denominator of the frame
  float* c_s = inv_s + GU_TF;                  // [L]
  __shared__ unsigned act_s[GU_NG];
  __shared__ float dmin_s[GU_THREADS / 32];
  __shared__ int band_s[4];                    // i_lo, i_hi (fp32 support of the tile), a_lo, a_hi (accumulated phonemes)

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

  // first phoneme with c > v (strict) or c >= v, by binary search on the monotone centres; lanes 0 / 1 of warp 0 find the
  // two ends of a band at the same time and publish them (512 threads repeating the searches cost a third of the kernel)
  auto lower = [&](int lo, int hi, float v, bool strict) {
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const float c = c_s[mid];
      if (strict ? c > v : c >= v) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  // phonemes that can reach the tile at all: c_i in (t0 - 104, t_last + 104), the fp32 support of the Gaussian
  if (tid < 2) {
    int v = tid == 0 ? 0 : L;
    if (is_mono) v = tid == 0 ? lower(0, L, (float)t0 - GU_CUT, true) : lower(0, L, (float)(t0 + n_w - 1) + GU_CUT, false);
    band_s[tid] = v;
  }
  __syncthreads();
  const int i_lo = band_s[0], i_hi = band_s[1];

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
    const float den = part + 1e-20f;
    if (f < n_w) den_mine = den;
    if (q == 0) inv_s[f] = 1.0f / den;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den_mine = fminf(den_mine, __shfl_xor_sync(0xffffffffu, den_mine, o));
  if (lane == 0) dmin_s[warp] = den_mine;
  __syncthreads();
  // smallest denominator of the tile -> how far a phoneme can sit and still reach a normalised weight of GU_SKIP somewhere
  // in the tile: exp(-0.01 D^2) / den_min >= GU_SKIP  <=>  D <= sqrt(-100 ln(GU_SKIP den_min)).  ~53 frames at den ~ 1; the
  // full 104 when a frame of the tile is far from every centre (tiny denominators make far phonemes matter).
  if (tid < 2) {
    int v = tid == 0 ? i_lo : i_hi;
    if (is_mono) {
      float den_min = dmin_s[0];
#pragma unroll
      for (int k = 1; k < GU_THREADS / 32; ++k) den_min = fminf(den_min, dmin_s[k]);
      const float d_cut = fminf(GU_CUT, sqrtf(-100.f * logf(GU_SKIP * den_min)) + 1.f);   // + 1 frame of slack on the bound
      v = tid == 0 ? lower(i_lo, i_hi, (float)t0 - d_cut, true) : lower(i_lo, i_hi, (float)(t0 + n_w - 1) + d_cut, false);
    }
    band_s[2 + tid] = v;
  }
  __syncthreads();
  const int a_lo = band_s[2], a_hi = max(band_s[2], band_s[3]);     // phonemes whose x rows are fetched and accumulated

  // w rows of the phonemes that are not accumulated: exact weights inside the fp32 support, zeros outside it
  if (a.w_out) {
    const int f = tid & 63;
    if (f < n_w) {
      const float tf = (float)(t0 + f), inv = inv_s[f];
      for (int i = tid >> 6; i < L; i += GU_NG) {
        if (i >= a_lo && i < a_hi) continue;
        float wv = 0.f;
        if (i >= i_lo && i < i_hi) { const float dl = tf - c_s[i]; wv = expf(-0.01f * (dl * dl)) * inv; }
        a.w_out[((size_t)b * L + i) * T_w + t0 + f] = wv;
      }
    }
  }

  // register tile: 4 channels x 8 frames held as frame PAIRS so that one FFMA2 (fma.rn.f32x2) updates two frames of a channel:
  // acc[c][p] = (frame 2p, frame 2p + 1) of channel c.  The weights of a phoneme arrive as aligned pairs straight from
  // shared memory; only the 4 channel values are duplicated (4 moves per 16 FFMA2).
  float2 acc[4][GU_FG / 2];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int p2 = 0; p2 < GU_FG / 2; ++p2) acc[c][p2] = make_float2(0.f, 0.f);

  for (int i0 = a_lo; i0 < a_hi; i0 += GU_CH) {
    const int n = min(GU_CH, a_hi - i0);
    const int n_x = max(0, min(n, n_src - i0));   // rows of the chunk that exist in x
    // x rows of the chunk -> shared memory (16 bytes per lane, coalesced 1 KB rows)
    for (int ii = tid >> 6; ii < n_x; ii += GU_NG)
      if (has_ch) cp_async16(x_s + ii * 256 + cq * 4, a.x + (src0 + i0 + ii) * a.ldx + cq * 4);
    cp_async_commit();
    // normalised weights of the chunk, once per (frame, phoneme)
    {
      const int f = tid & 63;
      const float tf = (float)(t0 + f);
      const float inv = f < n_w ? inv_s[f] : 0.f;   // pad rows (t >= T_w) stay zero
      for (int ii = tid >> 6; ii < n; ii += GU_NG) {
        const float dl = tf - c_s[i0 + ii];
        const float wv = expf(-0.01f * (dl * dl)) * inv;
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
        const float2 xd[4] = {make_float2(xv.x, xv.x), make_float2(xv.y, xv.y), make_float2(xv.z, xv.z), make_float2(xv.w, xv.w)};
#pragma unroll
        for (int j4 = 0; j4 < GU_FG / 4; ++j4) {
          const float4 wv = wr[j4];
          const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[c][2 * j4] = __ffma2_rn(w01, xd[c], acc[c][2 * j4]);
            acc[c][2 * j4 + 1] = __ffma2_rn(w23, xd[c], acc[c][2 * j4 + 1]);
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
        const float4 v = (j & 1) ? make_float4(acc[0][j >> 1].y, acc[1][j >> 1].y, acc[2][j >> 1].y, acc[3][j >> 1].y)
                                 : make_float4(acc[0][j >> 1].x, acc[1][j >> 1].x, acc[2][j >> 1].x, acc[3][j >> 1].x);
        *reinterpret_cast<float4*>(out_t + (size_t)f * a.ldo) = v;
        if (outb_t && a.out_planes > 0) store_planes4(outb_t + (size_t)f * a.ldo, a.plane_elems, a.out_planes, v);
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
