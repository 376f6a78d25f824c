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
//   * gaussian_upsample_kernel, one CTA of 4 warps per (utterance, tile of 32 frames), each warp owning 8 frames x all 256
//     channels (lane = 8 channels; 4 CTAs per SM).  exp(-0.01 D^2) is exactly 0 in fp32 for |D| >= 104 (0.01 D^2 > 103.97,
//     below the smallest denormal), so only the phonemes whose centre lies within 104 frames of the tile can matter (a
//     32-ary search of the monotone centres by warp ballots; negative durations fall back to the full range): the CTA
//     stages their x rows in shared memory once (cp.async, 16 bytes per lane).  Everything else is warp-local, so the
//     main loop has no CTA barrier:
//       - denominators of the warp's 8 frames, 4 lanes per frame + shuffles, over the phonemes within 64 frames (what is
//         farther adds < 1.6e-18 each); only a frame far from every centre (sum < 1e-6) makes the warp sum the full support;
//       - the smallest of them bounds how far a phoneme can sit and still reach a NORMALISED weight of 1e-12 on one of the
//         warp's frames (~53 frames; up to 104 where the denominators are tiny): ~17 phonemes survive at LJSpeech
//         durations.  What is skipped changes a sum by < 1e-12 |x| per phoneme, four orders below an fp32 ulp of the result;
//       - their normalised weights, evaluated once per (frame, phoneme), go to the warp's slice of shared memory as
//         (w, w) pairs, so that the accumulation is 2 + 4 shared-memory vector loads and 32 packed FMAs (fma.rn.f32x2: a
//         3-register FFMA issues every other cycle on this part, FFMA2 retires two per issue) per phoneme and lane.
//     Output rows leave as 512-byte warp stores (+ operand planes inside the forward).  The weight tensor `w`, when
//     requested, is written by a separate pass of the CTA, coalesced along t: exact weights inside the fp32 support,
//     zeros elsewhere.
//     Measured at the BASELINE configs[4] shape (batch 64 x 300 phonemes, T = 2160; profiles/r2_gaussian.md): 62 us per
//     call without `w` = 2.6 TB/s of algorithmic bytes (40 % of the measured HBM peak), 96 us with `w` (3.4 TB/s, 52 %).
//     The limiter is FP32 issue, not HBM: 17 phonemes x 256 channels per frame is 0.6 GFMA per call, the FMA pipe is 44 %
//     busy and half of the issued instructions are the per-warp bookkeeping around it.
#include "fs2_common.cuh"
#include <math.h>

#define LAUNCHED_ERR() (++g_fs2_launches, cudaGetLastError())

namespace {

#ifndef FS2_GU_FW
#define FS2_GU_FW 8
#endif
constexpr int GU_FW = FS2_GU_FW; // frames per warp (register tile: GU_FW frames x 8 channels per lane); 4 or 8
constexpr int GU_WARPS = 4;
constexpr int GU_TF = GU_WARPS * GU_FW;   // frames per CTA tile
constexpr int GU_QN = 32 / GU_FW;         // lanes per frame while denominators / weights are evaluated
static_assert(GU_FW == 4 || GU_FW == 8, "frames per warp");
constexpr int GU_THREADS = 32 * GU_WARPS;   // 256
constexpr int GU_XR = 40;        // phoneme rows of x staged per chunk (40 KB at D = 256)
constexpr float GU_CUT = 104.f;  // exp(-0.01 * 104^2) == 0 in fp32
constexpr float GU_NEAR = 64.f;  // exp(-0.01 * 64^2) = 1.6e-18: invisible next to a denominator >= GU_DEN_OK
constexpr float GU_DEN_OK = 1e-6f;
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

__global__ void __launch_bounds__(GU_THREADS, 4) gaussian_upsample_kernel(const GuArgs a) {
  FS2_PDL_PROLOGUE();
  extern __shared__ __align__(16) float gu_smem[];
  float* x_s = gu_smem;                        // [GU_XR][256]   phoneme rows of the current chunk
  float* w_s = x_s + GU_XR * 256;              // [GU_WARPS][GU_XR][2 * GU_FW]  per warp: weights as (w, w) pairs per frame
  float* inv_s = w_s + GU_WARPS * GU_XR * 2 * GU_FW;   // [GU_TF]  1 / denominator (only read by the `w` pass)
  float* c_s = inv_s + GU_TF;                  // [L]

  const int b = blockIdx.y, t0 = blockIdx.x * GU_TF, L = shape_or(a.L_dev, a.L), T_w = shape_or(a.Tw_dev, a.T_w);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = a.Dp >> 2;                    // channel quads of the slab; lane owns quads `lane` and `32 + lane`
  const bool has0 = lane < nq, has1 = 32 + lane < nq;
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
  const int f0 = warp * GU_FW;                 // first frame of this warp inside the tile
  float* out_w = a.out + (row0 + t0 + f0) * a.ldo + lane * 4;
  bf16* outb_w = a.out_b ? a.out_b + (row0 + t0 + f0) * a.ldo + lane * 4 : nullptr;
  auto store_row = [&](int j, float4 v0, float4 v1) {     // frame f0 + j: two 512-byte warp stores (+ operand planes)
    if (f0 + j >= n_t) return;
    float* o = out_w + (size_t)j * a.ldo;
    if (has0) *reinterpret_cast<float4*>(o) = v0;
    if (has1) *reinterpret_cast<float4*>(o + 128) = v1;
    if (outb_w && a.out_planes > 0) {
      bf16* ob = outb_w + (size_t)j * a.ldo;
      if (has0) store_planes4(ob, a.plane_elems, a.out_planes, v0);
      if (has1) store_planes4(ob + 128, a.plane_elems, a.out_planes, v1);
    }
  };
  if (n_w == 0) {                              // pure padding tile
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < GU_FW; ++j) store_row(j, z, z);
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

  // Bands on the monotone centres.  first(v, strict) = number of phonemes with c <= v (strict) or c < v: the index of the
  // first phoneme beyond v.  For the whole tile it is counted by the CTA (one predicate per thread and 256 phonemes,
  // __syncthreads_count); inside a warp, over the <= 64 phonemes of a window, by two ballots -- no serial search anywhere.
  // count(v, strict) over the whole utterance: a 32-ary search, the lanes probe 32 evenly spaced centres per level (two
  // levels up to 1024 phonemes).  Every warp of the CTA computes the same band: no exchange, no barrier.
  auto first_beyond = [&](float v, bool incl) {            // incl: count c <= v, else c < v
    int lo = 0, n = L;
    while (n > 32) {
      const int stride = (n + 31) >> 5, idx = lo + lane * stride;
      const float c = idx < lo + n ? c_s[idx] : INFINITY;
      const int k = __popc(__ballot_sync(0xffffffffu, incl ? c <= v : c < v));
      if (k == 0) return lo;
      const int nlo = lo + (k - 1) * stride + 1;           // probe k - 1 is before v, probe k (if any) is not
      n = min(lo + n, lo + k * stride) - nlo;
      lo = nlo;
    }
    const float c = lane < n ? c_s[lo + lane] : INFINITY;
    return lo + __popc(__ballot_sync(0xffffffffu, incl ? c <= v : c < v));
  };
  int i_lo = 0, i_hi = L;                      // phonemes inside the fp32 support of the tile: (t0 - 104, t_last + 104)
  if (is_mono) {
    i_lo = first_beyond((float)t0 - GU_CUT, true);
    i_hi = max(i_lo, first_beyond((float)(t0 + n_w - 1) + GU_CUT, false));
  }
  // [b_lo, b_hi) = phonemes of [lo, hi) with c in (v_lo - reach, v_hi + reach); all lanes get the same answer
  auto warp_band = [&](int lo, int hi, float v_lo, float v_hi, float reach, int& b_lo, int& b_hi) {
    b_lo = lo; b_hi = hi;
    if (!is_mono) return;
    const float a_ = v_lo - reach, z_ = v_hi + reach;
    if (hi - lo <= 64) {
      const float c0 = lo + lane < hi ? c_s[lo + lane] : INFINITY, c1 = lo + 32 + lane < hi ? c_s[lo + 32 + lane] : INFINITY;
      b_lo = lo + __popc(__ballot_sync(0xffffffffu, c0 <= a_)) + __popc(__ballot_sync(0xffffffffu, c1 <= a_));
      b_hi = lo + __popc(__ballot_sync(0xffffffffu, c0 < z_)) + __popc(__ballot_sync(0xffffffffu, c1 < z_));
    } else {                                   // many (zero-duration) phonemes in the window: binary searches
      int l0 = lo, h0 = hi;
      while (l0 < h0) { const int mid = (l0 + h0) >> 1; if (c_s[mid] > a_) h0 = mid; else l0 = mid + 1; }
      b_lo = l0; h0 = hi;
      while (l0 < h0) { const int mid = (l0 + h0) >> 1; if (c_s[mid] >= z_) h0 = mid; else l0 = mid + 1; }
      b_hi = l0;
    }
    b_hi = max(b_hi, b_lo);
  };
  auto stage_x = [&](int c0) {                 // rows [c0, c0 + GU_XR) of the band -> x_s (cp.async, 16 bytes per lane)
    const int n_x = max(0, min(min(GU_XR, i_hi - c0), n_src - c0));
    for (int r = warp; r < n_x; r += GU_WARPS) {
      const float* src = a.x + (src0 + c0 + r) * a.ldx + lane * 4;
      if (has0) cp_async16(x_s + r * 256 + lane * 4, src);
      if (has1) cp_async16(x_s + r * 256 + 128 + lane * 4, src + 128);
    }
    cp_async_commit();
  };
  stage_x(i_lo);

  // ---- this warp's frames: denominators.  Lane = (phoneme slot q, frame fr): 8 lanes per frame.  First over the
  // phonemes within GU_NEAR frames (what is farther adds < 1.6e-18 each); a frame whose sum stays below GU_DEN_OK is far
  // from every centre, the warp then sums the whole fp32 support (tiny denominators make far phonemes matter).
  const int fr = lane & (GU_FW - 1), q = lane / GU_FW;
  const int fw = f0 + fr;                      // this lane's frame inside the tile
  const bool f_ok = fw < n_w;
  const float tf = (float)(t0 + fw);
  const float tw_lo = (float)(t0 + f0), tw_hi = (float)(t0 + min(f0 + GU_FW, max(n_w, f0 + 1)) - 1);   // frames of the warp with weights
  int d_lo, d_hi;
  warp_band(i_lo, i_hi, tw_lo, tw_hi, GU_NEAR, d_lo, d_hi);
  auto den_over = [&](int lo, int hi) {
    float part = 0.f;
    if (f_ok)
      for (int i = lo + q; i < hi; i += GU_QN) { const float dl = tf - c_s[i]; part += expf(-0.01f * (dl * dl)); }
#pragma unroll
    for (int o = GU_FW; o < 32; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    return part;
  };
  float den = den_over(d_lo, d_hi);
  if (__any_sync(0xffffffffu, f_ok && den < GU_DEN_OK)) den = den_over(i_lo, i_hi);
  den += 1e-20f;
  const float inv = f_ok ? 1.0f / den : 0.f;   // pad rows (t >= T_w) keep zero weights
  // smallest denominator of the warp's frames -> how far a phoneme can sit and still reach a normalised weight of GU_SKIP
  // on one of them: exp(-0.01 D^2) / den_min >= GU_SKIP  <=>  D <= sqrt(-100 ln(GU_SKIP den_min))  (~53 frames at den ~ 1)
  float den_min = f_ok ? den : INFINITY;
#pragma unroll
  for (int o = 1; o < GU_FW; o <<= 1) den_min = fminf(den_min, __shfl_xor_sync(0xffffffffu, den_min, o));
  int a_lo = i_lo, a_hi = i_lo;                // phonemes this warp accumulates (none when it has no frame with weights)
  if (f0 < n_w) {
    const float d_cut = fminf(GU_CUT, sqrtf(-100.f * logf(GU_SKIP * den_min)) + 1.f);   // + 1 frame of slack on the bound
    warp_band(i_lo, i_hi, tw_lo, tw_hi, d_cut, a_lo, a_hi);
  }
  if (a.w_out && q == 0) inv_s[fw] = inv;

  // register tile: 4 frames x 8 channels as channel PAIRS, so that one FFMA2 (fma.rn.f32x2) updates two channels of a
  // frame: acc[j][p] = channels (pair p) of frame f0 + j.  x arrives from shared memory as natural pairs, the weights are
  // stored as (w, w) pairs by the lanes that evaluate them: no register shuffling around the 16 FFMA2 of a phoneme.
  float2 acc[GU_FW][4];
#pragma unroll
  for (int j = 0; j < GU_FW; ++j)
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) acc[j][p2] = make_float2(0.f, 0.f);
  float* w_w = w_s + warp * (GU_XR * 2 * GU_FW);

  for (int c0 = i_lo; c0 < i_hi; c0 += GU_XR) {
    if (c0 != i_lo) {
      __syncthreads();                          // every warp is done with the previous chunk's rows
      stage_x(c0);
    }
    const int lo = max(a_lo, c0), hi = min(min(a_hi, c0 + GU_XR), n_src);   // rows beyond n_src are zeros: nothing to add
    // normalised weights of this warp's frames, once per (frame, phoneme)
    for (int i = lo + q; i < hi; i += GU_QN) {
      const float dl = tf - c_s[i];
      const float wv = expf(-0.01f * (dl * dl)) * inv;
      *reinterpret_cast<float2*>(w_w + (i - c0) * (2 * GU_FW) + 2 * fr) = make_float2(wv, wv);
    }
    cp_async_wait_all();
    __syncthreads();                            // x rows (all warps' copies) and, warp-locally, the weights are visible
    const float* wr = w_w + (lo - c0) * (2 * GU_FW);
    const float* xr = x_s + (lo - c0) * 256 + lane * 4;
#pragma unroll 2
    for (int i = lo; i < hi; ++i, wr += 2 * GU_FW, xr += 256) {
      const float4 xa = *reinterpret_cast<const float4*>(xr);
      const float4 xb = *reinterpret_cast<const float4*>(xr + 128);
      const float2 xp[4] = {make_float2(xa.x, xa.y), make_float2(xa.z, xa.w), make_float2(xb.x, xb.y), make_float2(xb.z, xb.w)};
#pragma unroll
      for (int j2 = 0; j2 < GU_FW / 2; ++j2) {
        const float4 wq = *reinterpret_cast<const float4*>(wr + 4 * j2);     // (w, w) of frames 2 j2 and 2 j2 + 1
        const float2 wa = make_float2(wq.x, wq.y), wb = make_float2(wq.z, wq.w);
#pragma unroll
        for (int p2 = 0; p2 < 4; ++p2) {
          acc[2 * j2][p2] = __ffma2_rn(wa, xp[p2], acc[2 * j2][p2]);
          acc[2 * j2 + 1][p2] = __ffma2_rn(wb, xp[p2], acc[2 * j2 + 1][p2]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < GU_FW; ++j)
    store_row(j, make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y),
              make_float4(acc[j][2].x, acc[j][2].y, acc[j][3].x, acc[j][3].y));

  // ---- the weight tensor, when requested: w[b, i, t0 .. t0 + n_w) for every phoneme slot, coalesced along t (exact
  // weights inside the fp32 support, zeros outside it)
  if (a.w_out) {
    __syncthreads();                            // inv_s of every warp
    for (int f = lane; f < n_w; f += 32) {      // lane = frame: 128-byte rows of w
      const float tfw = (float)(t0 + f), invf = inv_s[f];
      for (int i = warp; i < L; i += GU_WARPS) {
        float wv = 0.f;
        if (i >= i_lo && i < i_hi) { const float dl = tfw - c_s[i]; wv = expf(-0.01f * (dl * dl)) * invf; }
        a.w_out[((size_t)b * L + i) * T_w + t0 + f] = wv;
      }
    }
  }
}

constexpr size_t GU_SMEM_MAX = 200 * 1024;
inline size_t gu_smem_bytes(int L) { return sizeof(float) * ((size_t)GU_XR * 256 + GU_WARPS * GU_XR * 2 * GU_FW + GU_TF + (size_t)L); }
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
