// fs2_rowops.cu -- the HBM-bound row operators of the path (sm_100a): gathers, scans, the
// length regulator, the Gaussian upsampler, bucketize+embedding, masks and weight repacking.
// All of them move 16-byte vectors per lane with coalesced row accesses; none has reuse
// worth staging beyond a per-CTA copy of the small per-utterance tables (cumulative
// durations / centres), so there is no shared-memory tiling of the payload.
#include "fs2_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

std::atomic<long long> g_fs2_launches{0};
int g_fs2_pdl = getenv("FS2_NO_PDL") ? 0 : 1;
thread_local int g_fs2_plain_next = 1;
thread_local int g_fs2_pdl_off = 0;

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return ld_act(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }


// ---------------------------------------------------------------------------------------------
// Ragged-grid layout (fs2_common.cuh): ext[b] = min(lens[b] + halo_keep, S); off = exclusive scan of ext + halo_rows
// rounded up to FS2_ROW_ALIGN rows (utterances start on 16-byte boundaries of the transposed-V operand).
// One CTA, chunks of 1024 utterances, block-wide scan by warp shuffles.
__global__ void __launch_bounds__(1024) build_layout_kernel(const int* lens, int Breal, int S_host, int halo_keep,
                                                            int halo_rows, int* off, int* ext,
                                                            int extra_ext, const int* S_dev) {
  FS2_PDL_PROLOGUE();
  const int S = shape_or(S_dev, S_host);
  const int B = Breal + (extra_ext > 0 ? 1 : 0);   // the pseudo utterance (index Breal) has min(extra_ext, S) grid rows
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int b = base + tid;
    int e = 0, v = 0;
    if (b < B) {
      const long long want = b >= Breal ? (long long)extra_ext : lens ? (long long)lens[b] + halo_keep : (long long)S;
      e = (int)(want < S ? want : S);
      if (e < 0) e = 0;
      ext[b] = e;
      v = halo_rows > 0 ? ((e + halo_rows + FS2_ROW_ALIGN - 1) / FS2_ROW_ALIGN) * FS2_ROW_ALIGN : e;
    }
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int incl = carry_s + x + (w > 0 ? warp_tot[w - 1] : 0);
    if (b < B) off[b] = incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) off[B] = carry_s;
}
__global__ void fill_rowmap_kernel(const int* off, const int* ext, int B, int R_cap,
                                   unsigned* rowmap) {
  FS2_PDL_PROLOGUE();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R_cap) return;
  unsigned code = FS2_ROW_NONE;
  if (r < off[B]) {
    int lo = 0, hi = B - 1;  // last b with off[b] <= r
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (off[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const int p = r - off[lo];
    if (p < ext[lo]) code = ((unsigned)lo << 16) | (unsigned)p;
  }
  rowmap[r] = code;
}

// transformer/Models.py:82-91  out[b,p,:] = src_word_emb[texts[b,p]] + PE[p]   (all grid rows, incl. PAD ids)
// one warp per flat row; D/4 float4 per row; optional bf16 / bf16x3 shadow for the first GEMM
__global__ void embed_pe_kernel(const int64_t* texts, const float* emb,
                                const float* pe, int vocab, const RowLayout lay, int D,
                                float* out_grid, bf16* out_b, int out_planes,
                                float* out_user) {
  FS2_PDL_PROLOGUE();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int R = ld_act(lay.off + lay.B);
  if (r >= R) return;
  const RowPos rp = row_pos(lay, r, R);
  const int nv = D >> 2;
  const size_t plane = (size_t)lay.R_cap * D;
  if (!rp.in_grid) {  // halo rows stay zero
    for (int c = lane; c < nv; c += 32) {
      if (out_grid) st4(out_grid + (size_t)r * D + c * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      if (out_b && out_planes > 0) store_planes4(out_b + (size_t)r * D + c * 4, plane, out_planes, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    return;
  }
  const int S = lay_S(lay);
  long long id = texts[(size_t)rp.b * S + rp.p];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // the reference would raise; clamp instead of faulting
  for (int c = lane; c < nv; c += 32) {
    const float4 e = ld4(emb + (size_t)id * D + c * 4), q = ld4(pe + (size_t)rp.p * D + c * 4);
    const float4 v = make_float4(e.x + q.x, e.y + q.y, e.z + q.z, e.w + q.w);
    if (out_grid) st4(out_grid + (size_t)r * D + c * 4, v);
    if (out_b && out_planes > 0) store_planes4(out_b + (size_t)r * D + c * 4, plane, out_planes, v);
    if (out_user) st4(out_user + ((size_t)rp.b * S + rp.p) * D + c * 4, v);
  }
}

__global__ void lens_to_i32_kernel(const int64_t* lens, int B, int cap_host, int* out, int* zero2, const int* cap_dev) {
  FS2_PDL_PROLOGUE();
  const int cap = shape_or(cap_dev, cap_host);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && zero2) { zero2[0] = 0; zero2[1] = 0; }   // the {T_max, frames} accumulators of this forward
  if (i < B) {
    long long v = lens[i];
    out[i] = (int)(v < 0 ? 0 : (v > cap ? cap : v));
  }
}

// utils/tools.py:89-97  mask[b,i] = i >= lens[b]
// zero0 / zero1 (optional, same [B, max_len] shape, fp32): cleared in the same pass -- the variance predictors only write
// the rows their packed layout carries (modules.py:285 masked_fill(mask, 0) for the rest)
__global__ void mask_kernel(const int64_t* lens64, const int* lens32, int B, int max_len_host,
                            uint8_t* mask, float* zero0, float* zero1, const int* max_len_dev) {
  FS2_PDL_PROLOGUE();
  const int max_len = shape_or(max_len_dev, max_len_host);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * max_len) return;
  const int b = (int)(i / max_len), p = (int)(i - (size_t)b * max_len);
  const long long len = lens64 ? lens64[b] : (long long)lens32[b];
  mask[i] = (uint8_t)(p >= len);
  if (zero0) zero0[i] = 0.f;
  if (zero1) zero1[i] = 0.f;
}

// model/modules.py:132-135  clamp(round(exp(log_d) - 1) * d_control, min=0); torch.round = half-to-even = rintf
__device__ __forceinline__ float round_duration(float log_d, float d_control) {
  return fmaxf(rintf(expf(log_d) - 1.0f) * d_control, 0.0f);
}
__global__ void round_durations_kernel(const float* log_d, int64_t n, float d_control,
                                       float* out) {
  FS2_PDL_PROLOGUE();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = round_duration(log_d[i], d_control);
}

// model/modules.py:206-222 bookkeeping: expand_size = max(int(d), 0); mel_len = sum.  One CTA per utterance,
// block-wide inclusive scan (warp shuffles + one smem hop), chunks of 1024 phonemes.
// log_d != null: the kernel first applies modules.py:132-135 (rounding) to log_d and writes d (fused forward path);
// log_d == null: d is an input (stand-alone operator).
__global__ void __launch_bounds__(1024) duration_scan_kernel(const float* log_d, float d_control, float* d, int L_host, int* cum,
                                                             int64_t* mel_lens,
                                                             int* mel_lens32, int* tmax, const int* L_dev) {
  FS2_PDL_PROLOGUE();
  const int L = shape_or(L_dev, L_host);
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < L; base += 1024) {
    const int i = base + tid;
    int v = 0;
    if (i < L) {
      float f;
      if (log_d) {
        f = round_duration(log_d[(size_t)b * L + i], d_control);
        d[(size_t)b * L + i] = f;
      } else {
        f = d[(size_t)b * L + i];
      }
      // int() truncates toward zero; guard the conversion against inf/NaN/huge values
      v = (f > 0.f) ? (f < 1.0e6f ? (int)f : 1000000) : 0;
    }
    // the running sums saturate at SAT (far above the 65535 frames an utterance may have, far below INT_MAX): garbage
    // log-durations (inf, untrained weights) must end in a loud "too many frames" error of stage 2, never in a wrapped,
    // plausible-looking length.  Every partial sum is clamped, so no int32 addition below can overflow (2 * SAT < 2^31).
    constexpr int SAT = 1 << 30;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x = min(x + y, SAT);
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t = min(t + y, SAT);
      }
      warp_tot[lane] = t;  // inclusive totals
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = min(min(carry + x, SAT) + (w > 0 ? warp_tot[w - 1] : 0), SAT);
    if (i < L) cum[(size_t)b * L + i] = incl;
    __syncthreads();
    if (tid == 1023) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) {
    const int total = carry_s;
    if (mel_lens) mel_lens[b] = total;
    if (mel_lens32) mel_lens32[b] = total;
    if (tmax) {               // tmax[0] = longest utterance (frames), tmax[1] = frames of the whole batch (saturating)
      atomicMax(tmax, total);
      // batch total: one plain atomicAdd per utterance (a compare-and-swap loop serialises the B blocks: 90 us at batch
      // 256).  Each addend is clamped so that B of them cannot wrap int32; the clamp (>= 32767 frames per utterance at the
      // largest batch) is never reached by a forward that stage 2 accepts (65535 frames per utterance, R_cap < 2^31 rows).
      atomicAdd(tmax + 1, min(total, 0x7fffffff / (int)gridDim.x));
    }
  }
}

// model/modules.py:220-226 + utils/tools.py:288-306: frame t of utterance b copies phoneme row i with
// cum[i-1] <= t < cum[i]; frames >= mel_len (kept padded rows and halo) are zero.  One warp per output row, the
// utterance's cumulative table staged in shared memory, binary search per row, 16-byte row copy (+ bf16 shadow).
__global__ void __launch_bounds__(256) length_regulate_kernel(const float* x, const int* src_off,
                                                              int src_stride, const int* cum, int L_host, int D,
                                                              const RowLayout lay, float* out,
                                                              bf16* out_b, int out_planes, const int* L_dev) {
  FS2_PDL_PROLOGUE();
  const int L = shape_or(L_dev, L_host);
  extern __shared__ int cum_s[];
  const int b = blockIdx.y;
  const int rows_b = ld_act(lay.off + b + 1) - ld_act(lay.off + b);   // ext + halo
  const int rows_per_cta = (blockDim.x >> 5) * 8;
  if ((int)blockIdx.x * rows_per_cta >= rows_b) return;
  for (int i = threadIdx.x; i < L; i += blockDim.x) cum_s[i] = cum[(size_t)b * L + i];
  __syncthreads();
  const int total = L > 0 ? cum_s[L - 1] : 0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nv = D >> 2;
  const size_t plane = (size_t)lay.R_cap * D;
  const size_t row0 = (size_t)ld_act(lay.off + b);
  const size_t src0 = src_off ? (size_t)ld_act(src_off + b) : (size_t)b * src_stride;
  for (int k = 0; k < 8; ++k) {
    const int t = blockIdx.x * rows_per_cta + k * (blockDim.x >> 5) + w;
    if (t >= rows_b) continue;
    float* dst = out + (row0 + t) * D;
    bf16* dstb = out_b ? out_b + (row0 + t) * D : nullptr;
    const float* src = nullptr;
    if (t < total && t < lay_S(lay)) {
      int lo = 0, hi = L - 1;  // first i with cum[i] > t
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum_s[mid] > t) hi = mid; else lo = mid + 1;
      }
      src = x + (src0 + lo) * D;
    }
    for (int c = lane; c < nv; c += 32) {
      const float4 v = src ? ld4(src + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      st4(dst + c * 4, v);
      if (dstb && out_planes > 0) store_planes4(dstb + c * 4, plane, out_planes, v);
    }
  }
}

// model/modules.py:80-100 (inference branch) fused with the decoder's positional add (Models.py:231-233):
//   pred <- pred * control ; idx = bucketize(pred, bins) ; x[row] += emb[idx] (+ pe[p])
// torch.bucketize(right=False) lower bound, incl. its behaviour on NaN boundaries (every compare false -> n_bins-1).
__global__ void variance_embed_kernel(float* pred, float control, const float* bins,
                                      int n_bins, const float* emb, const float* pe,
                                      float* x, bf16* xb, int xb_planes, const RowLayout lay,
                                      int D, int* idx_out) {
  FS2_PDL_PROLOGUE();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int R = ld_act(lay.off + lay.B);
  if (r >= R) return;
  const RowPos rp = row_pos(lay, r, R);
  if (!rp.in_grid) return;   // halo rows stay zero
  const size_t u = (size_t)rp.b * lay_S(lay) + rp.p;
  const float v = pred[u] * control;
  int start = 0, end = n_bins - 1;  // boundaries array has n_bins-1 entries
  while (start < end) {
    const int mid = start + ((end - start) >> 1);
    const float mv = ld_act(bins + mid);
    if (!(mv >= v)) start = mid + 1; else end = mid;
  }
  if (lane == 0) {
    pred[u] = v;
    if (idx_out) idx_out[u] = start;
  }
  const size_t row = (size_t)r;
  const int nv = D >> 2;
  for (int c = lane; c < nv; c += 32) {
    float4 a = *reinterpret_cast<const float4*>(x + row * D + c * 4);
    const float4 e = ld4(emb + (size_t)start * D + c * 4);
    a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    if (pe) {
      const float4 q = ld4(pe + (size_t)rp.p * D + c * 4);
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    st4(x + row * D + c * 4, a);
    if (xb && xb_planes > 0) store_planes4(xb + row * D + c * 4, (size_t)lay.R_cap * D, xb_planes, a);
  }
}

// Rows of the destination (PostNet) layout that the packed source layout does not carry (see fs2_common.cuh): padded
// grid rows get the bias row (what mel_linear makes of a zero decoder row, fastspeech2_align.py:83), the rows between an
// utterance's grid rows and the next utterance get zero (the convolutions' zero padding).  One warp per row.
__global__ void __launch_bounds__(256) fill_padded_rows_kernel(const float* bias, int N, const RowLayout src,
                                                               const RowLayout dst, float* out_grid,
                                                               bf16* out_b, int out_planes,
                                                               float* out_user) {
  FS2_PDL_PROLOGUE();
  const int b = blockIdx.y;
  const int e_src = b < src.B ? ld_act(src.ext + b) : 0;          // the pseudo utterance has no source rows
  const int p = e_src + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int o0 = ld_act(dst.off + b);
  const int rows_dst = ld_act(dst.off + b + 1) - o0;               // grid rows + zero rows of utterance b
  const int e_dst = ld_act(dst.ext + b);
  const int S_user = lay_S(src);
  const bool in_dst = p < rows_dst, in_user = b < src.B && p < S_user;
  if (!in_dst && !in_user) return;
  const int lane = threadIdx.x & 31, nv = N >> 2;
  const size_t g = (size_t)o0 + p;
  const size_t plane = (size_t)dst.R_cap * N;
  for (int c = lane; c < nv; c += 32) {
    const float4 bv = ld4(bias + c * 4);
    if (in_dst) {
      const float4 v = p < e_dst ? bv : make_float4(0.f, 0.f, 0.f, 0.f);
      if (out_grid) st4(out_grid + g * N + c * 4, v);
      if (out_b && out_planes > 0) store_planes4(out_b + g * N + c * 4, plane, out_planes, v);
    }
    if (out_user && in_user) st4(out_user + ((size_t)b * S_user + p) * N + c * 4, bv);
  }
}

// PostNet far rows (fs2_common.cuh): utterance b's rows p >= ext[b] - H (when the layout cut the utterance short of S)
// equal the rows of an all-bias utterance at the same distance from the end of the grid.  One warp per row.
__global__ void __launch_bounds__(256) postnet_far_rows_kernel(const float* post_grid, int N, const RowLayout pn,
                                                               int B, int H, float* out_user) {
  FS2_PDL_PROLOGUE();
  const int b = blockIdx.y, S = lay_S(pn);
  const int e = ld_act(pn.ext + b);
  if (e >= S) return;                                   // the utterance reaches the end of the grid: every row is exact
  const int p = max(e - H, 0) + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= S) return;
  const int Tp = ld_act(pn.ext + B);                     // rows of the pseudo utterance: min(S, 2H+1)
  const int j = S > Tp ? ((S - p <= H) ? Tp - (S - p) : H) : p;
  const float* src = post_grid + ((size_t)ld_act(pn.off + B) + j) * N;
  float* dst = out_user + ((size_t)b * S + p) * N;
  for (int c = threadIdx.x & 31; c < (N >> 2); c += 32) st4(dst + c * 4, ld4(src + c * 4));
}

// Same rows for a channel-major user tensor [B, N, S]: one thread per row p (consecutive threads = consecutive addresses
// in every channel), looping over the channels; the source row is the same for almost every thread (broadcast loads).
__global__ void __launch_bounds__(256) postnet_far_rows_cm_kernel(const float* post_grid, int N, const RowLayout pn,
                                                                  int B, int H, float* out_user) {
  FS2_PDL_PROLOGUE();
  const int b = blockIdx.y, S = lay_S(pn);
  const int e = ld_act(pn.ext + b);
  if (e >= S) return;
  const int p = max(e - H, 0) + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= S) return;
  const int Tp = ld_act(pn.ext + B);
  const int j = S > Tp ? ((S - p <= H) ? Tp - (S - p) : H) : p;
  const float* src = post_grid + ((size_t)ld_act(pn.off + B) + j) * N;
  float* dst = out_user + (size_t)b * N * S + p;
  for (int c = 0; c < N; ++c) dst[(size_t)c * S] = ld_act(src + c);
}

// row p = 0 of every utterance <- 0 (MelEncoder replaces the first mel frame by zeros, Models.py:144-145); one warp per utterance
__global__ void zero_first_rows_kernel(const RowLayout lay, int C, float* x) {
  FS2_PDL_PROLOGUE();
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= lay.B || ld_act(lay.ext + b) <= 0) return;
  float* row = x + (size_t)ld_act(lay.off + b) * C;
  for (int c = lane; c < (C >> 2); c += 32) st4(row + c * 4, make_float4(0.f, 0.f, 0.f, 0.f));
}

// dense user layout [B,S,C] <-> ragged grid layout (test / per-operator entry points)
__global__ void to_grid_kernel(const float* xu, const RowLayout lay, int C, float* out,
                               int ldo, int col_off, bf16* out_b) {
  FS2_PDL_PROLOGUE();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 per thread
  const int nv = C >> 2;
  if (i >= (size_t)lay.R_cap * nv) return;
  const size_t row = i / nv;
  const int c = (int)(i - row * nv);
  const RowPos rp = row_pos(lay, (int)row, ld_act(lay.off + lay.B));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rp.in_grid) v = ld4(xu + ((size_t)rp.b * lay_S(lay) + rp.p) * C + c * 4);
  if (out) st4(out + row * ldo + col_off + c * 4, v);
  if (out_b) *reinterpret_cast<uint2*>(out_b + row * C + c * 4) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
}
__global__ void from_grid_kernel(const float* xg, const RowLayout lay, int C, float* out) {
  FS2_PDL_PROLOGUE();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = C >> 2;
  if (i >= (size_t)lay.B * lay.S * nv) return;
  const size_t row = i / nv;
  const int c = (int)(i - row * nv);
  const int b = (int)(row / lay.S), p = (int)(row - (size_t)b * lay.S);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p < ld_act(lay.ext + b)) v = ld4(xg + ((size_t)ld_act(lay.off + b) + p) * C + c * 4);
  st4(out + row * C + c * 4, v);
}

// torch Conv1d / Linear weight [N][K][taps] -> fp32 [taps][K][n_total] (columns n_off..) and/or
// bf16 [taps][n_total][K] (rows n_off..), optionally scaled per output channel (BatchNorm fold).
__global__ void pack_weight_kernel(const float* src, int N, int K, int taps, const float* scale,
                                   float* dst_f, bf16* dst_b, int n_total, int n_off) {
  FS2_PDL_PROLOGUE();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * K * taps) return;
  const int t = (int)(i % taps);
  const int k = (int)((i / taps) % K);
  const int n = (int)(i / ((size_t)taps * K));
  float v = src[i];
  if (scale) v *= scale[n];
  if (dst_f) dst_f[((size_t)t * K + k) * n_total + n_off + n] = v;
  if (dst_b) {
    const size_t plane = (size_t)taps * n_total * K, idx = ((size_t)t * n_total + n_off + n) * K + k;
    const bf16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const bf16 mid = __float2bfloat16_rn(r1);
    dst_b[idx] = hi;
    dst_b[plane + idx] = mid;
    dst_b[2 * plane + idx] = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  }
}

// BatchNorm1d(eval) folded into the preceding conv (transformer/Layers.py:120-167, eps 1e-5):
//   scale = g / sqrt(var + eps) ; bias' = (conv_bias - mean) * scale + b
__global__ void bn_fold_kernel(const float* conv_bias, const float* g,
                               const float* b, const float* mean,
                               const float* var, int n, float eps, float* scale_out,
                               float* bias_out) {
  FS2_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = g[i] / sqrtf(var[i] + eps);
  scale_out[i] = s;
  bias_out[i] = (conv_bias[i] - mean[i]) * s + b[i];
}

__global__ void split_kernel(const float* src, int64_t n4, int planes, bf16* dst, int64_t plane_elems) {
  FS2_PDL_PROLOGUE();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) store_planes4(dst + i * 4, (size_t)plane_elems, planes, ld4(src + i * 4));
}

// max |w| of a tensor as the bit pattern of a non-negative float (monotonic as unsigned)
__global__ void absmax_kernel(const float* w, size_t n, unsigned* out) {
  FS2_PDL_PROLOGUE();
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fabsf(w[i]);
    if (v == v && v <= 3.0e38f) m = fmaxf(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// wf [taps][K][N] fp32 -> dst [2][taps][N][K] fp16 bit patterns of (w * scale): hi, lo
__global__ void pack_weight_f16x2_kernel(const float* wf, int N, int K, int taps, float scale, bf16* dst) {
  FS2_PDL_PROLOGUE();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into dst plane: ((t*N + n)*K + k)
  const size_t plane = (size_t)taps * N * K;
  if (i >= plane) return;
  const int k = (int)(i % K);
  const int n = (int)((i / K) % N);
  const int t = (int)(i / ((size_t)K * N));
  uint16_t hi, lo;
  split2h_scaled(wf[((size_t)t * K + k) * N + n] * scale, hi, lo);
  reinterpret_cast<uint16_t*>(dst)[i] = hi;
  reinterpret_cast<uint16_t*>(dst)[plane + i] = lo;
}

// f16x2 operand planes -> fp32: (hi + lo) / FS2_F16X2_ACT_SCALE
__global__ void unsplit2_kernel(const bf16* src, int64_t n, int64_t plane_elems, float* dst) {
  FS2_PDL_PROLOGUE();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
  const float hi = __half2float(__ushort_as_half(s16[i])), lo = __half2float(__ushort_as_half(s16[plane_elems + i]));
  dst[i] = (hi + lo) * (1.0f / FS2_F16X2_ACT_SCALE);
}

__global__ void fill_zero_kernel(uint32_t* p, size_t n) {
  FS2_PDL_PROLOGUE();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}

__global__ void f32_to_bf16_kernel(const float* src, int64_t n, bf16* dst) {
  FS2_PDL_PROLOGUE();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* src, int64_t n, float* dst) {
  FS2_PDL_PROLOGUE();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}

// V [R, D] bf16 (flat rows) -> V^T [D, Rv]: row c, column r (columns >= R zero).  Test helper for the tcgen05
// attention entry point; on the product path the QKV GEMM epilogue writes V^T directly.
__global__ void transpose_v_kernel(const bf16* v, int R, int Rv, int D, bf16* vt) {
  FS2_PDL_PROLOGUE();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)D * Rv) return;
  const int r = (int)(i % Rv), c = (int)(i / Rv);
  vt[i] = r < R ? v[(size_t)r * D + c] : __float2bfloat16_rn(0.f);
}

// Models.py:231-233 alone: x[b,p,:] += pe[p,:] on grid rows; used when no variance embedding is frame-level
__global__ void add_pe_kernel(float* x, const float* pe, const RowLayout lay, int D) {
  FS2_PDL_PROLOGUE();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = D >> 2;
  if (i >= (size_t)lay.R_cap * nv) return;
  const size_t row = i / nv;
  const int c = (int)(i - row * nv);
  const RowPos rp = row_pos(lay, (int)row, ld_act(lay.off + lay.B));
  if (!rp.in_grid) return;
  float4 a = *reinterpret_cast<const float4*>(x + row * D + c * 4);
  const float4 q = ld4(pe + (size_t)rp.p * D + c * 4);
  a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
  st4(x + row * D + c * 4, a);
}

inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

#define LAUNCHED() (++g_fs2_launches, cudaGetLastError())

cudaError_t rowops_build_layout(const int* lens32, int B, int S, int halo_keep, int halo_rows, int* off, int* ext,
                                unsigned* rowmap, int R_cap, cudaStream_t st, int extra_ext, const int* S_dev) {
  if (B <= 0 || R_cap <= 0) return cudaSuccess;
  const int Bt = B + (extra_ext > 0 ? 1 : 0);
  if (Bt > 65535 || S > FS2_MAX_ROWS_PER_UTT) return cudaErrorInvalidValue;
  (void)FS2_LAUNCH(build_layout_kernel, 1, 1024, 0, st, lens32, B, S, halo_keep, halo_rows, off, ext, extra_ext, S_dev);
  ++g_fs2_launches;
  (void)FS2_LAUNCH(fill_rowmap_kernel, blocks_for((size_t)R_cap, 256), 256, 0, st, off, ext, Bt, R_cap, rowmap);
  return LAUNCHED();
}
cudaError_t rowops_embed_pe(const int64_t* texts, const float* emb, const float* pe, int vocab, const RowLayout& lay,
                            int D, float* out_grid, bf16* out_b, int out_planes, float* out_user, cudaStream_t st) {
  if (lay.R_cap <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(embed_pe_kernel, blocks_for((size_t)lay.R_cap, 8), 256, 0, st, texts, emb, pe, vocab, lay, D, out_grid, out_b,
                                                                   out_planes, out_user);
  return LAUNCHED();
}
cudaError_t rowops_lens_to_i32(const int64_t* lens, int B, int cap, int* out, cudaStream_t st, int* zero2, const int* cap_dev) {
  if (B <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(lens_to_i32_kernel, blocks_for(B, 256), 256, 0, st, lens, B, cap, out, zero2, cap_dev);
  return LAUNCHED();
}
cudaError_t rowops_mask(const int64_t* lens64, const int* lens32, int B, int max_len, uint8_t* mask, cudaStream_t st,
                        float* zero0, float* zero1, const int* max_len_dev) {
  if ((size_t)B * max_len == 0) return cudaSuccess;
  (void)FS2_LAUNCH(mask_kernel, blocks_for((size_t)B * max_len, 256), 256, 0, st, lens64, lens32, B, max_len, mask, zero0, zero1, max_len_dev);
  return LAUNCHED();
}
cudaError_t rowops_round_durations(const float* log_d, int64_t n, float d_control, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(round_durations_kernel, blocks_for((size_t)n, 256), 256, 0, st, log_d, n, d_control, out);
  return LAUNCHED();
}
cudaError_t rowops_duration_scan(const float* d, int B, int L, int* cum, int64_t* mel_lens, int* mel_lens32,
                                 int* tmax_dev, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(duration_scan_kernel, B, 1024, 0, st, (const float*)nullptr, 1.0f, const_cast<float*>(d), L, cum, mel_lens,
                   mel_lens32, tmax_dev, (const int*)nullptr);
  return LAUNCHED();
}
cudaError_t rowops_round_scan(const float* log_d, float d_control, float* d_rounded, int B, int L, int* cum,
                              int64_t* mel_lens, int* mel_lens32, int* tmax_dev, cudaStream_t st, const int* L_dev) {
  if (B <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(duration_scan_kernel, B, 1024, 0, st, log_d, d_control, d_rounded, L, cum, mel_lens, mel_lens32, tmax_dev, L_dev);
  return LAUNCHED();
}
cudaError_t rowops_length_regulate(const float* x, const int* src_off, int src_stride, const int* cum, int L, int D,
                                   const RowLayout& lay, float* out, bf16* out_b, int out_planes, cudaStream_t st,
                                   const int* L_dev) {
  if (lay.B <= 0 || lay.R_cap <= 0) return cudaSuccess;
  const size_t smem = sizeof(int) * (size_t)(L > 0 ? L : 1);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  dim3 grid((FS2_ROWS_PER_UTT(lay.S, FS2_HALO) + 63) / 64, lay.B);   // covers ext + halo rows of the longest utterance
  (void)FS2_LAUNCH(length_regulate_kernel, grid, 256, smem, st, x, src_off, src_stride, cum, L, D, lay, out, out_b, out_planes, L_dev);
  return LAUNCHED();
}
cudaError_t rowops_variance_embed(float* pred, float control, const float* bins, int n_bins, const float* emb,
                                  const float* pe, float* x, bf16* xb, int xb_planes, const RowLayout& lay, int D,
                                  int* idx_out, cudaStream_t st) {
  if (lay.R_cap <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(variance_embed_kernel, blocks_for((size_t)lay.R_cap, 8), 256, 0, st, pred, control, bins, n_bins, emb, pe, x, xb,
                                                                         xb_planes, lay, D, idx_out);
  return LAUNCHED();
}
cudaError_t rowops_fill_padded_rows(const float* bias, int N, const RowLayout& src, const RowLayout& dst, float* out_grid,
                                    bf16* out_b, int out_planes, float* out_user, cudaStream_t st) {
  if (dst.B <= 0 || dst.S <= 0) return cudaSuccess;
  dim3 grid((FS2_ROWS_PER_UTT(dst.S, FS2_HALO) + 7) / 8, dst.B);
  (void)FS2_LAUNCH(fill_padded_rows_kernel, grid, 256, 0, st, bias, N, src, dst, out_grid, out_b, out_planes, out_user);
  return LAUNCHED();
}
cudaError_t rowops_postnet_far_rows(const float* post_grid, int N, const RowLayout& pn, int B, int H, float* out_user,
                                    int user_cm, cudaStream_t st) {
  if (B <= 0 || pn.S <= 0) return cudaSuccess;
  if (user_cm) {
    dim3 grid_cm((pn.S + 255) / 256, B);
    (void)FS2_LAUNCH(postnet_far_rows_cm_kernel, grid_cm, 256, 0, st, post_grid, N, pn, B, H, out_user);
    return LAUNCHED();
  }
  dim3 grid((pn.S + 7) / 8, B);
  (void)FS2_LAUNCH(postnet_far_rows_kernel, grid, 256, 0, st, post_grid, N, pn, B, H, out_user);
  return LAUNCHED();
}
cudaError_t rowops_pack_weight_f16x2(const float* wf, int N, int K, int taps, bf16* dst, float* w_scale_out, cudaStream_t st) {
  const size_t n = (size_t)N * K * taps;
  if (n == 0) return cudaSuccess;
  unsigned* d_max = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_max), sizeof(unsigned));
  if (e != cudaSuccess) return e;
  unsigned bits = 0;
  e = cudaMemsetAsync(d_max, 0, sizeof(unsigned), st);
  if (e == cudaSuccess) {
    const int blocks = (int)(blocks_for(n, 256) < 1024 ? blocks_for(n, 256) : 1024);
    (void)FS2_LAUNCH(absmax_kernel, blocks, 256, 0, st, wf, n, d_max);
    ++g_fs2_launches;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bits, d_max, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_max);
  if (e != cudaSuccess) return e;
  float mx;
  memcpy(&mx, &bits, sizeof mx);
  // largest power of two with max|w| * scale < 16384 (fp16 max is 65504: 4x head-room), clamped to a sane range
  int ex = 0;
  if (mx > 0.f) {
    (void)frexpf(mx, &ex);          // mx = m * 2^ex, m in [0.5, 1)
    ex = 14 - ex;
  }
  if (ex > 24) ex = 24;
  if (ex < -24) ex = -24;
  const float scale = ldexpf(1.0f, ex);
  *w_scale_out = scale;
  (void)FS2_LAUNCH(pack_weight_f16x2_kernel, blocks_for(n, 256), 256, 0, st, wf, N, K, taps, scale, dst);
  return LAUNCHED();
}
cudaError_t rowops_split(const float* src, int64_t n, int planes, bf16* dst, int64_t plane_elems, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (n % 4) return cudaErrorInvalidValue;
  (void)FS2_LAUNCH(split_kernel, blocks_for((size_t)(n / 4), 256), 256, 0, st, src, n / 4, planes, dst, plane_elems);
  return LAUNCHED();
}
cudaError_t rowops_zero_first_rows(const RowLayout& lay, int C, float* x, cudaStream_t st) {
  if (lay.B <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(zero_first_rows_kernel, blocks_for((size_t)lay.B, 8), 256, 0, st, lay, C, x);
  return LAUNCHED();
}
cudaError_t rowops_to_grid(const float* x_user, const RowLayout& lay, int C, float* out, int ldo, int col_off,
                           bf16* out_b, cudaStream_t st) {
  const size_t n = (size_t)lay.R_cap * (C / 4);
  if (n == 0) return cudaSuccess;
  (void)FS2_LAUNCH(to_grid_kernel, blocks_for(n, 256), 256, 0, st, x_user, lay, C, out, ldo, col_off, out_b);
  return LAUNCHED();
}
cudaError_t rowops_from_grid(const float* x_grid, const RowLayout& lay, int C, float* out_user, cudaStream_t st) {
  const size_t n = (size_t)lay.B * lay.S * (C / 4);
  if (n == 0) return cudaSuccess;
  (void)FS2_LAUNCH(from_grid_kernel, blocks_for(n, 256), 256, 0, st, x_grid, lay, C, out_user);
  return LAUNCHED();
}
cudaError_t rowops_pack_weight(const float* src, int N, int K, int taps, const float* scale, float* dst_f,
                               bf16* dst_b, int n_total, int n_off, cudaStream_t st) {
  const size_t n = (size_t)N * K * taps;
  if (n == 0) return cudaSuccess;
  (void)FS2_LAUNCH(pack_weight_kernel, blocks_for(n, 256), 256, 0, st, src, N, K, taps, scale, dst_f, dst_b, n_total, n_off);
  return LAUNCHED();
}
cudaError_t rowops_bn_fold(const float* conv_bias, const float* g, const float* b, const float* mean,
                           const float* var, int n, float eps, float* scale_out, float* bias_out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(bn_fold_kernel, blocks_for(n, 256), 256, 0, st, conv_bias, g, b, mean, var, n, eps, scale_out, bias_out);
  return LAUNCHED();
}
cudaError_t rowops_f32_to_bf16(const float* src, int64_t n, bf16* dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(f32_to_bf16_kernel, blocks_for((size_t)n, 256), 256, 0, st, src, n, dst);
  return LAUNCHED();
}
cudaError_t rowops_unsplit2(const bf16* src, int64_t n, int64_t plane_elems, float* dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(unsplit2_kernel, blocks_for((size_t)n, 256), 256, 0, st, src, n, plane_elems, dst);
  return LAUNCHED();
}
cudaError_t rowops_bf16_to_f32(const bf16* src, int64_t n, float* dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  (void)FS2_LAUNCH(bf16_to_f32_kernel, blocks_for((size_t)n, 256), 256, 0, st, src, n, dst);
  return LAUNCHED();
}
cudaError_t rowops_transpose_v(const bf16* v, int R, int Rv, int D, bf16* vt, cudaStream_t st) {
  const size_t n = (size_t)D * Rv;
  if (n == 0) return cudaSuccess;
  (void)FS2_LAUNCH(transpose_v_kernel, blocks_for(n, 256), 256, 0, st, v, R, Rv, D, vt);
  return LAUNCHED();
}
cudaError_t rowops_add_pe(float* x, const float* pe, const RowLayout& lay, int D, cudaStream_t st) {
  const size_t n = (size_t)lay.R_cap * (D / 4);
  if (n == 0) return cudaSuccess;
  (void)FS2_LAUNCH(add_pe_kernel, blocks_for(n, 256), 256, 0, st, x, pe, lay, D);
  return LAUNCHED();
}
// zero fill as a KERNEL (not cudaMemsetAsync), so that it is an ordinary link of the programmatic-dependent-launch chain
cudaError_t rowops_fill_zero(void* p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  if (bytes % 4 || (reinterpret_cast<uintptr_t>(p) & 3)) { g_fs2_plain_next = 1; return cudaMemsetAsync(p, 0, bytes, st); }
  const size_t n = bytes / 4;
  (void)FS2_LAUNCH(fill_zero_kernel, blocks_for(n, 256) < 4096 ? blocks_for(n, 256) : 4096, 256, 0, st,
                   reinterpret_cast<uint32_t*>(p), n);
  return LAUNCHED();
}
