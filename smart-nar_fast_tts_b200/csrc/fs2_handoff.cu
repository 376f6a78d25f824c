// Hand-off of the forward's results to their consumers: the vocoder and the
// per-utterance slicing of `synth_samples` (reference utils/tools.py:153-199, utils/model.py:70-88).
//
// The reference returns padded tensors and then, per utterance, calls `.item()` on two lengths and slices / copies
// four tensors to the host one by one (B * 6 synchronising round trips per batch), and converts the vocoder's fp32
// waveforms to int16 on the host after copying the PADDED fp32 batch.  Here both are one pass over HBM on the device:
//   * pack_valid_rows: the valid rows of a padded [B, S, C] (or channel-major [B, C, S]) tensor, back to back, plus
//     the B+1 offsets -- the host then needs ONE copy of the valid data only;
//   * wav_to_int16: int16 = numpy's `(wav * max_wav_value).astype("int16")` of the first lens[b] samples of every row,
//     back to back: half the bytes of the fp32 batch, and no padding, cross the PCIe link.
// Pure byte movement: bound by HBM (read valid bytes once, write them once); every load and store is coalesced.
#include "fs2_common.cuh"

namespace {

__device__ __forceinline__ long long ld_act64(const int64_t* p) {
  long long v;
  asm volatile("ld.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld4(const float* p) { return ld_act(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ long long clamp_len(long long v, long long cap) { return v < 0 ? 0 : (v > cap ? cap : v); }

// off = sum_{i<b} clamp(lens[i], 0, cap); len = clamp(lens[b], 0, cap).  Every thread of the (256-thread) block gets
// both.  B is at most a few thousand, so each block re-reads the lengths before it instead of waiting for a scan
// kernel: no extra launch, no dependency between blocks.
__device__ __forceinline__ void segment_of(const int64_t* lens, int b, long long cap, long long* off, long long* len) {
  __shared__ long long s_part[8];
  long long acc = 0;
  for (int i = threadIdx.x; i < b; i += blockDim.x) acc += clamp_len(ld_act64(lens + i), cap);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  long long t = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += s_part[w];
  *off = t;
  *len = clamp_len(ld_act64(lens + b), cap);
}

// grid (chunks, B), 256 threads.  cm = 0: src [B, S, C] -> dst rows [off_b, off_b + len_b) of [sum len, C];
// cm = 1: src [B, C, S] -> utterance b's block dst + off_b * C holds [C, len_b] (channel-major per utterance).
__global__ void __launch_bounds__(256) pack_valid_rows_kernel(const float* src, const int64_t* lens, int B, int S, int C,
                                                              int cm, int vec, int64_t* offsets, float* dst) {
  FS2_PDL_PROLOGUE();
  const int b = blockIdx.y;
  long long off, len;
  segment_of(lens, b, S, &off, &len);
  if (blockIdx.x == 0 && threadIdx.x == 0 && offsets) {
    offsets[b] = off;
    if (b == B - 1) offsets[B] = off + len;
  }
  const size_t n = (size_t)len * C;
  float* d = dst + (size_t)off * C;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  if (!cm) {
    const float* s = src + (size_t)b * S * C;     // the valid rows of an utterance are one contiguous block
    if (vec) {
      const size_t n4 = n >> 2;
      size_t i = tid;
      for (; i + 3 * nthr < n4; i += 4 * nthr) {        // four independent 16-byte loads in flight per thread
        const float4 v0 = ld4(s + 4 * i), v1 = ld4(s + 4 * (i + nthr)), v2 = ld4(s + 4 * (i + 2 * nthr)),
                     v3 = ld4(s + 4 * (i + 3 * nthr));
        *reinterpret_cast<float4*>(d + 4 * i) = v0;
        *reinterpret_cast<float4*>(d + 4 * (i + nthr)) = v1;
        *reinterpret_cast<float4*>(d + 4 * (i + 2 * nthr)) = v2;
        *reinterpret_cast<float4*>(d + 4 * (i + 3 * nthr)) = v3;
      }
      for (; i < n4; i += nthr) *reinterpret_cast<float4*>(d + 4 * i) = ld4(s + 4 * i);
    } else {
      for (size_t i = tid; i < n; i += nthr) d[i] = ld_act(s + i);
    }
  } else {
    // utterance block [C, len]: element i = c * len + p comes from s + c*S + p.  32-bit index arithmetic (the division by
    // len is the only non-trivial instruction); four independent loads in flight per thread
    const float* s = src + (size_t)b * C * S;
    const unsigned n32 = (unsigned)n, l32 = (unsigned)len, t32 = (unsigned)tid, nt32 = (unsigned)nthr;
    auto at = [=](unsigned i) { const unsigned c = i / l32; return s + (size_t)c * S + (i - c * l32); };
    unsigned i = t32;
    if (n < 0x40000000ull) {
      for (; i + 3 * nt32 < n32; i += 4 * nt32) {
        const float v0 = ld_act(at(i)), v1 = ld_act(at(i + nt32)), v2 = ld_act(at(i + 2 * nt32)), v3 = ld_act(at(i + 3 * nt32));
        d[i] = v0; d[i + nt32] = v1; d[i + 2 * nt32] = v2; d[i + 3 * nt32] = v3;
      }
      for (; i < n32; i += nt32) d[i] = ld_act(at(i));
    } else {
      for (size_t j = tid; j < n; j += nthr) {
        const size_t c = j / (size_t)len;
        d[j] = ld_act(s + c * S + (j - c * (size_t)len));
      }
    }
  }
}

// numpy's float32 -> int16 `astype` on x86-64: cvttss2si to int32 (NaN / out of range -> INT32_MIN), low 16 bits kept
__device__ __forceinline__ short f32_to_i16_numpy(float x) {
  const int v = (fabsf(x) < 2147483648.0f) ? __float2int_rz(x) : (int)0x80000000;
  return (short)(v & 0xffff);
}

// grid (chunks, B), 256 threads.  wav [B, N] fp32; lens[b] samples kept (NULL: all N), written back to back.
__global__ void __launch_bounds__(256) wav_to_int16_kernel(const float* wav, const int64_t* lens, int B, long long N,
                                                           float scale, int vec_ok, int64_t* offsets, short* dst) {
  FS2_PDL_PROLOGUE();
  const int b = blockIdx.y;
  long long off = (long long)b * N, len = N;
  if (lens) segment_of(lens, b, N, &off, &len);
  if (blockIdx.x == 0 && threadIdx.x == 0 && offsets) {
    offsets[b] = off;
    if (b == B - 1) offsets[B] = off + len;
  }
  const float* s = wav + (size_t)b * N;
  short* d = dst + off;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (vec_ok && (off & 3) == 0) {               // 16-byte loads, 8-byte stores
    const size_t n4 = (size_t)len >> 2;
    auto cvt4 = [scale](const float4 v) {
      short4 o;
      o.x = f32_to_i16_numpy(__fmul_rn(v.x, scale));
      o.y = f32_to_i16_numpy(__fmul_rn(v.y, scale));
      o.z = f32_to_i16_numpy(__fmul_rn(v.z, scale));
      o.w = f32_to_i16_numpy(__fmul_rn(v.w, scale));
      return o;
    };
    size_t i = tid;
    for (; i + 3 * nthr < n4; i += 4 * nthr) {          // four independent 16-byte loads in flight per thread
      const float4 v0 = ld4(s + 4 * i), v1 = ld4(s + 4 * (i + nthr)), v2 = ld4(s + 4 * (i + 2 * nthr)),
                   v3 = ld4(s + 4 * (i + 3 * nthr));
      *reinterpret_cast<short4*>(d + 4 * i) = cvt4(v0);
      *reinterpret_cast<short4*>(d + 4 * (i + nthr)) = cvt4(v1);
      *reinterpret_cast<short4*>(d + 4 * (i + 2 * nthr)) = cvt4(v2);
      *reinterpret_cast<short4*>(d + 4 * (i + 3 * nthr)) = cvt4(v3);
    }
    for (; i < n4; i += nthr) *reinterpret_cast<short4*>(d + 4 * i) = cvt4(ld4(s + 4 * i));
    done = n4 << 2;
  }
  for (size_t i = done + tid; i < (size_t)len; i += nthr) d[i] = f32_to_i16_numpy(__fmul_rn(ld_act(s + i), scale));
}

// Blocks per utterance.  Every block pays a fixed prologue (it re-derives its utterance's offset from the lengths), so
// blocks should be few and fat: aim at ~8 blocks per SM over the whole grid, but keep >= min_per_thread work items per
// thread (small batches: parallelism first), at least 1 and at most 1024 chunks.
inline unsigned chunks_for(int B, size_t items_per_utt, size_t min_per_thread) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  size_t want = ((size_t)8 * sms + B - 1) / B;
  const size_t most = items_per_utt / (256 * min_per_thread);
  if (want > most) want = most;
  return (unsigned)(want < 1 ? 1 : (want > 1024 ? 1024 : want));
}

}  // namespace

cudaError_t handoff_pack_valid_rows(const float* src, const int64_t* lens, int B, int S, int C, int channel_major,
                                    int64_t* offsets, float* dst, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  const int vec = (!channel_major && (C % 4 == 0) && aligned) ? 1 : 0;
  dim3 grid(chunks_for(B, (size_t)S * C / (vec ? 4 : 1), 4), B);
  (void)FS2_LAUNCH(pack_valid_rows_kernel, grid, 256, 0, st, src, lens, B, S, C, channel_major, vec, offsets, dst);
  ++g_fs2_launches;
  return cudaGetLastError();
}

cudaError_t handoff_wav_to_int16(const float* wav, const int64_t* lens, int B, int64_t N, float max_wav_value,
                                 int64_t* offsets, int16_t* dst, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  const bool aligned = (reinterpret_cast<uintptr_t>(wav) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0;
  const int vec_ok = (aligned && N % 4 == 0) ? 1 : 0;
  dim3 grid(chunks_for(B, (size_t)N / 4, 4), B);
  (void)FS2_LAUNCH(wav_to_int16_kernel, grid, 256, 0, st, wav, lens, B, (long long)N, max_wav_value, vec_ok, offsets,
                   reinterpret_cast<short*>(dst));
  ++g_fs2_launches;
  return cudaGetLastError();
}
