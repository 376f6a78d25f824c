// fs2_tc_gemm_staged.cu -- the tcgen05 implicit-GEMM of fs2_tc_gemm.cu with a shared-memory staged epilogue.
//
// Same mainloop (TMA producer warp, MMA warp, TMEM double-buffered accumulators, bf16 or bf16x3 operands) for the
// N % 256 == 0 GEMMs of the path; what changes is how the epilogue touches HBM.  With one thread per output row the
// direct version issues 16-byte accesses to 32 different 1-KB rows per warp instruction (32 L1 wavefronts each), which
// made the small-K GEMMs (attention output projection + LayerNorm, FFN conv k=1 + LayerNorm, QKV) epilogue-bound at
// 3-5x their HBM time (profiles/r1a).  Here every global access of the epilogue is a TMA transfer:
//   * the fp32 residual tile is TMA-loaded in [128 rows x 32 col] chunks into a 2-deep, 128B-swizzled staging ring
//     (prefetched while the mainloop of the tile is still running);
//   * results are written by the row-owning threads into swizzled staging tiles (conflict-free 16-byte st.shared) and
//     leave as TMA stores: fp32 [128 x 32], bf16 [128 x 32] per operand plane (1 or 3), Q / K tiles, and V^T as a
//     [32 d x 128 row] transposed tile.
// Row masking is by value (masked rows store zeros); rows past the end of the buffer are clipped by TMA.
// LayerNorm keeps the one-thread-per-row two-pass scheme (row parked in TMEM between the passes).
#include "fs2_tc_common.cuh"
#include "../../include/fs2_b200.h"

namespace {

using namespace tc;

constexpr int BM = 128, BN = 256, BKE = 64, STAGES = 3, NUM_THREADS = 192;
constexpr int A_BYTES = BM * BKE * 2, B_BYTES = BN * BKE * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int RING_BYTES = STAGES * STAGE_BYTES;          // 144 KB
constexpr int CH_F32 = BM * 32 * 4;                       // 16 KB: [128 rows][32 fp32], 128-byte rows, SWIZZLE_128B
constexpr int CH_B16 = BM * 32 * 2;                       //  8 KB: [128 rows][32 bf16],  64-byte rows, SWIZZLE_64B
constexpr int RES_OFF = RING_BYTES;                       // 2 x CH_F32
constexpr int OUTF_OFF = RES_OFF + 2 * CH_F32;            // CH_F32 (also the V^T chunk: [32 d][128 rows] bf16, 8 KB)
constexpr int OUTB_OFF = OUTF_OFF + CH_F32;               // 3 x CH_B16
constexpr int PARAM_OFF = OUTB_OFF + 3 * CH_B16;          // bias[256], ln_g[256], ln_b[256]
constexpr int BAR_OFF = PARAM_OFF + 3 * 256 * 4;
constexpr int NUM_BARS = 2 * STAGES + 4 + 2;
constexpr int SMEM_TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
constexpr uint32_t TMEM_COLS = 2 * BN;

__constant__ int s_combo_a[6] = {0, 2, 1, 0, 1, 0};       // bf16x3 cross products, smallest first (see fs2_tc_gemm.cu)
__constant__ int s_combo_b[6] = {2, 0, 1, 1, 0, 0};
__constant__ int s_combo2_a[3] = {0, 1, 0};               // f16x2 cross products (hi*lo, lo*hi, hi*hi)
__constant__ int s_combo2_b[3] = {1, 0, 0};

__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_conv_gemm_staged_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmOutF,
                           const __grid_constant__ CUtensorMap tmOutB0, const __grid_constant__ CUtensorMap tmOutB1,
                           const __grid_constant__ CUtensorMap tmVt, const ConvGemmArgs a, const int num_n_blocks) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  float* s_bias = reinterpret_cast<float*>(smem + PARAM_OFF);
  float* s_g = s_bias + 256;
  float* s_b = s_g + 256;
  const uint32_t bars = base + BAR_OFF;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
  auto res_bar = [&](int s) { return bars + 8u * (2 * STAGES + 4 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + NUM_BARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pad = (a.taps - 1) / 2;
  const int KB = (a.K + BKE - 1) / BKE;
  const int ncombo = a.planes == 3 ? 6 : a.planes == 2 ? 3 : 1;
  const int iters = a.taps * KB;
  const float asc = a.acc_scale;
  const bool ln = (a.epi == EPI_RES_LN || a.epi == EPI_RELU_LN);

  griddep_launch_dependents();   // PDL (fs2_common.cuh)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); mbar_init(res_bar(s), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    for (int i = t; i < 256; i += 128) {
      s_g[i] = ln ? __ldg(a.ln_g + i) : 1.f;
      s_b[i] = ln ? __ldg(a.ln_b + i) : 0.f;
    }
  }
  griddep_wait();   // first access to predecessor-written global memory is below
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int R = __ldg(a.lay.off + a.lay.B);
  const int num_tiles = ((R + BM - 1) / BM) * num_n_blocks;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
        const int r0 = m_blk * BM, n0 = n_blk * BN;
        for (int c = 0; c < ncombo; ++c) {
          const int pa = ncombo == 1 ? 0 : ncombo == 3 ? s_combo2_a[c] : s_combo_a[c];
          const int pb = ncombo == 1 ? 0 : ncombo == 3 ? s_combo2_b[c] : s_combo_b[c];
          for (int it = 0; it < iters; ++it) {
            const int t = it / KB, k0 = (it - t * KB) * BKE;
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            const uint32_t sa = base + stage * STAGE_BYTES;
            tma_load_3d(sa, &tmA, full_bar(stage), k0, r0 + t - pad, pa);
            tma_load_2d(sa + A_BYTES, &tmB, full_bar(stage), k0, (pb * a.taps + t) * a.N + n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16kind(BM, BN, a.planes == 2 ? 0u : 1u);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      const int steps = iters * ncombo;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int it = 0; it < steps; ++it) {
          mbar_wait(full_bar(stage), phase);
          fence_after_sync();
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BKE / 16; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5, 128 threads, thread = output row)
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                  // row in tile
    const bool elected = threadIdx.x == 64;
    const int sw128 = row & 7, sw64 = (row >> 1) & 3;
    uint8_t* resbuf = smem + RES_OFF;
    uint8_t* outf_row = smem + OUTF_OFF + row * 128;
    uint8_t* outb_row = smem + OUTB_OFF + row * 64;
    const bool has_res = a.epi == EPI_RES_LN;
    const int out_planes = a.out_planes == 3 ? 3 : a.out_planes == 2 ? 2 : 1;
    int res_cnt = 0;                                // residual chunks consumed so far by this thread (buffer / parity)
    int as = 0; uint32_t aphase = 0;

    // Stage one 32-column chunk of this thread's row and ship it: fp32 tile via mapF (if wf), bf16 plane tiles via mapB
    // (if wb).  Two 128-thread barriers: [A] the previous TMA stores have finished reading the staging tiles,
    // [B] all rows of the new tiles are written and visible to the async proxy.
    auto stage_out = [&](const float (&y)[32], int col, int r0, bool wf, bool wb, const CUtensorMap* mapB) {
      if (elected) tma_store_wait_read();
      bar_epi();
      if (wf) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(outf_row + ((j ^ sw128) << 4)) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
      }
      if (wb) {
        if (out_planes == 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(outb_row + ((j ^ sw64) << 4)) =
                make_uint4(pack_bf16x2(y[8 * j], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                           pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
        } else if (out_planes == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) split2h_pair(y[8 * j + 2 * u], y[8 * j + 2 * u + 1], h[u], l[u]);
            uint8_t* o = outb_row + ((j ^ sw64) << 4);
            *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(o + CH_B16) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float h[8], m[8], l[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) split3(y[8 * j + u], h[u], m[u], l[u]);
            uint8_t* o = outb_row + ((j ^ sw64) << 4);
            *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
            *reinterpret_cast<uint4*>(o + CH_B16) = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
            *reinterpret_cast<uint4*>(o + 2 * CH_B16) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
          }
        }
      }
      fence_proxy_async();
      bar_epi();
      if (elected) {
        if (wf) tma_store_2d(&tmOutF, base + OUTF_OFF, col, r0);
        if (wb)
          for (int p = 0; p < out_planes; ++p) tma_store_3d(mapB, base + OUTB_OFF + p * CH_B16, col, r0, p);
        tma_store_commit();
      }
    };

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
      const int n0 = n_blk * BN, r0 = m_blk * BM;
      // residual prefetch (overlaps the tile's mainloop): chunks 0 and 1
      if (has_res && elected) {
        for (int c = 0; c < 2; ++c) {
          const int buf = (res_cnt + c) & 1;
          mbar_expect_tx(res_bar(buf), CH_F32);
          tma_load_2d(base + RES_OFF + buf * CH_F32, &tmRes, res_bar(buf), c * 32, r0);
        }
      }
      // bias slice of this tile
      bar_epi();
      for (int i = threadIdx.x - 64; i < BN; i += 128) s_bias[i] = (n0 + i < a.N) ? __ldg(a.bias + n0 + i) : 0.f;
      bar_epi();

      const int r = r0 + row;
      const RowPos rp = row_pos(a.lay, r, R);
      const bool keep_len = rp.in_grid && (a.lay.lens == nullptr || rp.p < __ldg(a.lay.lens + rp.b));
      const bool keep = (a.mask_mode == MASK_LEN) ? keep_len : rp.in_grid;

      mbar_wait(tfull_bar(as), aphase);
      fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      uint32_t v[32];

      if (ln) {
        // ---- pass 1: pre-norm value, row statistics, park the row back in TMEM
        float sum = 0.f, sq = 0.f;
        for (int c = 0; c < BN / 32; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float rr[32];
          if (has_res) {
            const int buf = res_cnt & 1;
            mbar_wait(res_bar(buf), (uint32_t)((res_cnt >> 1) & 1));
            const uint8_t* rrow = resbuf + buf * CH_F32 + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t4 = *reinterpret_cast<const float4*>(rrow + ((j ^ sw128) << 4));
              rr[4 * j] = t4.x; rr[4 * j + 1] = t4.y; rr[4 * j + 2] = t4.z; rr[4 * j + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = fmaf(__uint_as_float(v[j]), asc, s_bias[c * 32 + j]) + rr[j];
            if (!has_res) x = fmaxf(x, 0.f);
            sum += x;
            sq = fmaf(x, x, sq);
            v[j] = __float_as_uint(x);
          }
          tmem_st32(t_row + c * 32, v);
          if (has_res) {
            bar_epi();                               // every row of this staging buffer has been read
            if (elected && c + 2 < BN / 32) {
              const int buf = res_cnt & 1;
              mbar_expect_tx(res_bar(buf), CH_F32);
              tma_load_2d(base + RES_OFF + buf * CH_F32, &tmRes, res_bar(buf), (c + 2) * 32, r0);
            }
            ++res_cnt;
          }
        }
        tmem_wait_st();
        const float mean = sum * (1.0f / 256.0f);
        const float var = fmaxf(sq * (1.0f / 256.0f) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        // ---- pass 2: normalise, affine, mask, stage + store
        for (int c = 0; c < BN / 32; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = (__uint_as_float(v[j]) - mean) * rstd * s_g[c * 32 + j] + s_b[c * 32 + j];
            y[j] = keep ? x : 0.f;
          }
          stage_out(y, c * 32, r0, a.out != nullptr, a.out_b != nullptr, &tmOutB0);
        }
      } else if (a.epi == EPI_QKV) {
        // n_blk 0 -> Q, 1 -> K (bf16 [R,256] tiles); 2 -> V transposed: vt[d, flat row]
        for (int c = 0; c < BN / 32; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = rp.in_grid ? fmaf(__uint_as_float(v[j]), asc, s_bias[c * 32 + j]) : 0.f;
          if (n_blk < 2) {
            stage_out(y, c * 32, r0, false, true, n_blk == 0 ? &tmOutB0 : &tmOutB1);
          } else {
            if (elected) tma_store_wait_read();
            bar_epi();
            bf16* vt_s = reinterpret_cast<bf16*>(smem + OUTF_OFF);       // [32 d][128 rows]
#pragma unroll
            for (int j = 0; j < 32; ++j) vt_s[j * BM + row] = __float2bfloat16_rn(y[j]);
            fence_proxy_async();
            bar_epi();
            if (elected) {
              tma_store_2d(&tmVt, base + OUTF_OFF, r0, c * 32);
              tma_store_commit();
            }
          }
        }
      } else {
        // ---- EPI_BIAS / EPI_RELU / EPI_TANH
        for (int c = 0; c < BN / 32; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = fmaf(__uint_as_float(v[j]), asc, s_bias[c * 32 + j]);
            if (a.epi == EPI_RELU) x = fmaxf(x, 0.f);
            if (a.epi == EPI_TANH) x = tanhf(x);
            y[j] = keep ? x : 0.f;
          }
          stage_out(y, n0 + c * 32, r0, a.out != nullptr, a.out_b != nullptr, &tmOutB0);
        }
      }
      // release the accumulator stage to the MMA warp
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (elected) tma_store_wait_all();   // staging tiles must outlive the last TMA store
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

bool tc_conv_gemm_staged_supported(const ConvGemmArgs& a) {
  if (a.N % 256 != 0 || a.dst_SA > 0 || a.out_user != nullptr) return false;
  if (a.epi == EPI_RES_LN || a.epi == EPI_RELU_LN) return a.N == 256 && (a.out == nullptr || a.ldo == 256) && (a.out_b == nullptr || a.ldob == 256);
  if (a.epi == EPI_QKV) return a.N == 768;
  if (a.epi == EPI_BIAS || a.epi == EPI_RELU || a.epi == EPI_TANH)
    return (a.out == nullptr || a.ldo == a.N) && (a.out_b == nullptr || a.ldob == a.N) && (a.out || a.out_b);
  return false;
}

int tc_conv_gemm_staged_launch(const ConvGemmArgs& a, cudaStream_t st) {
  const uint64_t R = (uint64_t)a.lay.R_cap;
  const int planes = a.planes == 3 ? 3 : a.planes == 2 ? 2 : 1;
  const int out_planes = a.out_planes == 3 ? 3 : a.out_planes == 2 ? 2 : 1;
  CUtensorMap tmA, tmB, tmRes, tmOutF, tmOutB0, tmOutB1, tmVt;
  if (!make_tmap_bf16_3d(&tmA, a.Ab, (uint64_t)planes, R, (uint64_t)a.K, (uint64_t)a.K, R * a.K, BM) ||
      !make_tmap_bf16(&tmB, a.Wb, (uint64_t)planes * a.taps * a.N, (uint64_t)a.K, (uint64_t)a.K, BN))
    return fs2_fail_cuda(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(A/W)");
  bool ok = true;
  auto f32_map = [&](CUtensorMap* m, const float* p, int ld) {
    const uint64_t dims[2] = {(uint64_t)ld, R}, strides[1] = {(uint64_t)ld * 4};
    const uint32_t box[2] = {32, (uint32_t)BM};
    return make_tmap_generic(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  auto b16_map = [&](CUtensorMap* m, const bf16* p, int ld, int np) {
    const uint64_t dims[3] = {(uint64_t)ld, R, (uint64_t)np}, strides[2] = {(uint64_t)ld * 2, R * ld * 2};
    const uint32_t box[3] = {32, (uint32_t)BM, 1};
    return make_tmap_generic(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
  };
  // unused maps alias a valid one (the kernel never dereferences them)
  ok = ok && f32_map(&tmRes, a.epi == EPI_RES_LN ? a.residual : (a.out ? a.out : reinterpret_cast<const float*>(a.Ab)), 256);
  ok = ok && f32_map(&tmOutF, a.out ? a.out : reinterpret_cast<const float*>(a.Ab), a.out ? a.ldo : 256);
  if (a.epi == EPI_QKV) {
    ok = ok && b16_map(&tmOutB0, a.q_b, 256, 1) && b16_map(&tmOutB1, a.k_b, 256, 1);
    const uint64_t dims[2] = {(uint64_t)a.Rv, 256}, strides[1] = {(uint64_t)a.Rv * 2};
    const uint32_t box[2] = {(uint32_t)BM, 32};
    ok = ok && make_tmap_generic(&tmVt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.vt_b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  } else {
    const bf16* ob = a.out_b ? a.out_b : a.Ab;
    ok = ok && b16_map(&tmOutB0, ob, a.out_b ? a.ldob : a.K, a.out_b ? out_planes : 1);
    tmOutB1 = tmOutB0;
    tmVt = tmOutB0;
  }
  if (!ok) return fs2_fail_cuda(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(staged epilogue)");
  static bool configured = false;
  const int smem = SMEM_TOTAL + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_gemm_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fs2_fail_cuda(e, "cudaFuncSetAttribute(tc_conv_gemm_staged)");
    configured = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int num_n_blocks = a.N / BN;
  const int tiles = (int)((R + BM - 1) / BM) * num_n_blocks;
  const int grid = tiles < num_sms ? tiles : num_sms;
  ConvGemmArgs b = a;
  if (a.epi == EPI_QKV) b.out_planes = 1;
  (void)FS2_LAUNCH(tc_conv_gemm_staged_kernel, grid, NUM_THREADS, smem, st, tmA, tmB, tmRes, tmOutF, tmOutB0, tmOutB1, tmVt, b, num_n_blocks);
  ++g_fs2_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fs2_fail_cuda(e, "tc_conv_gemm_staged_kernel launch");
  return FS2_OK;
}
