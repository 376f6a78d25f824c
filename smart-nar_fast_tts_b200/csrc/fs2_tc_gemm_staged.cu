// fs2_tc_gemm_staged.cu -- the tcgen05 implicit-GEMM of fs2_tc_gemm.cu with a shared-memory staged epilogue.
//
// Same mainloop (TMA producer warp, MMA warp, TMEM double-buffered accumulators, bf16 / bf16x3 / f16x2 operands) for
// the N % 256 == 0 GEMMs of the path; what changes is how the epilogue touches HBM.  With one thread per output row the
// direct version issues 16-byte accesses to 32 different 1-KB rows per warp instruction (32 L1 wavefronts each), which
// made the small-K GEMMs (attention output projection + LayerNorm, FFN conv k=1 + LayerNorm, QKV) epilogue-bound at
// 3-5x their HBM time (profiles/r1a).  Here every global access of the epilogue is a TMA transfer:
//   * the fp32 residual tile is TMA-loaded in [128 rows x 32 col] chunks into 128B-swizzled staging buffers;
//   * results are written by the row-owning threads into swizzled staging tiles (conflict-free 16-byte st.shared) and
//     leave as TMA stores: fp32 [128 x 32], bf16 [128 x 32] per operand plane (1..3), Q / K tiles, and V^T as a
//     [32 d x 128 row] transposed tile.
// Epilogue v2 (profiles/r1d: the 4-warp epilogue spent ~45 % of its time waiting on the 2-deep residual ring and on the
// single output staging tile): EIGHT epilogue warps in two independent groups.  Group g owns columns [128 g, 128 g + 128)
// of the tile (TMEM restricts a warp to the lane quarter warp % 4, so more warps can only split columns); each group
// has its own residual / staging buffers, its own named barrier and its own elected TMA thread, so the two halves run
// decoupled and twice as many bytes are in flight.  A group's two fp32 buffers carry the residual chunks in pass 1
// and double-buffer the fp32 output tiles in pass 2.  LayerNorm needs one exchange per tile: the two threads that
// share a row swap their partial (sum, sum of squares) through shared memory (64-thread named barrier per lane quarter).
// Row masking is by value (masked rows store zeros); rows past the end of the buffer are clipped by TMA.
//
// Tile width (Plan::bn): 256 columns, or 128 for latency-bound launches that would leave most SMs idle (at batch 32 the
// encoder has 22 row tiles; N = 128 MMAs only reach ~2/3 of the N = 256 rate, so wide tiles stay the rule; see
// tc_conv_gemm_staged_launch and profiles/experiments/README.md).  A LayerNorm epilogue needs the whole 256-column row: with 128-wide tiles the two CTAs that share a row
// tile form a 2-CTA cluster and swap their partial (sum, sum of squares) through distributed shared memory
// (st.shared::cluster + remote mbarrier arrive with release / acquire at cluster scope), one exchange per tile.
#include "fs2_tc_common.cuh"
#include "../../include/fs2_b200.h"
#include <stdlib.h>

namespace {

using namespace tc;

constexpr int BM = 128, BN_MAX = 256, BKE = 64;
constexpr int NUM_THREADS = 352;                          // warp 0 TMA (A), warp 1 MMA, warp 2 TMA (B), warps 3-10 epilogue
constexpr int EPI_T0 = 96;                                // first epilogue thread
constexpr int A_BYTES = BM * BKE * 2;                     // 16 KB per k-block; the weight k-block is bn x 128 B (32 / 16 KB)
constexpr int MAX_A = 4, MAX_B = 4;                       // ring depths (a k-step consumes one stage of each ring)
constexpr int CH_F32 = BM * 32 * 4;                       // 16 KB: [128 rows][32 fp32], 128-byte rows, SWIZZLE_128B
constexpr int CH_B16 = BM * 32 * 2;                       //  8 KB: [128 rows][32 bf16],  64-byte rows, SWIZZLE_64B
// shared-memory plan (bytes from the 1024-aligned base), chosen per launch (struct Plan):
//   A ring          nA x 16 KB        (activation k-blocks, own producer warp)
//   B ring          nB x 32 / 16 KB   (weight k-blocks of a 256 / 128-wide tile, own producer warp)
//   group g region  wide plan: F0, F1 (2 x 16 KB fp32 staging / residual) + B staging (out_planes x 8 KB)
//                   deep plan: B staging only
//   params          bias[2][128], ln_g[256], ln_b[256], stats[2 parities][2 groups][128] float2,
//                   xstat[2 parities][128] float2 (the peer CTA's LayerNorm partial sums, written remotely)
//   barriers
constexpr int PARAM_BYTES = (2 * 128 + 256 + 256) * 4 + 2 * 2 * 128 * 8 + 2 * 128 * 8;
constexpr int NUM_BARS = 2 * MAX_A + 2 * MAX_B + 4 + 4 + 2;
struct Plan {
  int nA, nB, wide, grp_bytes;
  int serial_planes;   // deep plan with several operand planes out: the planes go through ONE 8 KB staging tile in turn
  int bn;              // tile width: 256 or 128 columns
  int ln_pair;         // LayerNorm over a 2-CTA cluster (bn == 128): partial row statistics are exchanged through DSMEM
  __host__ __device__ int b_bytes() const { return bn * BKE * 2; }
  __host__ __device__ int ring_bytes() const { return nA * A_BYTES + nB * b_bytes(); }
  __host__ __device__ int total() const { return ring_bytes() + 2 * grp_bytes + PARAM_BYTES + NUM_BARS * 8 + 16; }
};
constexpr uint32_t TMEM_COLS = 2 * BN_MAX;

__constant__ int s_combo_a[6] = {0, 2, 1, 0, 1, 0};       // bf16x3 cross products, smallest first (see fs2_tc_gemm.cu)
__constant__ int s_combo_b[6] = {2, 0, 1, 1, 0, 0};
__constant__ int s_combo2_a[3] = {0, 1, 0};               // f16x2 cross products (hi*lo, lo*hi, hi*hi)
__constant__ int s_combo2_b[3] = {1, 0, 0};

template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_conv_gemm_staged_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmOutF,
                           const __grid_constant__ CUtensorMap tmOutB0, const __grid_constant__ CUtensorMap tmOutB1,
                           const __grid_constant__ CUtensorMap tmVt, const ConvGemmArgs a, const int num_n_blocks,
                           const Plan plan) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const int nA = plan.nA, nB = plan.nB;
  const int BN = plan.bn, B_BYTES = plan.b_bytes();
  const uint32_t ringB = base + nA * A_BYTES;      // B ring follows the A ring
  const int ring_bytes = plan.ring_bytes();
  const int grp_sz = plan.grp_bytes;
  const int param_off = ring_bytes + 2 * grp_sz;
  float* s_bias = reinterpret_cast<float*>(smem + param_off);           // [2 groups][128]
  float* s_g = s_bias + 256;
  float* s_b = s_g + 256;
  float2* s_stat = reinterpret_cast<float2*>(s_b + 256);                // [2 parities][2 groups][128]
  float2* s_xstat = s_stat + 2 * 2 * 128;                               // [2 parities][128], written by the peer CTA
  const int bar_off = param_off + PARAM_BYTES;
  const uint32_t bars = base + bar_off;
  auto fullA = [&](int s) { return bars + 8u * s; };
  auto emptyA = [&](int s) { return bars + 8u * (MAX_A + s); };
  auto fullB = [&](int s) { return bars + 8u * (2 * MAX_A + s); };
  auto emptyB = [&](int s) { return bars + 8u * (2 * MAX_A + MAX_B + s); };
  constexpr int TB0 = 2 * MAX_A + 2 * MAX_B;
  auto tfull_bar = [&](int s) { return bars + 8u * (TB0 + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (TB0 + 2 + s); };
  auto res_bar = [&](int g, int s) { return bars + 8u * (TB0 + 4 + 2 * g + s); };
  auto xfull_bar = [&](int s) { return bars + 8u * (TB0 + 8 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + NUM_BARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pad = (a.taps - 1) / 2;
  const int KB = (a.K + BKE - 1) / BKE;
  const int ncombo = a.planes == 3 ? 6 : a.planes == 2 ? 3 : 1;
  const int iters = a.taps * KB;
  // pinned in a register: as a plain read of the parameter block the compiler re-loads it from the constant bank next to
  // every use inside the (issue-bound) epilogue loops -- 31 LDC per 32-column chunk in the r2l capture of dec.qkv
  float asc;
  asm volatile("mov.f32 %0, %1;" : "=f"(asc) : "f"(a.acc_scale));
  const bool ln = (a.epi == EPI_RES_LN || a.epi == EPI_RELU_LN);

  griddep_launch_dependents();   // PDL (fs2_common.cuh)
  if (base & 1023u) __trap();    // the swizzled tiles assume a 1024-byte aligned dynamic shared memory window
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < MAX_A; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < MAX_B; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }
    for (int g = 0; g < 2; ++g) for (int s = 0; s < 2; ++s) mbar_init(res_bar(g, s), 1);
    for (int s = 0; s < 2; ++s) mbar_init(xfull_bar(s), 128);   // one remote arrive per row of the peer CTA
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 3) {
    const int i = threadIdx.x - EPI_T0;   // 0..255
    s_g[i] = ln ? __ldg(a.ln_g + i) : 1.f;
    s_b[i] = ln ? __ldg(a.ln_b + i) : 0.f;
  }
  griddep_wait();   // first access to predecessor-written global memory is below
  fence_before_sync();
  __syncthreads();
  if (plan.ln_pair) cluster_sync_all();   // the peer's barriers exist before any remote arrive can reach them
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int R = ld_act(a.lay.off + a.lay.B);
  const int num_tiles = ((R + BM - 1) / BM) * num_n_blocks;

  if (warp == 0) {
    // ===================================================== TMA producer, activations (A ring)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int r0 = (tile / num_n_blocks) * BM;
        for (int c = 0; c < ncombo; ++c) {
          const int pa = ncombo == 1 ? 0 : ncombo == 3 ? s_combo2_a[c] : s_combo_a[c];
          for (int it = 0; it < iters; ++it) {
            const int t = it / KB, k0 = (it - t * KB) * BKE;
            mbar_wait(emptyA(stage), phase ^ 1u);
            mbar_expect_tx(fullA(stage), A_BYTES);
            tma_load_3d(base + stage * A_BYTES, &tmA, fullA(stage), k0, r0 + t - pad, pa);
            if (++stage == nA) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================== TMA producer, weights (B ring)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % num_n_blocks) * BN;
        for (int c = 0; c < ncombo; ++c) {
          const int pb = ncombo == 1 ? 0 : ncombo == 3 ? s_combo2_b[c] : s_combo_b[c];
          for (int it = 0; it < iters; ++it) {
            const int t = it / KB, k0 = (it - t * KB) * BKE;
            mbar_wait(emptyB(stage), phase ^ 1u);
            mbar_expect_tx(fullB(stage), B_BYTES);
            tma_load_2d(ringB + stage * B_BYTES, &tmB, fullB(stage), k0, (pb * a.taps + t) * a.N + n0);
            if (++stage == nB) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16kind(BM, BN, a.planes == 2 ? 0u : 1u);   // N = tile width (256 / 128)
      int sa_i = 0, sb_i = 0; uint32_t pha = 0, phb = 0;
      int as = 0; uint32_t aphase = 0;
      const int steps = iters * ncombo;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int it = 0; it < steps; ++it) {
          mbar_wait(fullB(sb_i), phb);
          mbar_wait(fullA(sa_i), pha);
          fence_after_sync();
          const uint64_t adesc = make_smem_desc_sw128(base + sa_i * A_BYTES);
          const uint64_t bdesc = make_smem_desc_sw128(ringB + sb_i * B_BYTES);
#pragma unroll
          for (int k = 0; k < BKE / 16; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) ? 1u : 0u);
          umma_commit(emptyA(sa_i));
          umma_commit(emptyB(sb_i));
          if (++sa_i == nA) { sa_i = 0; pha ^= 1u; }
          if (++sb_i == nB) { sb_i = 0; phb ^= 1u; }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================================================== epilogue: 2 groups x 4 warps, thread = (output row, column half)
    const int g = (warp - 3) >> 2;                  // group = column half
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                  // row in tile
    const bool elected = threadIdx.x == EPI_T0 + g * 128;
    const int bar_grp_id = 1 + g, bar_pair_id = 3 + q;
    auto bar_grp = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_grp_id) : "memory"); };
    const int sw128 = row & 7, sw64 = (row >> 1) & 3;
    const int grp_off = ring_bytes + g * grp_sz;
    const bool wide = plan.wide != 0;               // F buffers exist
    uint8_t* Fbuf = smem + grp_off;                 // F0 | F1 (wide plan only)
    uint8_t* Bbuf = smem + grp_off + (wide ? 2 * CH_F32 : 0);
    const bool has_res = a.epi == EPI_RES_LN;
    const int out_planes = a.out_planes == 3 ? 3 : a.out_planes == 2 ? 2 : 1;
    const int HALF = BN >> 1;                       // columns per group: 128 or 64
    const int NCH = HALF >> 5;                      // 32-column chunks per group: 4 or 2
    const int gc0 = g * HALF;                       // this group's first column within the tile
    float* my_bias = s_bias + g * 128;
    int res_cnt = 0;                                // residual chunks consumed so far by this thread (buffer / parity)
    int as = 0; uint32_t aphase = 0;
    uint32_t tile_par = 0;
    uint32_t xphase = 0;                            // bit i = phase of peer-statistics barrier i

    // 32 per-column parameters (bias / gamma / beta slice of a chunk) from shared memory as 8 vector loads (they were 32
    // scalar broadcast loads per chunk and parameter)
    auto ld_cols = [&](const float* p, float (&o)[32]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t4 = *reinterpret_cast<const float4*>(p + 4 * j);
        o[4 * j] = t4.x; o[4 * j + 1] = t4.y; o[4 * j + 2] = t4.z; o[4 * j + 3] = t4.w;
      }
    };

    // Stage one 32-column chunk (index c within the group's half) of this thread's row and ship it.
    //   wf: fp32 tile via tmOutF, staged in F[c & 1]; wb: bf16 plane tiles via mapB, staged in B (or, when bF, in F[c & 1]).
    // Commit order is B group then F group, so that wait_group.read 1 (all but the newest group have been read out) frees
    // both B and F[c & 1] while the previous chunk's fp32 store may still be in flight.
    auto stage_out = [&](const float (&y)[32], int c, int gcol, int r0, bool wf, bool wb, bool bF, const CUtensorMap* mapB) {
      uint8_t* fb = Fbuf + (c & 1) * CH_F32;
      uint8_t* bb = bF ? fb : Bbuf;
      if (plan.serial_planes && wb && !wf && out_planes == 2) {
        // mainloop-bound producers with two operand planes out (encoder FFN conv k=9 in f16x2): the staging tile is kept to
        // 8 KB so that the weight ring gets a fourth stage; hi and lo plane leave one after the other
        uint32_t hh[16], ll[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) split2h_pair(y[2 * u], y[2 * u + 1], hh[u], ll[u]);
        uint8_t* orow = Bbuf + row * 64;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          if (elected) tma_store_wait_read_n<0>();
          bar_grp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(orow + ((j ^ sw64) << 4)) =
                pl == 0 ? make_uint4(hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3])
                        : make_uint4(ll[4 * j], ll[4 * j + 1], ll[4 * j + 2], ll[4 * j + 3]);
          fence_proxy_async();
          bar_grp();
          if (elected) {
            tma_store_3d(mapB, base + (uint32_t)(Bbuf - smem), gcol, r0, pl);
            tma_store_commit();
          }
        }
        return;
      }
      if (elected) {
        if (wf || bF) tma_store_wait_read_n<1>(); else tma_store_wait_read_n<0>();
      }
      bar_grp();
      if (wf) {
        uint8_t* o = fb + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(o + ((j ^ sw128) << 4)) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
      }
      if (wb) {
        uint8_t* orow = bb + row * 64;
        if (out_planes == 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(orow + ((j ^ sw64) << 4)) =
                make_uint4(pack_bf16x2(y[8 * j], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                           pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
        } else if (out_planes == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) split2h_pair(y[8 * j + 2 * u], y[8 * j + 2 * u + 1], h[u], l[u]);
            uint8_t* o = orow + ((j ^ sw64) << 4);
            *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(o + CH_B16) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float h[8], m[8], l[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) split3(y[8 * j + u], h[u], m[u], l[u]);
            uint8_t* o = orow + ((j ^ sw64) << 4);
            *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
            *reinterpret_cast<uint4*>(o + CH_B16) = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
            *reinterpret_cast<uint4*>(o + 2 * CH_B16) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
          }
        }
      }
      fence_proxy_async();
      bar_grp();
      if (elected) {
        if (wb) {
          const uint32_t bsm = base + (uint32_t)(bb - smem);
          for (int p = 0; p < out_planes; ++p) tma_store_3d(mapB, bsm + p * CH_B16, gcol, r0, p);
          tma_store_commit();
        }
        if (wf) {
          tma_store_2d(&tmOutF, base + (uint32_t)(fb - smem), gcol, r0);
          tma_store_commit();
        }
      }
    };

    // (b, p) code and valid length of this thread's row, fetched AHEAD of the tile that needs them: the code two tiles
    // ahead, the length (a load that depends on the code) one tile ahead.  Fetched at the point of use, the dependent pair
    // cost every tile a full global-memory round trip at the top of its epilogue (8 % of the stall samples of dec.qkv).
    auto code_of_tile = [&](int t) -> unsigned {
      const int rn = (t / num_n_blocks) * BM + row;
      return (t < num_tiles && rn < R) ? ld_act(a.lay.rowmap + rn) : FS2_ROW_NONE;
    };
    auto len_of_code = [&](unsigned cd) -> int {
      return (cd != FS2_ROW_NONE && a.lay.lens != nullptr) ? ld_act(a.lay.lens + (int)(cd >> 16)) : 0x7fffffff;
    };
    unsigned next_code = code_of_tile((int)blockIdx.x);
    unsigned next2_code = code_of_tile((int)blockIdx.x + (int)gridDim.x);
    int next_len = len_of_code(next_code);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
      const int n0 = n_blk * BN, r0 = m_blk * BM;
      const float* my_g = s_g + (n0 & 255) + gc0;    // LayerNorm parameters of this thread's columns (N == 256)
      const float* my_b = s_b + (n0 & 255) + gc0;
      // residual prefetch: chunks 0 and 1 of this group's half (the F buffers double as output staging in pass 2 of the
      // previous tile: wait until those stores have been read out)
      if (has_res && elected) {
        tma_store_wait_read_n<0>();
        for (int c = 0; c < 2; ++c) {
          const int buf = (res_cnt + c) & 1;
          mbar_expect_tx(res_bar(g, buf), CH_F32);
          tma_load_2d(base + grp_off + buf * CH_F32, &tmRes, res_bar(g, buf), n0 + gc0 + c * 32, r0);
        }
      }
      // bias slice of this group's half of the tile
      bar_grp();
      {
        const int i = threadIdx.x - EPI_T0 - g * 128;
        if (i < HALF) my_bias[i] = (n0 + gc0 + i < a.N) ? __ldg(a.bias + n0 + gc0 + i) : 0.f;
      }
      bar_grp();

      // (b, p) of this thread's row; the code of the next tile's row is fetched now so its latency is off the critical path
      const unsigned code = next_code;
      const int row_len = next_len;
      next_code = next2_code;                                   // loaded one tile ago
      next2_code = code_of_tile(tile + 2 * (int)gridDim.x);
      next_len = len_of_code(next_code);
      const bool in_grid = code != FS2_ROW_NONE;
      const int rpp = in_grid ? (int)(code & 0xFFFFu) : 0;
      const bool keep_len = in_grid && rpp < row_len;
      const bool keep = (a.mask_mode == MASK_LEN) ? keep_len : in_grid;

      mbar_wait(tfull_bar(as), aphase);
      fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + gc0);
      uint32_t v[32];

      if (ln) {
        // ---- pass 1: pre-norm value, partial row statistics, park the row back in TMEM
        // partial sums over 64-column quarters of the row, combined below in ONE fixed association order
        // ((q0 + q1) + (q2 + q3)) whatever the tile width, so that the statistics -- and with them every output bit -- do
        // not depend on how the launcher tiled the GEMM (batch-size invariance, tests/test_gpu_properties.py)
        float psum[2] = {0.f, 0.f}, psq[2] = {0.f, 0.f};
        for (int c = 0; c < NCH; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float rr[32];
          if (has_res) {
            const int buf = res_cnt & 1;
            mbar_wait(res_bar(g, buf), (uint32_t)((res_cnt >> 1) & 1));
            const uint8_t* rrow = smem + grp_off + buf * CH_F32 + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t4 = *reinterpret_cast<const float4*>(rrow + ((j ^ sw128) << 4));
              rr[4 * j] = t4.x; rr[4 * j + 1] = t4.y; rr[4 * j + 2] = t4.z; rr[4 * j + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = 0.f;
          }
          float bb[32];
          ld_cols(my_bias + c * 32, bb);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = fmaf(__uint_as_float(v[j]), asc, bb[j]) + rr[j];
            if (!has_res) x = fmaxf(x, 0.f);
            psum[c >> 1] += x;
            psq[c >> 1] = fmaf(x, x, psq[c >> 1]);
            v[j] = __float_as_uint(x);
          }
          tmem_st32(t_row + c * 32, v);
          if (has_res) {
            bar_grp();                               // every row of this staging buffer has been read
            if (elected && c + 2 < NCH) {
              const int buf = res_cnt & 1;
              mbar_expect_tx(res_bar(g, buf), CH_F32);
              tma_load_2d(base + grp_off + buf * CH_F32, &tmRes, res_bar(g, buf), n0 + gc0 + (c + 2) * 32, r0);
            }
            ++res_cnt;
          }
        }
        tmem_wait_st();
        const float sum = NCH == 4 ? psum[0] + psum[1] : psum[0];
        const float sq = NCH == 4 ? psq[0] + psq[1] : psq[0];
        // the two threads of a row (one per group) exchange their partial sums
        s_stat[(tile_par * 2 + g) * 128 + row] = make_float2(sum, sq);
        asm volatile("bar.sync %0, 64;" ::"r"(bar_pair_id) : "memory");
        const float2 other = s_stat[(tile_par * 2 + (g ^ 1)) * 128 + row];
        // add in a fixed order (group 0 first) so that both threads of the row derive bit-identical statistics
        float tsum = g == 0 ? sum + other.x : other.x + sum;
        float tsq = g == 0 ? sq + other.y : other.y + sq;
        if (plan.ln_pair) {
          // 128-wide tiles: the other half of the row lives in the peer CTA of the cluster.  Group 0 ships this CTA's
          // partial sums into the peer's shared memory and arrives (release.cluster) on the peer's barrier; every
          // epilogue thread then waits (acquire.cluster) for the peer's partial.  Double-buffered by tile parity.
          const uint32_t peer = cluster_ctarank() ^ 1u;
          if (g == 0) {
            const uint32_t slot = base + (uint32_t)(reinterpret_cast<uint8_t*>(s_xstat + tile_par * 128 + row) - smem);
            st_remote_f32x2(slot, peer, tsum, tsq);
            mbar_arrive_remote_release(xfull_bar((int)tile_par), peer);
          }
          mbar_wait_cluster(xfull_bar((int)tile_par), (xphase >> tile_par) & 1u);
          xphase ^= 1u << tile_par;
          const float2 px = s_xstat[tile_par * 128 + row];
          // fixed order (rank 0's half first) so that both CTAs derive bit-identical statistics
          const bool first = (cluster_ctarank() == 0);
          tsum = first ? tsum + px.x : px.x + tsum;
          tsq = first ? tsq + px.y : px.y + tsq;
        }
        const float mean = tsum * (1.0f / 256.0f);
        const float var = fmaxf(tsq * (1.0f / 256.0f) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        // ---- pass 2: normalise, affine, mask, stage + store
        for (int c = 0; c < NCH; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float y[32], gg[32], be[32];
          ld_cols(my_g + c * 32, gg);
          ld_cols(my_b + c * 32, be);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = (__uint_as_float(v[j]) - mean) * rstd * gg[j] + be[j];
            y[j] = keep ? x : 0.f;
          }
          stage_out(y, c, n0 + gc0 + c * 32, r0, a.out != nullptr, a.out_b != nullptr, false, &tmOutB0);
        }
      } else if (a.epi == EPI_QKV) {
        // columns [0,256) -> Q, [256,512) -> K (bf16 [R,256] tiles); [512,768) -> V transposed: vt[d, flat row]
        const int part = n0 >> 8, pc0 = (n0 & 255) + gc0;      // Q / K / V and the first column inside that part
        for (int c = 0; c < NCH; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float y[32], bb[32];
          ld_cols(my_bias + c * 32, bb);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = fmaf(__uint_as_float(v[j]), asc, bb[j]);
            y[j] = in_grid ? x : 0.f;
          }
          if (part < 2) {
            stage_out(y, c, pc0 + c * 32, r0, false, true, true, part == 0 ? &tmOutB0 : &tmOutB1);
          } else {
            uint8_t* fb = Fbuf + (c & 1) * CH_F32;
            if (elected) tma_store_wait_read_n<1>();
            bar_grp();
            uint16_t* vt_s = reinterpret_cast<uint16_t*>(fb);           // [out_planes][32 d][128 rows]
            if (out_planes == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) vt_s[j * BM + row] = __bfloat16_as_ushort(__float2bfloat16_rn(y[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                uint16_t hi, lo;
                split2h_scaled(y[j] * FS2_F16X2_ACT_SCALE, hi, lo);
                vt_s[j * BM + row] = hi;
                vt_s[32 * BM + j * BM + row] = lo;
              }
            }
            fence_proxy_async();
            bar_grp();
            if (elected) {
              for (int pl = 0; pl < out_planes; ++pl)
                tma_store_3d(&tmVt, base + (uint32_t)(fb - smem) + pl * CH_B16, r0, pc0 + c * 32, pl);
              tma_store_commit();
            }
          }
        }
      } else {
        // ---- EPI_BIAS / EPI_RELU / EPI_TANH
        for (int c = 0; c < NCH; ++c) {
          __syncwarp();
          tmem_ld32(t_row + c * 32, v);
          tmem_wait_ld();
          float y[32], bb[32];
          ld_cols(my_bias + c * 32, bb);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = fmaf(__uint_as_float(v[j]), asc, bb[j]);
            if (a.epi == EPI_RELU) x = fmaxf(x, 0.f);
            if (a.epi == EPI_TANH) x = tanhf(x);
            y[j] = keep ? x : 0.f;
          }
          stage_out(y, c, n0 + gc0 + c * 32, r0, a.out != nullptr, a.out_b != nullptr, false, &tmOutB0);
        }
      }
      // release the accumulator stage to the MMA warp
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
      tile_par ^= 1u;
    }
    if (elected) tma_store_wait_all();   // staging tiles must outlive the last TMA store
  }

  fence_before_sync();
  __syncthreads();
  if (plan.ln_pair) cluster_sync_all();   // no remote store / arrive may target a CTA that has already exited
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Wide plan (fp32 staging buffers) whenever the epilogue moves fp32 tiles or V^T; deep plan for the bf16-only
// producers (FFN conv k=9, PostNet k=5, predictor conv1).  Ring depths fill what the staging leaves of the 227 KB.
Plan plan_for(const ConvGemmArgs& a, int out_planes, int bn) {
  Plan p;
  p.bn = bn;
  p.ln_pair = (bn == 128 && (a.epi == EPI_RES_LN || a.epi == EPI_RELU_LN)) ? 1 : 0;
  p.wide = (a.epi == EPI_RES_LN || a.epi == EPI_QKV || a.out != nullptr) ? 1 : 0;
  p.serial_planes = (!p.wide && out_planes == 2) ? 1 : 0;
  p.grp_bytes = (p.wide ? 2 * CH_F32 : 0) + (p.serial_planes ? 1 : out_planes) * CH_B16;
  // a k-step consumes one A and one B stage: the pipeline is as deep as the shallower ring.  Deepest B ring that
  // fits next to a full A ring, then shrink the A ring if even two B stages do not fit.
  p.nA = MAX_A;
  p.nB = MAX_B;
  while (p.nB > 2 && p.total() > 227 * 1024) --p.nB;
  while (p.nA > 2 && p.total() > 227 * 1024) --p.nA;
  return p;
}

}  // namespace

bool tc_conv_gemm_staged_supported(const ConvGemmArgs& a) {
  if (a.N % 256 != 0 || a.dst_off != nullptr || a.out_user != nullptr) return false;
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
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  // Tile width: 128 columns when that needs fewer rounds of the persistent grid than 256 (a round of 128-wide tiles costs
  // half a round of 256-wide ones): small launches that leave SMs idle, and row counts just past a multiple of 148 tiles.
  const uint64_t rows = a.lay.rows_hint > 0 ? (uint64_t)a.lay.rows_hint : R;
  const int m_tiles = (int)((rows + BM - 1) / BM);
  // Measured (profiles/r1n): N = 128 MMAs deliver ~2/3 of the N = 256 rate, so 128-wide tiles only pay where the launch
  // is latency-bound: the 256-wide tiles would occupy at most half of the SMs (twice as many CTAs then share the same
  // single round) and the mainloop is short.
  const int ksteps = a.taps * ((a.K + BKE - 1) / BKE) * (planes == 3 ? 6 : planes == 2 ? 3 : 1);
  static const char* bn_env = getenv("FS2_TILE_N");    // experiment switch: force 128 / 256
  // (5/8 rather than 1/2 of the SMs: rows_hint / R_cap over-estimate the rows in use, by up to ~1.4x for the encoder)
  int BN = (m_tiles * (a.N / 256) <= num_sms * 5 / 8 && ksteps <= 64) ? 128 : 256;
  if (bn_env) BN = atoi(bn_env) == 128 ? 128 : 256;
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
    const int qp = a.planes == 2 ? 2 : 1;           // Q, K, V^T operand planes for the attention kernel
    ok = ok && b16_map(&tmOutB0, a.q_b, 256, qp) && b16_map(&tmOutB1, a.k_b, 256, qp);
    const uint64_t dims[3] = {(uint64_t)a.Rv, 256, (uint64_t)qp}, strides[2] = {(uint64_t)a.Rv * 2, (uint64_t)a.Rv * 512};
    const uint32_t box[3] = {(uint32_t)BM, 32, 1};
    ok = ok && make_tmap_generic(&tmVt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a.vt_b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  } else {
    const bf16* ob = a.out_b ? a.out_b : a.Ab;
    ok = ok && b16_map(&tmOutB0, ob, a.out_b ? a.ldob : a.K, a.out_b ? out_planes : 1);
    tmOutB1 = tmOutB0;
    tmVt = tmOutB0;
  }
  if (!ok) return fs2_fail_cuda(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(staged epilogue)");
  static std::atomic<bool> configured{false};   // handles on several host threads may race here: benign, but formally atomic
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_gemm_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fs2_fail_cuda(e, "cudaFuncSetAttribute(tc_conv_gemm_staged)");
    configured = true;
  }
  const int num_n_blocks = a.N / BN;
  const int tiles = (int)((R + BM - 1) / BM) * num_n_blocks;   // upper bound (allocated rows); the kernel reads the real count
  int grid = tiles < num_sms ? tiles : num_sms;
  ConvGemmArgs b = a;
  if (a.epi == EPI_QKV) b.out_planes = a.planes == 2 ? 2 : 1;
  const Plan plan = plan_for(a, a.epi == EPI_QKV ? 1 : out_planes, BN);   // Q / K / V^T are staged in the fp32 buffers
  if (plan.total() > 227 * 1024) return fs2_fail_cuda(cudaErrorInvalidValue, "tc_conv_gemm_staged: shared-memory plan");
  if (plan.ln_pair) {
    // the two column halves of a row tile run on the two CTAs of a cluster (tile = 2 * row tile + cluster rank)
    grid &= ~1;
    if (grid < 2) grid = 2;
    (void)FS2_LAUNCH_CLUSTER(2, tc_conv_gemm_staged_kernel, grid, NUM_THREADS, plan.total(), st, tmA, tmB, tmRes, tmOutF,
                             tmOutB0, tmOutB1, tmVt, b, num_n_blocks, plan);
  } else {
    (void)FS2_LAUNCH(tc_conv_gemm_staged_kernel, grid, NUM_THREADS, plan.total(), st, tmA, tmB, tmRes, tmOutF, tmOutB0,
                     tmOutB1, tmVt, b, num_n_blocks, plan);
  }
  ++g_fs2_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fs2_fail_cuda(e, "tc_conv_gemm_staged_kernel launch");
  return FS2_OK;
}
