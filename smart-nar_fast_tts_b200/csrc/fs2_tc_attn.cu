// fs2_tc_attn.cu -- tcgen05 flash attention for the decoder FFT blocks (sm_100a).
//
// Restates transformer/Modules.py:14-25 with the head split of SubLayers.py:39-55 (H = 2, d_k = 128):
//   out[b, q, h*128:(h+1)*128] = softmax_k( Q_h[b,q,:] . K_h[b,k,:] / sqrt(128), keys >= len_b masked ) @ V_h[b]
// One CTA = one (128-query tile, head, utterance); 112 KB of shared memory and 256 TMEM columns so that TWO CTAs share an
// SM (one's softmax overlaps the other's MMAs and loads).  The [S,S] score matrix never leaves the SM:
//   warp 0   : TMA.  Q tile once, then two independent 2-stage rings: K blocks [64 keys x 128] and V^T blocks
//              [128 d x 64 keys].  A K stage is free as soon as its score MMA has completed, long before the V stage of
//              the same block, which is what lets the score MMAs run two blocks ahead.
//   warp 1   : tcgen05.mma.  S = Q K^T (M=128,N=64,K=128) into one of TWO score buffers, TMEM columns [0,64) / [64,128);
//              O += P V (M=128,N=128,K=64 keys) into TMEM columns [128,256).  S(j+2) is issued right after P(j)V(j), i.e.
//              while the softmax of block j+1 runs, so a finished softmax always finds its next scores waiting
//              (profiles/r1g: with a single score buffer the softmax warps spent 22 % of their time waiting for S).
//   warps 2-5: online softmax, one thread per query row (= TMEM lane): one tcgen05.ld pass brings the row's 64 scores
//              into registers; running max on the raw scores, p = exp2(s * scale - m * scale) as one FFMA + one EX2 per
//              key (the key mask only touches the utterance's last block).  P never touches shared memory: the thread
//              packs its 64 probabilities as bf16x2 into 32 columns of the score buffer it has just read (tcgen05.st) and
//              the PV MMA takes its A operand straight from TMEM -- no staging tile, no proxy fence, and because block j
//              writes P into score buffer j & 1 the softmax of block j+1 never waits for the PV MMA of block j.  O is
//              rescaled in TMEM by exp2(m_old - m_new) only when a row of the warp moved its reference maximum.  The
//              normalised output tile leaves through a swizzled staging tile + TMA store when the whole tile lies
//              inside the utterance.
// V is consumed K-major as V^T ([d, flat row]); the QKV GEMM epilogue (fs2_tc_gemm.cu, EPI_QKV) writes it in that layout.
// Rows follow the ragged layout of fs2_common.cuh: utterance b owns flat rows [off[b], off[b+1]).
// Query rows >= len_b are written as zeros (masked by the caller anyway, Layers.py:43).
#include "fs2_tc_common.cuh"
#include "../../include/fs2_b200.h"

namespace {

using namespace tc;

// Operand planes NP: 1 = bf16 (decoder); 2 = f16x2 (fs2_common.cuh): Q, K, V^T and P are carried as scaled fp16 hi / lo
// planes and every MMA becomes three (hi*lo, lo*hi, hi*hi) -- the fp32-faithful attention of the encoder on the tensor
// cores.  With two planes the tiles double (224 KB) and one CTA owns the SM.
constexpr int BQ = 128, BKV = 64, DK = 128;
constexpr int Q_ATOM = BQ * 128;                 // 128 query rows x 64 elements (128-byte swizzled rows)
constexpr int K_ATOM = BKV * 128;                // 64 key rows x 64 elements
constexpr int V_TILE = DK * 128;                 // [128 d rows x 64 keys]
template <int NP>
struct AttSmem {
  static constexpr int Q_OFF = 0;                               // NP planes x 2 atoms (d 0..63, 64..127)
  static constexpr int K_OFF = Q_OFF + NP * 2 * Q_ATOM;         // 2 stages x NP planes x 2 atoms
  static constexpr int V_OFF = K_OFF + 2 * NP * 2 * K_ATOM;     // 2 stages x NP planes
  static constexpr int BAR_OFF = V_OFF + 2 * NP * V_TILE;       // 96 KB (NP = 1) / 192 KB (NP = 2)
  static constexpr int NUM_BARS = 16;   // 14 used (see the barrier table in the kernel); the TMEM slot follows them
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
};
constexpr float P_SCALE = 2048.0f;               // f16x2: probabilities are split as fp16 terms of p * 2^11
constexpr int ATT_THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;              // S0: columns [0,64); S1: [64,128); O: [128,256); 2 CTAs/SM -> 512

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NP>
__global__ void __launch_bounds__(ATT_THREADS, NP == 1 ? 2 : 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                    const RowLayout lay, bf16* out_b, float scale_log2, float out_scale) {
  using L = AttSmem<NP>;
  constexpr int Q_OFF = L::Q_OFF, K_OFF = L::K_OFF, V_OFF = L::V_OFF, BAR_OFF = L::BAR_OFF, NUM_BARS = L::NUM_BARS;
  constexpr int NCOMBO = NP == 2 ? 3 : 1;      // (A plane, B plane): hi*lo, lo*hi, hi*hi -- smallest first
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bars = base + BAR_OFF;
  // barrier table (8 bytes each): 0 q_full | 1-2 k_full | 3-4 k_empty | 5-6 v_full | 7-8 v_empty | 9-10 s_full |
  // 11-12 p_ready (one per score buffer) | 13-14 o_ready (one per block parity)
  const uint32_t q_full = bars, k_full0 = bars + 8, k_empty0 = bars + 24, v_full0 = bars + 40, v_empty0 = bars + 56,
                 s_full0 = bars + 72, p_ready0 = bars + 88, o_ready0 = bars + 104;
  static_assert(NUM_BARS * 8 > 112, "barrier table overlaps the TMEM slot");
  // mbarrier waits are by phase PARITY: a waiter must be within one phase of the barrier or the test is answered by the
  // wrong phase (it passes early, or blocks on a later phase).  The softmax threads only need "P(j)V(j) complete" when they
  // rescale O and at the very end, i.e. they skip phases -- hence one o_ready barrier per block parity: S(j) is issued
  // after P(j-2)V(j-2) and MMAs complete in issue order, so when a softmax thread works on block j (it has seen S(j)) every
  // P V up to j-2 is complete and barrier (j-1) & 1 is either in the phase of block j-1 or just past it: unambiguous.
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + NUM_BARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, p0 = blockIdx.x * BQ;
  FS2_PDL_PROLOGUE();   // the layout tables read next are written by a predecessor; every CTA waits before it may exit
  const size_t row0 = (size_t)ld_act(lay.off + b);
  const int SA = ld_act(lay.off + b + 1) - (int)row0;   // this utterance's rows (grid + halo)
  const int len = min(ld_act(lay.lens + b), ld_act(lay.ext + b));
  if (p0 >= SA) return;   // uniform over the CTA, before any barrier / TMEM use
  if (base & 1023u) __trap();   // the swizzled tiles assume a 1024-byte aligned dynamic shared memory window

  if (p0 >= len) {  // tile is all padding: zeros, no tensor work (uniform over the CTA)
    for (int idx = threadIdx.x; idx < BQ * (DK / 8); idx += ATT_THREADS) {
      const int q = idx / (DK / 8), c = idx % (DK / 8);
      if (p0 + q < SA)
        for (int pl = 0; pl < NP; ++pl)
          *reinterpret_cast<uint4*>(out_b + ((size_t)pl * lay.R_cap + row0 + p0 + q) * 256 + h * DK + c * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  const int nb = (len + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full0 + 8 * s, 1); mbar_init(k_empty0 + 8 * s, 1);
      mbar_init(v_full0 + 8 * s, 1); mbar_init(v_empty0 + 8 * s, 1);
      mbar_init(s_full0 + 8 * s, 1);
    }
    // Two p_ready barriers, one per score buffer: a softmax warp may run ONE block ahead of the slowest warp (the scores
    // of block j+1 exist before P(j) is complete) but never two (S(j+2) is issued after P(j)V(j)), so arrivals of
    // consecutive blocks must land on different barriers or a fast warp's second arrival would complete the phase of a
    // block whose P is still being written.
    mbar_init(p_ready0, 128);
    mbar_init(p_ready0 + 8, 128);
    mbar_init(o_ready0, 1);
    mbar_init(o_ready0 + 8, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;   // score buffer i at tmem_S + 64 i

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, NP * 2 * Q_ATOM);
      for (int pl = 0; pl < NP; ++pl) {
        tma_load_3d(base + Q_OFF + pl * 2 * Q_ATOM, &tmQ, q_full, h * DK, (int)row0 + p0, pl);
        tma_load_3d(base + Q_OFF + pl * 2 * Q_ATOM + Q_ATOM, &tmQ, q_full, h * DK + 64, (int)row0 + p0, pl);
      }
      for (int j = 0; j < nb; ++j) {
        const int s = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1u;
        const uint32_t ks = base + K_OFF + s * NP * 2 * K_ATOM, vs = base + V_OFF + s * NP * V_TILE;
        mbar_wait(k_empty0 + 8 * s, par);
        mbar_expect_tx(k_full0 + 8 * s, NP * 2 * K_ATOM);
        for (int pl = 0; pl < NP; ++pl) {
          tma_load_3d(ks + pl * 2 * K_ATOM, &tmK, k_full0 + 8 * s, h * DK, (int)row0 + j * BKV, pl);
          tma_load_3d(ks + pl * 2 * K_ATOM + K_ATOM, &tmK, k_full0 + 8 * s, h * DK + 64, (int)row0 + j * BKV, pl);
        }
        mbar_wait(v_empty0 + 8 * s, par);
        mbar_expect_tx(v_full0 + 8 * s, NP * V_TILE);
        for (int pl = 0; pl < NP; ++pl) tma_load_3d(vs + pl * V_TILE, &tmV, v_full0 + 8 * s, (int)row0 + j * BKV, h * DK, pl);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t FMT = NP == 2 ? 0u : 1u;   // fp16 / bf16 operands
      constexpr uint32_t idesc_s = make_idesc_f16kind(BQ, BKV, FMT), idesc_o = make_idesc_f16kind(BQ, DK, FMT);
      constexpr int CA[3] = {0, 1, 0}, CB[3] = {1, 0, 0};
      auto issue_S = [&](int j) {    // scores of block j into score buffer j & 1 (K stage j & 1)
        const int s = j & 1;
        mbar_wait(k_full0 + 8 * s, (j >> 1) & 1);
        fence_after_sync();
#pragma unroll
        for (int c = 0; c < NCOMBO; ++c) {
          const int pa = NP == 2 ? CA[c] : 0, pb = NP == 2 ? CB[c] : 0;
#pragma unroll
          for (int kk = 0; kk < DK / 16; ++kk) {
            const uint64_t ad = make_smem_desc_sw128(base + Q_OFF + pa * 2 * Q_ATOM + (kk >> 2) * Q_ATOM) + (uint64_t)(2 * (kk & 3));
            const uint64_t bd = make_smem_desc_sw128(base + K_OFF + (s * NP + pb) * 2 * K_ATOM + (kk >> 2) * K_ATOM) + (uint64_t)(2 * (kk & 3));
            umma_bf16(tmem_S + (uint32_t)(s * 64), ad, bd, idesc_s, (c | kk) ? 1u : 0u);
          }
        }
        // the K stage is reusable as soon as these MMAs have read it; the producer only waits for that when block j + 2
        // exists (an arrival nobody waits for is flagged by compute-sanitizer's synccheck)
        if (j + 2 < nb) umma_commit(k_empty0 + 8 * s);
        umma_commit(s_full0 + 8 * s);
      };
      mbar_wait(q_full, 0);
      issue_S(0);
      if (nb > 1) issue_S(1);
      for (int j = 0; j < nb; ++j) {
        const int s = j & 1;
        mbar_wait(p_ready0 + 8 * s, (j >> 1) & 1);     // P(j) in TMEM (score buffer j & 1), O rescaled
        mbar_wait(v_full0 + 8 * s, (j >> 1) & 1);
        fence_after_sync();
#pragma unroll
        for (int c = 0; c < NCOMBO; ++c) {
          const int pa = NP == 2 ? CA[c] : 0, pb = NP == 2 ? CB[c] : 0;
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk) {
            // A = P from TMEM: 16 keys = 8 columns of packed 16-bit pairs; plane pa at columns [32 pa, 32 pa + 32)
            const uint32_t at = tmem_S + (uint32_t)(s * 64 + pa * 32 + kk * 8);
            const uint64_t bd = make_smem_desc_sw128(base + V_OFF + (s * NP + pb) * V_TILE) + (uint64_t)(2 * kk);
            umma_bf16_ts(tmem_O, at, bd, idesc_o, (j | c | kk) ? 1u : 0u);
          }
        }
        if (j + 2 < nb) umma_commit(v_empty0 + 8 * s);
        umma_commit(o_ready0 + 8 * s);
        // S(j+2) overwrites score buffer j & 1, which P(j)V(j) -- issued just above -- reads as its A operand.  MMAs of one
        // thread execute in issue order, so the write cannot overtake the read (the same aliasing CUTLASS's sm100 FMHA
        // relies on); tests/test_gpu_properties.py checks run-to-run bit equality at batch 256.  It runs under the softmax
        // of block j + 1.
        if (j + 2 < nb) issue_S(j + 2);
      }
    }
  } else {
    // ===================================================== softmax warps (one thread = one query row)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int sw = row & 7;
    float m = -INFINITY, l = 0.f;
    uint32_t v0[32], v1[32];
    // m: reference maximum of the RAW scores of this row; the softmax runs in the exp2 domain, p = exp2((s - m) * scale)
    for (int j = 0; j < nb; ++j) {
      const uint32_t sbuf = tmem_S + lane_off + (uint32_t)((j & 1) * 64);
      mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
      fence_after_sync();
      // the whole 64-key block of this row in registers: one TMEM pass
      tmem_ld32(sbuf, v0);
      tmem_ld32(sbuf + 32, v1);
      tmem_wait_ld();
      const int kbase = j * BKV;
      if (kbase + BKV > len) {                   // only the utterance's last block holds masked keys (CTA-uniform branch)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (kbase + i >= len) v0[i] = 0xff800000u;        // -inf
          if (kbase + 32 + i >= len) v1[i] = 0xff800000u;
        }
      }
      // block maximum: four independent chains (a single 64-deep fmax / fadd chain is pure latency with 1-2 warps per SMSP)
      float bm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) bm4[i & 3] = fmaxf(bm4[i & 3], fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
      const float bm = fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3]));
      // m is the REFERENCE maximum the running sum and O are expressed against.  It only moves when a block's maximum
      // exceeds it by more than SLACK in the exp2 domain: probabilities may then reach 2^SLACK (exact in fp32, harmless
      // in bf16 / scaled fp16), and the O rescale pass -- a full TMEM round trip -- practically never runs after the first
      // block (with the exact running maximum nearly every block of a short utterance raised some row of the warp).
      constexpr float SLACK = NP == 1 ? 8.0f : 4.0f;   // f16x2: p * 2^11 must stay below the fp16 maximum
      const bool raise = (bm - m) * scale_log2 > SLACK;          // first block: m = -inf -> true; bm is finite (>= 1 valid key)
      const float m_new = raise ? bm : m;
      const float alpha = raise ? fast_exp2((m - m_new) * scale_log2) : 1.0f;   // 0 on the first block
      const float neg_ms = -m_new * scale_log2;
      // probabilities -> packed 16-bit pairs, written over the scores just read: P(j) = columns [0,32) (and [32,64) for the
      // lo plane) of score buffer j & 1; key 2c sits in the low half of column c
      // two keys per instruction where the pipe allows it: one packed FMA (fma.rn.f32x2) scales and shifts a pair of
      // scores, one packed add feeds the pair into the row sum (four independent pair chains); the exponentials stay two
      // MUFU.EX2.  5 instead of 7 instructions per key pair in the loop that bounds this kernel.
      float2 bs2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(neg_ms, neg_ms);
      uint32_t ph[32], pl[NP == 2 ? 32 : 1];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float t0 = __uint_as_float(i < 16 ? v0[2 * i] : v1[2 * i - 32]);
        const float t1 = __uint_as_float(i < 16 ? v0[2 * i + 1] : v1[2 * i - 31]);
        const float2 e2 = __ffma2_rn(make_float2(t0, t1), sc2, nm2);
        const float p0 = fast_exp2(e2.x);   // exp2(-inf) = 0 for masked keys
        const float p1 = fast_exp2(e2.y);
        bs2[i & 3] = __fadd2_rn(bs2[i & 3], make_float2(p0, p1));
        if (NP == 1) {
          ph[i] = pack_bf16x2(p0, p1);
        } else {
          uint16_t h0, l0, h1, l1;
          split2h_scaled(p0 * P_SCALE, h0, l0);
          split2h_scaled(p1 * P_SCALE, h1, l1);
          ph[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
          pl[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
      }
      tmem_st32(sbuf, ph);
      if (NP == 2) tmem_st32(sbuf + 32, reinterpret_cast<uint32_t(&)[32]>(pl));
      l = l * alpha + (((bs2[0].x + bs2[0].y) + (bs2[1].x + bs2[1].y)) + ((bs2[2].x + bs2[2].y) + (bs2[3].x + bs2[3].y)));
      m = m_new;
      // O is stable once P(j-1)V(j-1) has completed.  Every phase of o_ready is waited for (not only the ones a rescale
      // needs): P(j-1)V(j-1) was issued before this block's exponentials started, so the wait is already satisfied, and no
      // phase of the barrier ever completes unobserved (what compute-sanitizer's synccheck reports as "missing wait").
      if (j > 0) mbar_wait(o_ready0 + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
      // rescale the running output only when some row of the warp moved its reference maximum (warp-uniform branch)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        fence_after_sync();
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(tmem_O + lane_off + c * 32, v0);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(__uint_as_float(v0[i]) * alpha);
          tmem_st32(tmem_O + lane_off + c * 32, v0);
        }
      }
      tmem_wait_st();        // P (and a rescaled O) are in TMEM before the MMA warp is released
      fence_before_sync();
      mbar_arrive(p_ready0 + 8 * (j & 1));
    }
    // final: O / l -> bf16
    mbar_wait(o_ready0 + 8 * ((nb - 1) & 1), ((nb - 1) >> 1) & 1);   // the last block's P V
    fence_after_sync();
    const int p = p0 + row;
    const bool valid = p < len;
    const float inv = valid ? out_scale / l : 0.f;   // out_scale undoes the operand pre-scales of the f16x2 mode
    const bool interior = p0 + BQ <= SA;       // the whole 128-row tile belongs to this utterance: TMA store
    if (interior) {
      // stage the [128 x 128] tile (per plane) in the (now idle) Q buffers: two 128B-swizzled [128 x 64] atoms per plane
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem_O + lane_off + c * 32, v0);
        tmem_wait_ld();
        uint8_t* o_row = smem + Q_OFF + (c >> 1) * Q_ATOM + row * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(v0[u * 8 + i]) * inv;
          uint8_t* o16 = o_row + ((((c & 1) * 4 + u) ^ sw) << 4);
          if (NP == 1) {
            *reinterpret_cast<uint4*>(o16) =
                make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
          } else {
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split2h_pair(y[2 * i], y[2 * i + 1], hh[i], ll[i]);
            *reinterpret_cast<uint4*>(o16) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(o16 + 2 * Q_ATOM) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          }
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        for (int pl = 0; pl < NP; ++pl) {
          tma_store_3d(&tmO, base + Q_OFF + pl * 2 * Q_ATOM, h * DK, (int)row0 + p0, pl);
          tma_store_3d(&tmO, base + Q_OFF + pl * 2 * Q_ATOM + Q_ATOM, h * DK + 64, (int)row0 + p0, pl);
        }
        tma_store_commit();
        tma_store_wait_all();
      }
    } else {
      const bool writable = p < SA;              // rows past SA belong to the next utterance
      bf16* o = out_b + (row0 + (writable ? p : 0)) * 256 + h * DK;
      for (int c = 0; c < 4; ++c) {
        __syncwarp();
        tmem_ld32(tmem_O + lane_off + c * 32, v0);
        tmem_wait_ld();
        if (writable) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float y[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) y[u] = __uint_as_float(v0[i + u]) * inv;
            if (NP == 1) {
              *reinterpret_cast<uint4*>(o + c * 32 + i) =
                  make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
            } else {
              uint32_t hh[4], ll[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) split2h_pair(y[2 * t], y[2 * t + 1], hh[t], ll[t]);
              *reinterpret_cast<uint4*>(o + c * 32 + i) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
              *reinterpret_cast<uint4*>(o + (size_t)lay.R_cap * 256 + c * 32 + i) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
          }
        }
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}



template <int NP>
int launch_attention(const bf16* q, const bf16* k, const bf16* vt, const RowLayout& lay, int Rv, int H, bf16* out_b,
                     cudaStream_t st) {
  const uint64_t R = (uint64_t)lay.R_cap;
  CUtensorMap tmQ, tmK, tmV, tmO;
  // [planes][rows][cols] maps; the V^T planes are [256 d][Rv] each
  if (!tc::make_tmap_bf16_3d(&tmQ, q, NP, R, 256, 256, R * 256, BQ) || !tc::make_tmap_bf16_3d(&tmK, k, NP, R, 256, 256, R * 256, BKV) ||
      !tc::make_tmap_bf16_3d(&tmV, vt, NP, 256, (uint64_t)Rv, (uint64_t)Rv, (uint64_t)256 * Rv, DK) ||
      !tc::make_tmap_bf16_3d(&tmO, out_b, NP, R, 256, 256, R * 256, BQ))
    return fs2_fail_cuda(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(attention)");
  static std::atomic<bool> configured{false};   // handles on several host threads may race here: benign, but formally atomic
  const int smem = AttSmem<NP>::TOTAL;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_attention_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fs2_fail_cuda(e, "cudaFuncSetAttribute(tc_attention)");
    configured = true;
  }
  dim3 grid((FS2_ROWS_PER_UTT(lay.S, FS2_HALO) + BQ - 1) / BQ, H, lay.B);
  // softmax(QK^T / sqrt(dk)) in the exp2 domain; f16x2 operands carry 2^4 (Q, K, V) and 2^11 (P) pre-scales
  const double a = FS2_F16X2_ACT_SCALE;
  const float scale_log2 = (float)(1.4426950408889634 / sqrt((double)DK) / (NP == 2 ? a * a : 1.0));
  const float out_scale = NP == 2 ? (float)(1.0 / ((double)P_SCALE * a)) : 1.0f;
  (void)FS2_LAUNCH((tc_attention_kernel<NP>), grid, ATT_THREADS, smem, st, tmQ, tmK, tmV, tmO, lay, out_b, scale_log2, out_scale);
  ++g_fs2_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fs2_fail_cuda(e, "tc_attention_kernel launch");
  return FS2_OK;
}

}  // namespace

// planes: 1 = bf16 operands / bf16 output; 2 = f16x2 operand planes in, f16x2 operand planes out (fs2_common.cuh)
int tc_attention_launch(const bf16* q, const bf16* k, const bf16* vt, const RowLayout& lay, int Rv, int H, int planes,
                        bf16* out_b, cudaStream_t st) {
  if (lay.B <= 0 || lay.S <= 0) return FS2_OK;
  if (H != 2) return fs2_fail_cuda(cudaErrorInvalidValue, "tc_attention: built for H = 2, d_k = 128");
  if (!lay.off || !lay.ext || !lay.lens) return fs2_fail_cuda(cudaErrorInvalidValue, "tc_attention: row layout missing");
  if (planes == 2) return launch_attention<2>(q, k, vt, lay, Rv, H, out_b, st);
  if (planes == 1) return launch_attention<1>(q, k, vt, lay, Rv, H, out_b, st);
  return fs2_fail_cuda(cudaErrorInvalidValue, "tc_attention: planes must be 1 or 2");
}
