// fs2_api.cu -- C ABI (include/fs2_b200.h): handle, weight repacking, workspace and the
// two-stage orchestration of FastSpeech2Align.forward (inference branch,
// reference model/fastspeech2_align.py:30-100).  No torch types, no CPU compute path:
// every function needs a CUDA device and fails with FS2_ERR_NO_DEVICE / FS2_ERR_CUDA otherwise.
#include "fs2_common.cuh"
#include "../../include/fs2_b200.h"

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

static thread_local std::string g_last_error;  // creation-time / handle-less errors

int fs2_fail_cuda(cudaError_t e, const char* what) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  g_last_error = buf;
  return FS2_ERR_CUDA;
}

namespace {

struct RawT {
  float* d = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

struct GemmW {
  float* wf = nullptr;  // [taps][K][N]
  bf16* wb = nullptr;   // [3][taps][N][K]: plane 0 = bf16(w), planes 1-2 = split residuals (bf16x3 mode)
  bf16* wh = nullptr;   // [2][taps][N][K]: fp16 hi / lo terms of w * w_scale (f16x2 mode)
  float w_scale = 1.f;  // power of two
  float* bias = nullptr;
  int N = 0, K = 0, taps = 1;
};
struct FftW {
  GemmW qkv, fc, w1, w2;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};
struct CrossW {   // FFTBlock2 (Layers.py:51-70): cross-attention (queries from one sequence, keys / values from another) + FFN
  GemmW q, kv, fc, w1, w2;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};
struct PredW {
  GemmW c1, c2;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr, *lin_w = nullptr;
  float lin_b = 0.f;
};

}  // namespace

// Raw + packed weights of one model: read-only once fs2_load_weights has returned, so any number of handles on the
// same device (one per CUDA stream: StreamedSynthesizer) can run on ONE copy (fs2_share_weights).  Freed with the last
// handle that references it.
struct fs2_weights {
  int device = 0;
  std::map<std::string, RawT> raw;
  std::vector<void*> owned;  // packed weight allocations
  std::vector<FftW> enc, dec;
  PredW pred[3];
  GemmW mel_linear;
  std::vector<GemmW> postnet;
  // training-side aligner (mel_encoder.*): raw tensors are kept at load time, packed on the first fs2_op_mel_encoder call
  std::mutex menc_mu;
  bool menc_built = false;
  GemmW menc_pre1, menc_pre2;
  std::vector<CrossW> menc;
  ~fs2_weights() {
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    cudaDeviceSynchronize();   // kernels of another handle that still read these buffers
    for (auto& kv : raw) cudaFree(kv.second.d);
    for (void* p : owned) cudaFree(p);
    if (cur >= 0) cudaSetDevice(cur);
  }
};

struct fs2_handle {
  fs2_dims dims{};
  int device = 0;
  int prec_enc = FS2_PREC_BF16X3, prec_dec = FS2_PREC_BF16;
  int halo_keep = 2;  // padded rows kept per utterance (2 = packed; >= max length = the reference's padded grid)
  int mel_post_cm = 0;  // stage 2 / fs2_op_mel_postnet write mel_post channel-major [B, n_mel, T] (vocoder hand-off)
  int upsampler = 0;    // FS2_UPSAMPLER_HARD (LengthRegulator, the reference's wiring) / FS2_UPSAMPLER_GAUSSIAN
  bool loaded = false;
  std::string err;
  std::shared_ptr<fs2_weights> w;   // possibly shared with other handles (fs2_share_weights)
  float* pe_ext[2] = {nullptr, nullptr};  // on-the-fly tables for S > max_seq_len (encoder / decoder)
  int pe_ext_n[2] = {0, 0};
  std::map<std::string, std::pair<void*, size_t>> ws;  // growable workspace
  int* host_tmax = nullptr;                            // pinned
  cudaStream_t fill_stream = nullptr;                  // zero-fill of fresh workspace (see ensure)
  // tracing: per-kernel-class device time from CUDA events on the launching stream (fs2_profile_*)
  int prof_on = 0;   // 0 off, 1 every kernel class bracketed by events, 2 only whole segments (encoder / decoder stack, ...)
  struct ProfPending { int slot; cudaEvent_t a, b; };
  std::vector<ProfPending> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  std::vector<std::string> prof_names;
  std::vector<double> prof_ms;
  std::vector<long long> prof_launches;
  int prof_slot(const std::string& name) {
    for (size_t i = 0; i < prof_names.size(); ++i) if (prof_names[i] == name) return (int)i;
    prof_names.push_back(name); prof_ms.push_back(0.0); prof_launches.push_back(0);
    return (int)prof_names.size() - 1;
  }
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  // CUDA-graph cache (fs2_forward_stage*_graph): one instantiated graph per (stage, B, bucket, pointers, controls).  A graph
  // bakes in every workspace / weight / table pointer and every setting its launches read, so `gen` is bumped whenever one
  // of those may have changed (workspace growth, position-table growth, weights, precision, packing, layouts): an entry
  // captured under an older generation is re-captured at its next use.
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    unsigned long long gen = 0;
    int seen = 0;            // 0 = never run, 1 = ran eagerly once (workspace sized), -1 = capture failed: always eager
    long long kernels = 0;   // kernel nodes (for fs2_launch_count)
    // host-side state stage 1 leaves for stage 2: a replay launches no host code, so it is restored from here
    int st_B = 0, st_L = 0; float* st_enc_out = nullptr; RowLayout st_lay1{}; const int* st_Ldev = nullptr;
  };
  std::map<std::string, GraphEntry> graphs;
  unsigned long long gen = 1;
  long long graph_replays = 0, graph_captures = 0;
  cudaStream_t capture_stream = nullptr;   // private stream the stages are captured on (the caller's may be the legacy
                                           // default stream, which cannot capture); graphs are LAUNCHED on the caller's
  int* shape_host = nullptr;   // pinned int[4]: {L, T} of the forward in flight (copied to shape_dev before a graph launch)
  int* shape_dev = nullptr;    // device int[4] the replayed kernels read the true L / T from
  const int* cur_Ldev = nullptr;  // non-null while a *_graph entry point enqueues: true L / T live in shape_dev
  const int* cur_Tdev = nullptr;
  const int* st_Ldev = nullptr;   // stage 1 ran with L as an upper bound: [B, L] tensors (cum, ...) have the TRUE row stride
  void drop_graphs() {
    for (auto& kv : graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
    ++gen;
  }
  // stage-1 -> stage-2 state
  bool have_stage1 = false;
  int st_B = 0, st_L = 0, st_Tmax = 0;
  int st_frames = 0;            // frames of the whole batch (sum of mel_lens): sizes the tile-shape decisions of stage 2
  float* st_enc_out = nullptr;  // encoder output, rows in the stage-1 layout
  RowLayout st_lay1{};          // stage-1 (phoneme) row layout; its buffers live in the workspace

  int fail(int code, const std::string& m) {
    err = m;
    g_last_error = m;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    int rc = fs2_fail_cuda(e, what);
    err = g_last_error;
    return rc;
  }
  void* ensure(const std::string& name, size_t bytes) {
    auto it = ws.find(name);
    if (it != ws.end() && it->second.second >= bytes) return it->second.first;
    if (it != ws.end()) {
      cudaFree(it->second.first);
      ws.erase(it);
    }
    ++gen;   // a buffer moved (or appeared): graphs captured before this do not know it
    size_t cap = bytes + bytes / 4 + 256;
    void* p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
    // fresh workspace is zeroed once so that no kernel can ever multiply a masked 0 with a NaN bit pattern.  The fill
    // must have FINISHED before this returns: the caller's kernels run on its own stream, which (torch side streams are
    // cudaStreamNonBlocking) is not ordered against the legacy default stream a plain cudaMemset would use -- a delayed
    // fill would wipe what the first kernels of the forward wrote (seen as T = 0 on a fresh engine of a
    // StreamedSynthesizer while other threads kept stream 0 busy).  Allocation is rare (first use / growth).
    // A private non-blocking stream: the fill waits for nobody else's work either.
    if (!fill_stream && cudaStreamCreateWithFlags(&fill_stream, cudaStreamNonBlocking) != cudaSuccess) {
      fill_stream = nullptr; cudaFree(p); return nullptr;
    }
    if (cudaMemsetAsync(p, 0, cap, fill_stream) != cudaSuccess || cudaStreamSynchronize(fill_stream) != cudaSuccess) {
      cudaFree(p); return nullptr;
    }
    ws[name] = {p, cap};
    return p;
  }
};

#define HCHECK(expr)                                         \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return h->cuda_fail(_e, #expr);   \
  } while (0)
#define RCHECK(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != FS2_OK) return _rc; \
  } while (0)
#define WS(T, var, name, count)                                                   \
  T* var = reinterpret_cast<T*>(h->ensure(name, sizeof(T) * (size_t)(count)));    \
  if (!var) return h->fail(FS2_ERR_CUDA, std::string("workspace allocation failed: ") + name)

namespace {

// Times everything enqueued on `st` during its lifetime under `name` when tracing is enabled (otherwise free).
struct ProfScope {
  fs2_handle* h; cudaStream_t st; int slot = -1; cudaEvent_t a = nullptr; long long l0 = 0;
  ProfScope(fs2_handle* h_, const std::string& name, cudaStream_t st_, int mode = 1) : h(h_), st(st_) {
    if (!h || h->prof_on != mode) return;
    slot = h->prof_slot(name);
    a = h->prof_event();
    l0 = g_fs2_launches;
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (slot < 0) return;
    cudaEvent_t b = h->prof_event();
    cudaEventRecord(b, st);
    h->prof_pending.push_back({slot, a, b});
    h->prof_launches[slot] += g_fs2_launches - l0;
  }
};
#define PROF(name) ProfScope _prof_scope(h, name, st)
#define PROF_SEG(name) ProfScope _prof_seg(h, name, st, 2)   /* segment-level tracing (mode 2): no events between the kernels inside */

const float* raw_ptr(fs2_handle* h, const std::string& k) {
  auto it = h->w->raw.find(k);
  return it == h->w->raw.end() ? nullptr : it->second.d;
}

int need(fs2_handle* h, const std::string& k, std::initializer_list<int64_t> shape, const float** out) {
  auto it = h->w->raw.find(k);
  if (it == h->w->raw.end()) return h->fail(FS2_ERR_MISSING_WEIGHT, "missing state_dict key: " + k);
  std::vector<int64_t> want(shape);
  if (it->second.shape != want) {
    std::string m = "bad shape for " + k + ": got [";
    for (auto v : it->second.shape) m += std::to_string(v) + ",";
    m += "] want [";
    for (auto v : want) m += std::to_string(v) + ",";
    return h->fail(FS2_ERR_INVALID, m + "]");
  }
  *out = it->second.d;
  return FS2_OK;
}

template <typename T>
int dev_alloc(fs2_handle* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, sizeof(T) * (count ? count : 1));
  if (e != cudaSuccess) return h->cuda_fail(e, "cudaMalloc(packed weight)");
  h->w->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return FS2_OK;
}

// Build one GEMM weight from `parts` torch tensors [N_i, K, taps] stacked along N (QKV concat).
int build_gemm(fs2_handle* h, GemmW& g, const std::vector<std::string>& wkeys, const std::vector<std::string>& bkeys,
               int N_each, int K, int taps, const float* scale, const float* bias_override, cudaStream_t st) {
  const int parts = (int)wkeys.size();
  g.N = N_each * parts;
  g.K = K;
  g.taps = taps;
  RCHECK(dev_alloc(h, &g.wf, (size_t)taps * K * g.N));
  RCHECK(dev_alloc(h, &g.wb, (size_t)3 * taps * K * g.N));
  RCHECK(dev_alloc(h, &g.bias, (size_t)g.N));
  for (int i = 0; i < parts; ++i) {
    const float* w = nullptr;
    if (taps == 1 && h->w->raw.count(wkeys[i]) && h->w->raw[wkeys[i]].shape.size() == 2)
      RCHECK(need(h, wkeys[i], {N_each, K}, &w));
    else
      RCHECK(need(h, wkeys[i], {N_each, K, taps}, &w));
    HCHECK(rowops_pack_weight(w, N_each, K, taps, scale, g.wf, g.wb, g.N, i * N_each, st));
    if (bias_override) {
      HCHECK(cudaMemcpyAsync(g.bias + i * N_each, bias_override, sizeof(float) * N_each, cudaMemcpyDeviceToDevice, st));
    } else {
      const float* b = nullptr;
      RCHECK(need(h, bkeys[i], {N_each}, &b));
      HCHECK(cudaMemcpyAsync(g.bias + i * N_each, b, sizeof(float) * N_each, cudaMemcpyDeviceToDevice, st));
    }
  }
  RCHECK(dev_alloc(h, &g.wh, (size_t)2 * taps * K * g.N));
  HCHECK(rowops_pack_weight_f16x2(g.wf, g.N, K, taps, g.wh, &g.w_scale, st));
  return FS2_OK;
}

int build_fft(fs2_handle* h, FftW& L, const std::string& p, cudaStream_t st) {
  const int D = h->dims.d_model, F = h->dims.d_ffn;
  const std::string a = p + ".slf_attn.", f = p + ".pos_ffn.";
  RCHECK(build_gemm(h, L.qkv, {a + "w_qs.weight", a + "w_ks.weight", a + "w_vs.weight"},
                    {a + "w_qs.bias", a + "w_ks.bias", a + "w_vs.bias"}, D, D, 1, nullptr, nullptr, st));
  RCHECK(build_gemm(h, L.fc, {a + "fc.weight"}, {a + "fc.bias"}, D, D, 1, nullptr, nullptr, st));
  RCHECK(build_gemm(h, L.w1, {f + "w_1.weight"}, {f + "w_1.bias"}, F, D, h->dims.ffn_k1, nullptr, nullptr, st));
  RCHECK(build_gemm(h, L.w2, {f + "w_2.weight"}, {f + "w_2.bias"}, D, F, h->dims.ffn_k2, nullptr, nullptr, st));
  const float* t = nullptr;
  RCHECK(need(h, a + "layer_norm.weight", {D}, &t)); L.ln1_g = const_cast<float*>(t);
  RCHECK(need(h, a + "layer_norm.bias", {D}, &t));   L.ln1_b = const_cast<float*>(t);
  RCHECK(need(h, f + "layer_norm.weight", {D}, &t)); L.ln2_g = const_cast<float*>(t);
  RCHECK(need(h, f + "layer_norm.bias", {D}, &t));   L.ln2_b = const_cast<float*>(t);
  return FS2_OK;
}

int build_pred(fs2_handle* h, PredW& P, const std::string& p, cudaStream_t st) {
  const int D = h->dims.d_model, C = h->dims.vp_filter, k = h->dims.vp_kernel;
  const std::string c = p + ".conv_layer.";
  RCHECK(build_gemm(h, P.c1, {c + "conv1d_1.conv.weight"}, {c + "conv1d_1.conv.bias"}, C, D, k, nullptr, nullptr, st));
  RCHECK(build_gemm(h, P.c2, {c + "conv1d_2.conv.weight"}, {c + "conv1d_2.conv.bias"}, C, C, k, nullptr, nullptr, st));
  const float* t = nullptr;
  RCHECK(need(h, c + "layer_norm_1.weight", {C}, &t)); P.ln1_g = const_cast<float*>(t);
  RCHECK(need(h, c + "layer_norm_1.bias", {C}, &t));   P.ln1_b = const_cast<float*>(t);
  RCHECK(need(h, c + "layer_norm_2.weight", {C}, &t)); P.ln2_g = const_cast<float*>(t);
  RCHECK(need(h, c + "layer_norm_2.bias", {C}, &t));   P.ln2_b = const_cast<float*>(t);
  RCHECK(need(h, p + ".linear_layer.weight", {1, C}, &t)); P.lin_w = const_cast<float*>(t);
  RCHECK(need(h, p + ".linear_layer.bias", {1}, &t));
  HCHECK(cudaMemcpyAsync(&P.lin_b, t, sizeof(float), cudaMemcpyDeviceToHost, st));
  HCHECK(cudaStreamSynchronize(st));
  return FS2_OK;
}

// transformer/Models.py:10-30 in float64 on the host (libm), cast to fp32, uploaded.
int sinusoid_table_host(fs2_handle* h, int n_pos, int D, float* dev_out, cudaStream_t st) {
  std::vector<float> tab((size_t)n_pos * D);
  std::vector<double> denom(D);
  for (int j = 0; j < D; ++j) denom[j] = pow(10000.0, 2.0 * (double)(j / 2) / (double)D);
  for (int p = 0; p < n_pos; ++p)
    for (int j = 0; j < D; ++j) {
      const double ang = (double)p / denom[j];
      tab[(size_t)p * D + j] = (float)((j & 1) ? cos(ang) : sin(ang));
    }
  HCHECK(cudaMemcpyAsync(dev_out, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice, st));
  HCHECK(cudaStreamSynchronize(st));  // `tab` is pageable and dies at return
  return FS2_OK;
}

// position table for a stack (0 = encoder, 1 = decoder) covering S positions (Models.py:82-91, 218-233)
int position_table(fs2_handle* h, int stack, int S, const float** out, cudaStream_t st) {
  const int D = h->dims.d_model;
  if (S <= h->dims.max_seq_len) {
    *out = raw_ptr(h, stack == 0 ? "txt_encoder.position_enc" : "mel_decoder.position_enc");
    return FS2_OK;
  }
  if (h->pe_ext_n[stack] < S) {
    if (h->pe_ext[stack]) cudaFree(h->pe_ext[stack]);
    h->pe_ext[stack] = nullptr;
    ++h->gen;
    const int n = S + S / 4;
    HCHECK(cudaMalloc(reinterpret_cast<void**>(&h->pe_ext[stack]), sizeof(float) * (size_t)n * D));
    RCHECK(sinusoid_table_host(h, n, D, h->pe_ext[stack], st));
    h->pe_ext_n[stack] = n;
  }
  *out = h->pe_ext[stack];
  return FS2_OK;
}

inline int planes_of(int prec) {
  return prec == FS2_PREC_BF16X3 ? 3 : prec == FS2_PREC_F16X2 ? 2 : prec == FS2_PREC_BF16 ? 1 : 0;
}
inline bool is_split(int prec) { return prec == FS2_PREC_BF16X3 || prec == FS2_PREC_F16X2; }

// bf16 shadow of an fp32 activation tensor for a tensor-core consumer of precision `prec` (no-op for fp32 consumers)
cudaError_t make_shadow(const float* x, size_t n, int prec, bf16* xb, cudaStream_t st) {
  if (prec == FS2_PREC_BF16) return rowops_f32_to_bf16(x, (int64_t)n, xb, st);
  if (is_split(prec)) return rowops_split(x, (int64_t)n, planes_of(prec), xb, (int64_t)n, st);
  return cudaSuccess;
}

// Device-resident ragged row layout (fs2_common.cuh) in workspace buffers named `name`.* ; lens32 = null -> every
// utterance has S grid rows.  halo_keep: padded rows kept after the valid ones (>= S: the reference's padded grid).
// extra_ext > 0: one pseudo utterance with extra_ext grid rows is appended (index B; the layout then has B + 1
// utterances and carries no lens pointer: it is used with MASK_GRID only).
int make_layout(fs2_handle* h, const std::string& name, const int* lens32, int B, int S, int halo_keep, int halo_rows,
                RowLayout* out, cudaStream_t st, int extra_ext = 0, const int* S_dev = nullptr) {
  const int Bt = B + (extra_ext > 0 ? 1 : 0);
  if (Bt > 65535) return h->fail(FS2_ERR_UNSUPPORTED, "more than 65535 utterances in one call");
  if (S > FS2_MAX_ROWS_PER_UTT) return h->fail(FS2_ERR_UNSUPPORTED, "more than 65535 rows per utterance");
  // host-known upper bound of the rows in use: a real utterance never needs more than lens + halo_keep <= S grid rows
  const int R_cap = B * FS2_ROWS_PER_UTT(S, halo_rows) + (extra_ext > 0 ? FS2_ROWS_PER_UTT(extra_ext, halo_rows) : 0);
  WS(int, off, name + ".off", (size_t)Bt + 1);
  WS(int, ext, name + ".ext", (size_t)Bt);
  WS(unsigned, rowmap, name + ".map", (size_t)R_cap);
  HCHECK(rowops_build_layout(lens32, B, S, halo_keep, halo_rows, off, ext, rowmap, R_cap, st, extra_ext, S_dev));
  out->B = Bt; out->S = S; out->R_cap = R_cap; out->off = off; out->ext = ext; out->lens = extra_ext > 0 ? nullptr : lens32;
  out->rowmap = rowmap;
  out->rows_hint = 0;
  out->S_dev = S_dev;
  return FS2_OK;
}

// handle-less variant for the stand-alone operator entry points
struct TmpLayout {
  RowLayout lay{};
  int *off = nullptr, *ext = nullptr;
  unsigned* map = nullptr;
  cudaError_t build(const int* lens32, int B, int S, int halo_keep, int halo_rows, cudaStream_t st) {
    const int R_cap = B * FS2_ROWS_PER_UTT(S, halo_rows);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&off), sizeof(int) * ((size_t)B + 1));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ext), sizeof(int) * (size_t)B);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&map), sizeof(unsigned) * (size_t)(R_cap > 0 ? R_cap : 1));
    if (e == cudaSuccess) e = rowops_build_layout(lens32, B, S, halo_keep, halo_rows, off, ext, map, R_cap, st);
    lay = RowLayout{B, S, R_cap, off, ext, lens32, map, 0, nullptr};
    return e;
  }
  ~TmpLayout() { cudaFree(off); cudaFree(ext); cudaFree(map); }
};

ConvGemmArgs base_args(const GemmW& w, const RowLayout& lay) {
  ConvGemmArgs a;
  memset(&a, 0, sizeof a);
  a.K = w.K; a.N = w.N; a.taps = w.taps;
  a.Wf = w.wf; a.Wb = w.wb; a.Wh = w.wh; a.bias = w.bias;
  a.acc_scale = 1.f;
  a.acc_scale_f16x2 = 1.f / (FS2_F16X2_ACT_SCALE * w.w_scale);
  a.lay = lay;
  return a;
}

int run_gemm(fs2_handle* h, int prec, const ConvGemmArgs& a, cudaStream_t st, const std::string& tag = std::string()) {
  ProfScope _ps(tag.empty() ? nullptr : h, tag, st);
  if (prec == FS2_PREC_FP32) {
    cudaError_t e = simt_conv_gemm_launch(a, st);
    if (e != cudaSuccess) return h ? h->cuda_fail(e, "simt_conv_gemm_launch") : fs2_fail_cuda(e, "simt_conv_gemm_launch");
    return FS2_OK;
  }
  ConvGemmArgs b = a;
  b.planes = planes_of(prec);
  b.acc_scale = 1.f;
  if (prec == FS2_PREC_F16X2) { b.Wb = a.Wh; b.acc_scale = a.acc_scale_f16x2; }
  if (b.out_planes == 0) b.out_planes = b.planes;
  int rc = tc_conv_gemm_launch(b, st);
  if (rc != FS2_OK && h) h->err = g_last_error;
  return rc;
}

// Layers.py:39-48 x n layers on ragged-grid activations.  x (fp32) and, in tensor-core modes, xb (its bf16 / bf16x3
// shadow) are updated in place.  Returns through x/xb.
int run_fft_stack(fs2_handle* h, std::vector<FftW>& Ls, int l0, int l1, int prec, float* x, bf16* xb, const RowLayout& lay,
                  cudaStream_t st) {
  const int D = h->dims.d_model, F = h->dims.d_ffn, H = h->dims.n_heads, dk = D / H;
  const size_t R = (size_t)lay.R_cap;
  const std::string tg = (&Ls == &h->w->enc) ? "enc." : "dec.";
  PROF_SEG(tg + "fft_stack");
  WS(float, y, "fft.y", R * D);
  if (prec == FS2_PREC_FP32) {
    WS(float, qkv, "fft.qkv", R * 3 * D);
    WS(float, att, "fft.att", R * D);
    WS(float, hid, "fft.hid", R * F);
    for (int l = l0; l < l1; ++l) {
      FftW& L = Ls[l];
      ConvGemmArgs a = base_args(L.qkv, lay);
      a.A = x; a.epi = EPI_BIAS; a.mask_mode = MASK_GRID; a.out = qkv; a.ldo = 3 * D;
      RCHECK(run_gemm(h, prec, a, st, tg + "qkv"));
      {
        PROF(tg + "attn");
        HCHECK(simt_attention_launch(qkv, 3 * D, 0, D, 2 * D, lay, H, dk, att, D, st));
      }
      a = base_args(L.fc, lay);
      a.A = att; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = x; a.ln_g = L.ln1_g; a.ln_b = L.ln1_b;
      a.out = y; a.ldo = D;
      RCHECK(run_gemm(h, prec, a, st, tg + "fc_ln"));
      a = base_args(L.w1, lay);
      a.A = y; a.epi = EPI_RELU; a.mask_mode = MASK_GRID; a.out = hid; a.ldo = F;
      RCHECK(run_gemm(h, prec, a, st, tg + "ffn_w1"));
      a = base_args(L.w2, lay);
      a.A = hid; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = y; a.ln_g = L.ln2_g; a.ln_b = L.ln2_b;
      a.out = x; a.ldo = D;
      RCHECK(run_gemm(h, prec, a, st, tg + "ffn_w2_ln"));
    }
    return FS2_OK;
  }
  if (prec == FS2_PREC_BF16X3) {
    // fp32-faithful tensor-core path with 3-term bf16 operands: split-operand GEMMs, fp32 FFMA attention
    const size_t np = (size_t)planes_of(prec);
    WS(float, qkv, "fft.qkv", R * 3 * D);
    WS(float, att, "fft.att", R * D);
    WS(bf16, yb3, "fft.yb3", np * R * D);
    WS(bf16, attb3, "fft.attb3", np * R * D);
    WS(bf16, hidb3, "fft.hidb3", np * R * F);
    for (int l = l0; l < l1; ++l) {
      FftW& L = Ls[l];
      ConvGemmArgs a = base_args(L.qkv, lay);
      a.Ab = xb; a.epi = EPI_BIAS; a.mask_mode = MASK_GRID; a.out = qkv; a.ldo = 3 * D;
      RCHECK(run_gemm(h, prec, a, st, tg + "qkv"));
      {
        PROF(tg + "attn");
        HCHECK(simt_attention_launch(qkv, 3 * D, 0, D, 2 * D, lay, H, dk, att, D, st));
        HCHECK(rowops_split(att, (int64_t)(R * D), (int)np, attb3, (int64_t)(R * D), st));
      }
      a = base_args(L.fc, lay);
      a.Ab = attb3; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = x; a.ln_g = L.ln1_g; a.ln_b = L.ln1_b;
      a.out = y; a.ldo = D; a.out_b = yb3; a.ldob = D;
      RCHECK(run_gemm(h, prec, a, st, tg + "fc_ln"));
      a = base_args(L.w1, lay);
      a.Ab = yb3; a.epi = EPI_RELU; a.mask_mode = MASK_GRID; a.out_b = hidb3; a.ldob = F;
      RCHECK(run_gemm(h, prec, a, st, tg + "ffn_w1"));
      a = base_args(L.w2, lay);
      a.Ab = hidb3; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = y; a.ln_g = L.ln2_g; a.ln_b = L.ln2_b;
      a.out = x; a.ldo = D; a.out_b = xb; a.ldob = D;
      RCHECK(run_gemm(h, prec, a, st, tg + "ffn_w2_ln"));
    }
    return FS2_OK;
  }
  // tcgen05 path, attention included: bf16 (1 operand plane) or f16x2 (2 scaled fp16 planes, fp32-faithful)
  const int Rv = (lay.R_cap + 7) & ~7;
  const size_t np = (size_t)planes_of(prec);
  WS(bf16, yb, "fft.yb", np * R * D);
  WS(bf16, qb, "fft.qb", np * R * D);
  WS(bf16, kb, "fft.kb", np * R * D);
  WS(bf16, vtb, "fft.vtb", np * (size_t)D * Rv);
  WS(bf16, attb, "fft.attb", np * R * D);
  WS(bf16, hidb, "fft.hidb", np * R * F);
  for (int l = l0; l < l1; ++l) {
    FftW& L = Ls[l];
    ConvGemmArgs a = base_args(L.qkv, lay);
    a.Ab = xb; a.epi = EPI_QKV; a.mask_mode = MASK_GRID; a.q_b = qb; a.k_b = kb; a.vt_b = vtb; a.Rv = Rv;
    RCHECK(run_gemm(h, prec, a, st, tg + "qkv"));
    {
      PROF(tg + "attn");
      int rc = tc_attention_launch(qb, kb, vtb, lay, Rv, H, (int)np, attb, st);
      if (rc != FS2_OK) { h->err = g_last_error; return rc; }
    }
    a = base_args(L.fc, lay);
    a.Ab = attb; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = x; a.ln_g = L.ln1_g; a.ln_b = L.ln1_b;
    a.out = y; a.ldo = D; a.out_b = yb; a.ldob = D;
    RCHECK(run_gemm(h, prec, a, st, tg + "fc_ln"));
    a = base_args(L.w1, lay);
    a.Ab = yb; a.epi = EPI_RELU; a.mask_mode = MASK_GRID; a.out_b = hidb; a.ldob = F;
    RCHECK(run_gemm(h, prec, a, st, tg + "ffn_w1"));
    a = base_args(L.w2, lay);
    a.Ab = hidb; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = y; a.ln_g = L.ln2_g; a.ln_b = L.ln2_b;
    a.out = x; a.ldo = D; a.out_b = xb; a.ldob = D;
    RCHECK(run_gemm(h, prec, a, st, tg + "ffn_w2_ln"));
  }
  return FS2_OK;
}

// modules.py:278-286 with the reference's padded-grid semantics (halo leak, SURVEY.md section 8(a) note 1): the layout
// keeps the 2 padded rows that can reach a valid output.  out_user[B,S]; rows the layout does not carry are 0 (masked).
int run_predictor(fs2_handle* h, PredW& P, int prec, const float* x, const bf16* xb, const RowLayout& lay,
                  float* out_user, cudaStream_t st, bool out_cleared = false) {
  const int C = h->dims.vp_filter;
  const std::string tg = &P == &h->w->pred[0] ? "dur." : &P == &h->w->pred[1] ? "pitch." : "energy.";
  const size_t R = (size_t)lay.R_cap;
  WS(float, p1, "pred.h1", R * C);
  bf16* p1b = nullptr;
  if (prec != FS2_PREC_FP32) {
    WS(bf16, t, "pred.h1b", (size_t)planes_of(prec) * R * C);
    p1b = t;
  }
  if (!out_cleared) HCHECK(rowops_fill_zero(out_user, sizeof(float) * (size_t)lay.B * lay.S, st));
  ConvGemmArgs a = base_args(P.c1, lay);
  a.A = x; a.Ab = xb; a.epi = EPI_RELU_LN; a.mask_mode = MASK_GRID; a.ln_g = P.ln1_g; a.ln_b = P.ln1_b;
  a.out = prec == FS2_PREC_FP32 ? p1 : nullptr; a.ldo = C; a.out_b = p1b; a.ldob = C;
  RCHECK(run_gemm(h, prec, a, st, tg + "conv1"));
  a = base_args(P.c2, lay);
  a.A = p1; a.Ab = p1b; a.epi = EPI_RELU_LN_DOT; a.mask_mode = MASK_LEN; a.ln_g = P.ln2_g; a.ln_b = P.ln2_b;
  a.dot_w = P.lin_w; a.dot_b = P.lin_b; a.out_user = out_user; a.ldu = 1;
  RCHECK(run_gemm(h, prec, a, st, tg + "conv2_dot"));
  return FS2_OK;
}

// fastspeech2_align.py:83-85: mel_linear, PostNet (BatchNorm folded), residual.  dec follows `lay` (possibly packed).
// The reference runs PostNet over the whole padded grid [B, T] and returns its padded rows, which are not masked
// (SURVEY.md a15).  A row farther than H = layers * (k-1)/2 from the last valid frame sees only mel_linear's bias row (the
// decoder output there is exactly zero) and the zero padding past T: its value depends only on its distance to T.  So
// PostNet runs on a packed grid -- utterance b keeps min(len_b + 2H, T) rows, of which the first len_b + H come out
// exact -- plus ONE pseudo utterance of min(T, 2H+1) all-bias rows whose outputs are copied into every farther row.
int run_mel_postnet(fs2_handle* h, int prec, const float* dec, const bf16* decb, const RowLayout& lay, float* mel,
                    float* mel_post, cudaStream_t st) {
  const int M = h->dims.n_mel, P = h->dims.pn_dim, NL = h->dims.pn_layers;
  const int B = lay.B, T = lay.S;
  const int H = NL * ((h->dims.pn_kernel - 1) / 2);
  const bool tc = prec != FS2_PREC_FP32;
  const size_t np = (size_t)planes_of(prec);
  PROF_SEG("mel_postnet");
  RowLayout pn;
  // never fewer rows than the source layout carries (mel_linear scatters every source grid row into this grid)
  const int keep = !lay.lens ? T : (h->halo_keep > 2 * H ? h->halo_keep : 2 * H);
  // the pseudo utterance has min(T, 2H + 1) rows: the layout kernel clamps to the true T itself when T is a device value
  RCHECK(make_layout(h, "pn.lay", lay.lens, B, T, keep, FS2_HALO, &pn, st, T < 2 * H + 1 ? T : 2 * H + 1, lay.S_dev));
  if (lay.rows_hint > 0 && keep < T) {
    const long long est = (long long)lay.rows_hint + (long long)B * (keep - h->halo_keep) + 4 * H;
    pn.rows_hint = (int)(est < pn.R_cap ? est : pn.R_cap);
  }
  const size_t R = (size_t)pn.R_cap;
  WS(float, melg, "pn.mel", R * M);
  WS(float, postg, "pn.post", R * M);
  bf16 *melb = nullptr, *pa_b = nullptr, *pb_b = nullptr;
  float *pa = nullptr, *pb = nullptr;
  if (tc) {
    WS(bf16, t0, "pn.melb", np * R * M); melb = t0;
    WS(bf16, t1, "pn.a_b", np * R * P);  pa_b = t1;
    WS(bf16, t2, "pn.b_b", np * R * P);  pb_b = t2;
  } else {
    WS(float, t1, "pn.a", R * P); pa = t1;
    WS(float, t2, "pn.b", R * P); pb = t2;
  }
  ConvGemmArgs a = base_args(h->w->mel_linear, lay);
  a.A = dec; a.Ab = decb; a.epi = EPI_BIAS; a.mask_mode = MASK_GRID; a.dst_off = pn.off; a.dst_R_cap = pn.R_cap;
  a.out = melg; a.ldo = M; a.out_b = melb; a.ldob = M; a.out_user = mel; a.ldu = M;
  RCHECK(run_gemm(h, prec, a, st, "mel_linear"));
  {
    PROF("rows.fill_padded");
    HCHECK(rowops_fill_padded_rows(h->w->mel_linear.bias, M, lay, pn, melg, melb, (int)np, mel, st));
  }

  const float* in_f = melg; const bf16* in_b = melb;
  for (int i = 0; i < NL; ++i) {
    a = base_args(h->w->postnet[i], pn);
    a.A = in_f; a.Ab = in_b; a.mask_mode = MASK_GRID;
    if (i < NL - 1) {
      a.epi = EPI_TANH;
      float* of = (i & 1) ? pb : pa; bf16* ob = (i & 1) ? pb_b : pa_b;
      a.out = of; a.ldo = P; a.out_b = ob; a.ldob = P;
      in_f = of; in_b = ob;
    } else {
      a.epi = EPI_RES; a.residual = melg; a.out = postg; a.ldo = M; a.out_user = mel_post; a.ldu = M; a.out_user_B = B;
      a.user_cm = h->mel_post_cm;
    }
    RCHECK(run_gemm(h, prec, a, st, "postnet." + std::to_string(i)));
  }
  {
    PROF("rows.postnet_far");
    HCHECK(rowops_postnet_far_rows(postg, M, pn, B, H, mel_post, h->mel_post_cm, st));
  }
  return FS2_OK;
}

// packs mel_encoder.* (prenet, 4 x FFTBlock2) on first use; the weight block may be shared by several handles
int build_mel_encoder(fs2_handle* h, cudaStream_t st) {
  fs2_weights& W = *h->w;
  std::lock_guard<std::mutex> lock(W.menc_mu);
  if (W.menc_built) return FS2_OK;
  struct PdlOff { PdlOff() { ++g_fs2_pdl_off; } ~PdlOff() { --g_fs2_pdl_off; } } pdl_off;   // memcpys between small kernels
  const fs2_dims& d = h->dims;
  const int D = d.d_model, F = d.d_ffn;
  const float* t = nullptr;
  RCHECK(need(h, "mel_encoder.position_enc", {1, d.max_seq_len + 1, D}, &t));
  RCHECK(build_gemm(h, W.menc_pre1, {"mel_encoder.prenet.w_1.weight"}, {"mel_encoder.prenet.w_1.bias"}, D, d.n_mel, 1, nullptr, nullptr, st));
  RCHECK(build_gemm(h, W.menc_pre2, {"mel_encoder.prenet.w_2.weight"}, {"mel_encoder.prenet.w_2.bias"}, D, D, 1, nullptr, nullptr, st));
  W.menc.assign(d.n_dec_layers, CrossW());
  for (int l = 0; l < d.n_dec_layers; ++l) {
    CrossW& L = W.menc[l];
    const std::string p = "mel_encoder.layer_stack." + std::to_string(l), a = p + ".crs_attn.", f = p + ".pos_ffn.";
    RCHECK(build_gemm(h, L.q, {a + "w_qs.weight"}, {a + "w_qs.bias"}, D, D, 1, nullptr, nullptr, st));
    RCHECK(build_gemm(h, L.kv, {a + "w_ks.weight", a + "w_vs.weight"}, {a + "w_ks.bias", a + "w_vs.bias"}, D, D, 1, nullptr, nullptr, st));
    RCHECK(build_gemm(h, L.fc, {a + "fc.weight"}, {a + "fc.bias"}, D, D, 1, nullptr, nullptr, st));
    RCHECK(build_gemm(h, L.w1, {f + "w_1.weight"}, {f + "w_1.bias"}, F, D, d.ffn_k1, nullptr, nullptr, st));
    RCHECK(build_gemm(h, L.w2, {f + "w_2.weight"}, {f + "w_2.bias"}, D, F, d.ffn_k2, nullptr, nullptr, st));
    RCHECK(need(h, a + "layer_norm.weight", {D}, &t)); L.ln1_g = const_cast<float*>(t);
    RCHECK(need(h, a + "layer_norm.bias", {D}, &t));   L.ln1_b = const_cast<float*>(t);
    RCHECK(need(h, f + "layer_norm.weight", {D}, &t)); L.ln2_g = const_cast<float*>(t);
    RCHECK(need(h, f + "layer_norm.bias", {D}, &t));   L.ln2_b = const_cast<float*>(t);
  }
  HCHECK(cudaStreamSynchronize(st));
  W.menc_built = true;
  return FS2_OK;
}

// transformer/Models.py:140-173 MelEncoder.forward (eval): Prenet on the mels with frame 0 zeroed, + positional table, then
// n_dec_layers x FFTBlock2 (cross-attention over the phoneme sequence, FFN).  Both sequences live on their PADDED grids
// (ext = S: the reference computes -- and returns attention rows for -- padded query positions too).  GEMMs run in `prec`
// (fp32 FFMA or tcgen05 with bf16 / split operands); the cross-attention itself is the fp32 kernel of fs2_simt_attn.cu,
// which materialises the alignment.
int run_mel_encoder(fs2_handle* h, int prec, const float* src_seq, const float* mels, const int64_t* src_lens,
                    const int64_t* mel_lens, int B, int L, int T, float* out, float* attn, cudaStream_t st) {
  const fs2_dims& d = h->dims;
  const int D = d.d_model, F = d.d_ffn, H = d.n_heads, dk = D / H, M = d.n_mel;
  fs2_weights& W = *h->w;
  const bool tc = prec != FS2_PREC_FP32;
  const size_t np = (size_t)planes_of(prec);
  WS(int, sl32, "me.sl32", B);
  WS(int, ml32, "me.ml32", B);
  HCHECK(rowops_lens_to_i32(src_lens, B, L, sl32, st));
  HCHECK(rowops_lens_to_i32(mel_lens, B, T, ml32, st));
  RowLayout ls, lt;
  RCHECK(make_layout(h, "me.slay", sl32, B, L, L, FS2_HALO, &ls, st));
  RCHECK(make_layout(h, "me.tlay", ml32, B, T, T, FS2_HALO, &lt, st));
  const size_t Rs = (size_t)ls.R_cap, Rt = (size_t)lt.R_cap;
  WS(float, src_g, "me.src", Rs * D);
  WS(float, mel_g, "me.mel", Rt * M);
  WS(float, x, "me.x", Rt * D);
  WS(float, y, "me.y", Rt * D);
  WS(float, p1, "me.p1", Rt * D);
  WS(float, qbuf, "me.q", Rt * D);
  WS(float, kvbuf, "me.kv", Rs * 2 * D);
  WS(float, att, "me.att", Rt * D);
  float* hid = nullptr;
  bf16 *src_b = nullptr, *mel_b = nullptr, *xb = nullptr, *yb = nullptr, *p1b = nullptr, *attb = nullptr, *hidb = nullptr;
  if (tc) {
    WS(bf16, t0, "me.src_b", np * Rs * D); src_b = t0;
    WS(bf16, t1, "me.mel_b", np * Rt * M); mel_b = t1;
    WS(bf16, t2, "me.xb", np * Rt * D);    xb = t2;
    WS(bf16, t3, "me.yb", np * Rt * D);    yb = t3;
    WS(bf16, t4, "me.p1b", np * Rt * D);   p1b = t4;
    WS(bf16, t5, "me.attb", np * Rt * D);  attb = t5;
    WS(bf16, t6, "me.hidb", np * Rt * F);  hidb = t6;
  } else {
    WS(float, t7, "me.hid", Rt * F); hid = t7;
  }
  HCHECK(rowops_to_grid(src_seq, ls, D, src_g, D, 0, nullptr, st));
  HCHECK(make_shadow(src_g, Rs * D, prec, src_b, st));
  HCHECK(rowops_to_grid(mels, lt, M, mel_g, M, 0, nullptr, st));
  HCHECK(rowops_zero_first_rows(lt, M, mel_g, st));                     // Models.py:144-145
  HCHECK(make_shadow(mel_g, Rt * M, prec, mel_b, st));
  // Layers.py:24-28 Prenet: relu(w_2(relu(w_1(x)))) on every grid row (dropout is the identity in eval mode)
  ConvGemmArgs a = base_args(W.menc_pre1, lt);
  a.A = mel_g; a.Ab = mel_b; a.epi = EPI_RELU; a.mask_mode = MASK_GRID; a.out = tc ? nullptr : p1; a.ldo = D; a.out_b = p1b; a.ldob = D;
  RCHECK(run_gemm(h, prec, a, st, "menc.prenet1"));
  a = base_args(W.menc_pre2, lt);
  a.A = p1; a.Ab = p1b; a.epi = EPI_RELU; a.mask_mode = MASK_GRID; a.out = x; a.ldo = D;
  if (tc) a.out_planes = 1;
  RCHECK(run_gemm(h, prec, a, st, "menc.prenet2"));
  const float* pe = nullptr;
  if (T <= d.max_seq_len) pe = raw_ptr(h, "mel_encoder.position_enc");
  else RCHECK(position_table(h, 1, T, &pe, st));                        // same closed form as the decoder's (Models.py:148-153)
  HCHECK(rowops_add_pe(x, pe, lt, D, st));
  HCHECK(make_shadow(x, Rt * D, prec, xb, st));
  for (int l = 0; l < d.n_dec_layers; ++l) {
    CrossW& Lw = W.menc[l];
    a = base_args(Lw.q, lt);
    a.A = x; a.Ab = xb; a.epi = EPI_BIAS; a.mask_mode = MASK_GRID; a.out = qbuf; a.ldo = D;
    if (tc) a.out_planes = 1;
    RCHECK(run_gemm(h, prec, a, st, "menc.q"));
    a = base_args(Lw.kv, ls);
    a.A = src_g; a.Ab = src_b; a.epi = EPI_BIAS; a.mask_mode = MASK_GRID; a.out = kvbuf; a.ldo = 2 * D;
    if (tc) a.out_planes = 1;
    RCHECK(run_gemm(h, prec, a, st, "menc.kv"));
    {
      PROF("menc.cross_attn");
      HCHECK(simt_cross_attention_launch(qbuf, D, kvbuf, 2 * D, 0, D, lt, ls, H, dk, att, D,
                                         attn ? attn + (size_t)l * B * H * T * L : nullptr, st));
      HCHECK(make_shadow(att, Rt * D, prec, attb, st));
    }
    a = base_args(Lw.fc, lt);
    a.A = att; a.Ab = attb; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = x; a.ln_g = Lw.ln1_g; a.ln_b = Lw.ln1_b;
    a.out = y; a.ldo = D; a.out_b = yb; a.ldob = D;
    RCHECK(run_gemm(h, prec, a, st, "menc.fc_ln"));
    a = base_args(Lw.w1, lt);
    a.A = y; a.Ab = yb; a.epi = EPI_RELU; a.mask_mode = MASK_GRID; a.out = hid; a.ldo = F; a.out_b = hidb; a.ldob = F;
    RCHECK(run_gemm(h, prec, a, st, "menc.ffn_w1"));
    a = base_args(Lw.w2, lt);
    a.A = hid; a.Ab = hidb; a.epi = EPI_RES_LN; a.mask_mode = MASK_LEN; a.residual = y; a.ln_g = Lw.ln2_g; a.ln_b = Lw.ln2_b;
    a.out = x; a.ldo = D; a.out_b = xb; a.ldob = D;
    RCHECK(run_gemm(h, prec, a, st, "menc.ffn_w2_ln"));
  }
  HCHECK(rowops_from_grid(x, lt, D, out, st));
  return FS2_OK;
}

int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_last_error = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                   "); this library has no CPU fallback";
    (void)cudaGetLastError();
    return FS2_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    g_last_error = "device index out of range";
    return FS2_ERR_INVALID;
  }
  return FS2_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* fs2_version(void) { return "fs2_b200 0.1 (sm_100a)"; }

const char* fs2_last_error(const fs2_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int64_t fs2_launch_count(const fs2_handle*) { return (int64_t)g_fs2_launches.load(); }

int64_t fs2_last_frame_count(const fs2_handle* h) { return (h && h->have_stage1) ? (int64_t)h->st_frames : -1; }

int fs2_create(fs2_handle** out, const fs2_dims* d, int device) {
  if (!out || !d) { g_last_error = "null argument"; return FS2_ERR_INVALID; }
  *out = nullptr;
  RCHECK(check_device(device));
  // what the kernels are built for (config/LJSpeech/model.yaml is the only shipped config)
  if (d->d_model != 256 || d->vp_filter != 256) { g_last_error = "d_model and vp_filter must be 256"; return FS2_ERR_UNSUPPORTED; }
  if (d->n_heads <= 0 || d->d_model % d->n_heads || (d->d_model / d->n_heads != 128 && d->d_model / d->n_heads != 64)) {
    g_last_error = "head dim must be 64 or 128"; return FS2_ERR_UNSUPPORTED; }
  if (d->ffn_k1 % 2 == 0 || d->ffn_k2 % 2 == 0 || d->ffn_k1 > 2 * FS2_HALO + 1 || d->ffn_k2 > 2 * FS2_HALO + 1 ||
      d->vp_kernel != 3 || d->pn_kernel % 2 == 0 || d->pn_kernel > 2 * FS2_HALO + 1) {
    g_last_error = "conv kernels must be odd and <= 9; variance predictor kernel must be 3"; return FS2_ERR_UNSUPPORTED; }
  if (d->d_ffn % 256 || d->pn_dim % 256 || d->n_mel % 16 || d->n_mel > 128 || d->pn_layers < 2 || d->n_bins < 2 ||
      d->vocab < 1 || d->n_enc_layers < 1 || d->n_dec_layers < 1 || d->max_seq_len < 1) {
    g_last_error = "unsupported dims (d_ffn, pn_dim multiples of 256; n_mel multiple of 16 and <= 128)"; return FS2_ERR_UNSUPPORTED; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fs2_fail_cuda(e, "cudaSetDevice");
  fs2_handle* h = new fs2_handle();
  h->dims = *d;
  h->device = device;
  e = cudaMallocHost(reinterpret_cast<void**>(&h->host_tmax), 2 * sizeof(int));
  if (e != cudaSuccess) { delete h; return fs2_fail_cuda(e, "cudaMallocHost"); }
  *out = h;
  return FS2_OK;
}

void fs2_destroy(fs2_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  h->drop_graphs();
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  if (h->shape_host) cudaFreeHost(h->shape_host);
  if (h->shape_dev) cudaFree(h->shape_dev);
  h->w.reset();   // frees the weights when this was the last handle using them
  for (auto& kv : h->ws) cudaFree(kv.second.first);
  for (int i = 0; i < 2; ++i) if (h->pe_ext[i]) cudaFree(h->pe_ext[i]);
  for (auto& p : h->prof_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  if (h->host_tmax) cudaFreeHost(h->host_tmax);
  if (h->fill_stream) cudaStreamDestroy(h->fill_stream);
  delete h;
}

int fs2_set_precision(fs2_handle* h, int32_t enc, int32_t dec) {
  if (!h) return FS2_ERR_INVALID;
  if (enc < FS2_PREC_FP32 || enc > FS2_PREC_F16X2 || dec < FS2_PREC_FP32 || dec > FS2_PREC_F16X2)
    return h->fail(FS2_ERR_INVALID, "precision must be FS2_PREC_FP32, FS2_PREC_BF16, FS2_PREC_BF16X3 or FS2_PREC_F16X2");
  h->prec_enc = enc;
  h->prec_dec = dec;
  h->drop_graphs();
  return FS2_OK;
}

int fs2_set_row_packing(fs2_handle* h, int32_t keep_rows) {
  if (!h) return FS2_ERR_INVALID;
  if (keep_rows < 2) return h->fail(FS2_ERR_INVALID, "keep_rows must be >= 2 (the variance predictors' halo)");
  h->halo_keep = keep_rows;
  h->have_stage1 = false;
  h->drop_graphs();
  return FS2_OK;
}

int fs2_set_mel_post_layout(fs2_handle* h, int32_t channel_major) {
  if (!h) return FS2_ERR_INVALID;
  if (channel_major != 0 && channel_major != 1) return h->fail(FS2_ERR_INVALID, "channel_major must be 0 or 1");
  h->mel_post_cm = channel_major;
  h->drop_graphs();
  return FS2_OK;
}

int fs2_set_upsampler(fs2_handle* h, int32_t upsampler) {
  if (!h) return FS2_ERR_INVALID;
  if (upsampler != FS2_UPSAMPLER_HARD && upsampler != FS2_UPSAMPLER_GAUSSIAN)
    return h->fail(FS2_ERR_INVALID, "upsampler must be FS2_UPSAMPLER_HARD or FS2_UPSAMPLER_GAUSSIAN");
  h->upsampler = upsampler;
  h->drop_graphs();
  return FS2_OK;
}

int fs2_load_weights(fs2_handle* h, const fs2_weight_desc* descs, int32_t n) {
  // weight repacking interleaves cudaMemcpyAsync with small kernels: no programmatic overlap here at all
  struct PdlOff { PdlOff() { ++g_fs2_pdl_off; } ~PdlOff() { --g_fs2_pdl_off; } } pdl_off;
  if (!h || !descs || n <= 0) return h ? h->fail(FS2_ERR_INVALID, "null/empty weight list") : FS2_ERR_INVALID;
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = 0;
  // a fresh weight block: handles that share the previous one keep running on it until they load or share again
  h->loaded = false;
  h->drop_graphs();
  h->w = std::make_shared<fs2_weights>();
  h->w->device = h->device;
  h->have_stage1 = false;

  for (int i = 0; i < n; ++i) {
    const fs2_weight_desc& w = descs[i];
    if (!w.name || !w.data || w.ndim < 0 || w.ndim > 4) return h->fail(FS2_ERR_INVALID, "malformed weight descriptor");
    const std::string name(w.name);
    if (name.size() > 19 && name.compare(name.size() - 19, 19, "num_batches_tracked") == 0) continue;
    RawT t;
    t.numel = 1;
    for (int k = 0; k < w.ndim; ++k) { t.shape.push_back(w.shape[k]); t.numel *= w.shape[k]; }
    if (t.numel <= 0) return h->fail(FS2_ERR_INVALID, "empty tensor: " + name);
    HCHECK(cudaMalloc(reinterpret_cast<void**>(&t.d), sizeof(float) * (size_t)t.numel));
    cudaError_t e = cudaMemcpyAsync(t.d, w.data, sizeof(float) * (size_t)t.numel,
                                    w.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFree(t.d); return h->cuda_fail(e, "cudaMemcpyAsync(weight)"); }
    if (h->w->raw.count(name)) cudaFree(h->w->raw[name].d);
    h->w->raw[name] = t;
  }
  HCHECK(cudaStreamSynchronize(st));

  const fs2_dims& d = h->dims;
  const int D = d.d_model;
  const float* t = nullptr;
  RCHECK(need(h, "txt_encoder.src_word_emb.weight", {d.vocab, D}, &t));
  RCHECK(need(h, "txt_encoder.position_enc", {1, d.max_seq_len + 1, D}, &t));
  RCHECK(need(h, "mel_decoder.position_enc", {1, d.max_seq_len + 1, D}, &t));
  RCHECK(need(h, "variance_adaptor.pitch_bins", {d.n_bins - 1}, &t));
  RCHECK(need(h, "variance_adaptor.energy_bins", {d.n_bins - 1}, &t));
  RCHECK(need(h, "variance_adaptor.pitch_embedding.weight", {d.n_bins, D}, &t));
  RCHECK(need(h, "variance_adaptor.energy_embedding.weight", {d.n_bins, D}, &t));
  h->w->enc.assign(d.n_enc_layers, FftW());
  h->w->dec.assign(d.n_dec_layers, FftW());
  for (int l = 0; l < d.n_enc_layers; ++l) RCHECK(build_fft(h, h->w->enc[l], "txt_encoder.layer_stack." + std::to_string(l), st));
  for (int l = 0; l < d.n_dec_layers; ++l) RCHECK(build_fft(h, h->w->dec[l], "mel_decoder.layer_stack." + std::to_string(l), st));
  const char* which[3] = {"duration", "pitch", "energy"};
  for (int i = 0; i < 3; ++i) RCHECK(build_pred(h, h->w->pred[i], std::string("variance_adaptor.") + which[i] + "_predictor", st));
  RCHECK(build_gemm(h, h->w->mel_linear, {"mel_linear.weight"}, {"mel_linear.bias"}, d.n_mel, D, 1, nullptr, nullptr, st));
  h->w->postnet.assign(d.pn_layers, GemmW());
  for (int i = 0; i < d.pn_layers; ++i) {
    const int cin = i == 0 ? d.n_mel : d.pn_dim, cout = i == d.pn_layers - 1 ? d.n_mel : d.pn_dim;
    const std::string p = "postnet.convolutions." + std::to_string(i);
    const float *cb, *g, *b, *mean, *var;
    RCHECK(need(h, p + ".0.conv.bias", {cout}, &cb));
    RCHECK(need(h, p + ".1.weight", {cout}, &g));
    RCHECK(need(h, p + ".1.bias", {cout}, &b));
    RCHECK(need(h, p + ".1.running_mean", {cout}, &mean));
    RCHECK(need(h, p + ".1.running_var", {cout}, &var));
    float *scale = nullptr, *fb = nullptr;
    RCHECK(dev_alloc(h, &scale, (size_t)cout));
    RCHECK(dev_alloc(h, &fb, (size_t)cout));
    HCHECK(rowops_bn_fold(cb, g, b, mean, var, cout, 1e-5f, scale, fb, st));
    RCHECK(build_gemm(h, h->w->postnet[i], {p + ".0.conv.weight"}, {}, cout, cin, d.pn_kernel, scale, fb, st));
  }
  HCHECK(cudaStreamSynchronize(st));
  h->loaded = true;
  return FS2_OK;
}

int fs2_share_weights(fs2_handle* h, const fs2_handle* src) {
  if (!h || !src) return h ? h->fail(FS2_ERR_INVALID, "share_weights: null handle") : FS2_ERR_INVALID;
  if (h == src) return FS2_OK;
  if (!src->loaded || !src->w) return h->fail(FS2_ERR_STATE, "share_weights: the source handle has no weights loaded");
  if (h->device != src->device) return h->fail(FS2_ERR_INVALID, "share_weights: handles live on different devices");
  if (memcmp(&h->dims, &src->dims, sizeof(fs2_dims)) != 0) return h->fail(FS2_ERR_INVALID, "share_weights: dims differ");
  h->drop_graphs();
  h->w = src->w;
  h->loaded = true;
  h->have_stage1 = false;
  return FS2_OK;
}

// ---------------------------------------------------------------------------------------------
}  // extern "C"

// stage 1 up to and including the duration scan; leaves {max mel_len, sum mel_lens} in the workspace int pair *tmax_out
static int stage1_enqueue(fs2_handle* h, const int64_t* texts, const int64_t* src_lens, int32_t B, int32_t L,
                          float p_control, float e_control, float d_control, float* log_d, float* d_rounded,
                          int64_t* mel_lens, uint8_t* src_mask, float* pitch_ph, float* energy_ph, int** tmax_out,
                          void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h) return FS2_ERR_INVALID;
  if (!h->loaded) return h->fail(FS2_ERR_STATE, "fs2_load_weights has not succeeded on this handle");
  if (!texts || !src_lens || !log_d || !d_rounded || !mel_lens || B <= 0 || L <= 0)
    return h->fail(FS2_ERR_INVALID, "stage1: null pointer or empty batch");
  if (L > 8192) return h->fail(FS2_ERR_UNSUPPORTED, "stage1: L > 8192");
  const fs2_dims& d = h->dims;
  if ((d.pitch_phoneme_level && !pitch_ph) || (d.energy_phoneme_level && !energy_ph))
    return h->fail(FS2_ERR_INVALID, "stage1: phoneme-level pitch/energy output pointer missing");
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = d.d_model;
  h->have_stage1 = false;

  WS(int, lens32, "s1.lens32", B);
  WS(int, cum, "s1.cum", (size_t)B * L);
  WS(int, mlens32, "s1.mel_lens32", B);
  WS(int, tmax_dev, "s1.tmax", 2);   // [0] max frames of an utterance, [1] frames of the batch
  const int* Ldev = h->cur_Ldev;   // graph entry point: L is the bucket's bound, the true L is read from device memory
  HCHECK(rowops_lens_to_i32(src_lens, B, L, lens32, st, tmax_dev, Ldev));   // also clears the {T_max, frames} accumulators
  // packed phoneme rows: valid rows + the 2 padded rows the duration predictor's convolutions can see (+ zero halo)
  RowLayout lay;
  RCHECK(make_layout(h, "s1.lay", lens32, B, L, h->halo_keep, FS2_HALO, &lay, st, 0, Ldev));
  const size_t R = (size_t)lay.R_cap;
  WS(float, x, "s1.x", R * D);
  bf16* xb = nullptr;
  if (h->prec_enc != FS2_PREC_FP32) {
    WS(bf16, t, "s1.xb", (size_t)planes_of(h->prec_enc) * R * D);
    xb = t;
  }
  const float* pe = nullptr;
  RCHECK(position_table(h, 0, L, &pe, st));
  {
    PROF("rows.embed_pe");
    if (src_mask) HCHECK(rowops_mask(src_lens, nullptr, B, L, src_mask, st, log_d, nullptr, Ldev));   // + log_d cleared in the same pass
    HCHECK(rowops_embed_pe(texts, raw_ptr(h, "txt_encoder.src_word_emb.weight"), pe, d.vocab, lay, D, x, xb,
                           planes_of(h->prec_enc), nullptr, st));
  }
  RCHECK(run_fft_stack(h, h->w->enc, 0, d.n_enc_layers, h->prec_enc, x, xb, lay, st));
  // modules.py:116 duration predictor on the encoder output
  RCHECK(run_predictor(h, h->w->pred[0], h->prec_enc, x, xb, lay, log_d, st, src_mask != nullptr));
  // modules.py:117-126 phoneme-level variants
  if (d.pitch_phoneme_level) {
    RCHECK(run_predictor(h, h->w->pred[1], h->prec_enc, x, xb, lay, pitch_ph, st));
    HCHECK(rowops_variance_embed(pitch_ph, p_control, raw_ptr(h, "variance_adaptor.pitch_bins"), d.n_bins,
                                 raw_ptr(h, "variance_adaptor.pitch_embedding.weight"), nullptr, x, xb,
                                 planes_of(h->prec_enc), lay, D, nullptr, st));
  }
  if (d.energy_phoneme_level) {
    RCHECK(run_predictor(h, h->w->pred[2], h->prec_enc, x, xb, lay, energy_ph, st));
    HCHECK(rowops_variance_embed(energy_ph, e_control, raw_ptr(h, "variance_adaptor.energy_bins"), d.n_bins,
                                 raw_ptr(h, "variance_adaptor.energy_embedding.weight"), nullptr, x, xb,
                                 planes_of(h->prec_enc), lay, D, nullptr, st));
  }
  // modules.py:132-135 + LengthRegulator bookkeeping
  {
    PROF("rows.round_scan");
    HCHECK(rowops_round_scan(log_d, d_control, d_rounded, B, L, cum, mel_lens, mlens32, tmax_dev, st, Ldev));
  }
  h->st_B = B; h->st_L = L;
  h->st_Ldev = Ldev;
  h->st_enc_out = x;
  h->st_lay1 = lay;
  *tmax_out = tmax_dev;
  return FS2_OK;
}

extern "C" {

int fs2_forward_stage1(fs2_handle* h, const int64_t* texts, const int64_t* src_lens, int32_t B, int32_t L,
                       float p_control, float e_control, float d_control, float* log_d, float* d_rounded,
                       int64_t* mel_lens, uint8_t* src_mask, float* pitch_ph, float* energy_ph, int32_t* T_max_out,
                       void* stream) {
  if (h && !T_max_out) return h->fail(FS2_ERR_INVALID, "stage1: null T_max_out");
  int* tmax_dev = nullptr;
  RCHECK(stage1_enqueue(h, texts, src_lens, B, L, p_control, e_control, d_control, log_d, d_rounded, mel_lens, src_mask,
                        pitch_ph, energy_ph, &tmax_dev, stream));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HCHECK(cudaMemcpyAsync(h->host_tmax, tmax_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  HCHECK(cudaStreamSynchronize(st));  // the one data-dependent size of the path
  *T_max_out = h->host_tmax[0];
  h->have_stage1 = true;
  h->st_Tmax = h->host_tmax[0];
  h->st_frames = h->host_tmax[1];
  return FS2_OK;
}

int fs2_forward_stage1_async(fs2_handle* h, const int64_t* texts, const int64_t* src_lens, int32_t B, int32_t L,
                             float p_control, float e_control, float d_control, float* log_d, float* d_rounded,
                             int64_t* mel_lens, uint8_t* src_mask, float* pitch_ph, float* energy_ph, int32_t* tmax_user,
                             void* stream) {
  if (h && !tmax_user) return h->fail(FS2_ERR_INVALID, "stage1_async: null tmax_dev");
  int* tmax_dev = nullptr;
  RCHECK(stage1_enqueue(h, texts, src_lens, B, L, p_control, e_control, d_control, log_d, d_rounded, mel_lens, src_mask,
                        pitch_ph, energy_ph, &tmax_dev, stream));
  HCHECK(cudaMemcpyAsync(tmax_user, tmax_dev, 2 * sizeof(int), cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  h->st_Tmax = -1;   // unknown until fs2_forward_stage1_commit
  return FS2_OK;
}

int fs2_forward_stage1_commit(fs2_handle* h, int32_t T_max, int32_t frames) {
  if (!h) return FS2_ERR_INVALID;
  if (!h->loaded || h->st_B <= 0 || h->st_Tmax != -1)
    return h->fail(FS2_ERR_STATE, "stage1_commit without a preceding fs2_forward_stage1_async");
  if (T_max < 0 || frames < 0) return h->fail(FS2_ERR_INVALID, "stage1_commit: negative size");
  h->st_Tmax = T_max;
  h->st_frames = frames;
  h->have_stage1 = true;
  return FS2_OK;
}

}  // extern "C"

static int stage2_check(fs2_handle* h, int32_t T, float* mel, float* mel_post, float* pitch, float* energy) {
  if (!h->loaded || !h->have_stage1) return h->fail(FS2_ERR_STATE, "stage2 called before a successful stage1");
  if (T < h->st_Tmax) return h->fail(FS2_ERR_INVALID, "stage2: T smaller than the stage-1 maximum mel length");
  const fs2_dims& d = h->dims;
  if (T == 0) return FS2_OK;
  if (T > FS2_MAX_ROWS_PER_UTT) return h->fail(FS2_ERR_UNSUPPORTED, "stage2: more than 65535 mel frames per utterance");
  if (!mel || !mel_post || (!d.pitch_phoneme_level && !pitch) || (!d.energy_phoneme_level && !energy))
    return h->fail(FS2_ERR_INVALID, "stage2: null output pointer");
  if (h->upsampler == FS2_UPSAMPLER_GAUSSIAN && (d.pitch_phoneme_level || d.energy_phoneme_level))
    return h->fail(FS2_ERR_UNSUPPORTED, "stage2: the Gaussian upsampler is built for frame-level pitch / energy (with "
                   "phoneme-level features the padded phoneme rows are not zero and would need their embedding rows)");
  return FS2_OK;
}

// everything stage 2 enqueues.  T = rows per utterance of the outputs: the exact value, or (h->cur_Tdev set) the bucket's
// upper bound while the true value is read from device memory by the kernels.
static int stage2_enqueue(fs2_handle* h, int32_t T, float p_control, float e_control, float* mel, float* mel_post,
                          float* pitch, float* energy, uint8_t* mel_mask, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  const fs2_dims& d = h->dims;
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int B = h->st_B, L = h->st_L, D = d.d_model;
  const int* Tdev = h->cur_Tdev;
  const int* cum = reinterpret_cast<int*>(h->ws["s1.cum"].first);
  const int* mlens32 = reinterpret_cast<int*>(h->ws["s1.mel_lens32"].first);
  const bool pitch_fl = !d.pitch_phoneme_level, energy_fl = !d.energy_phoneme_level;

  // packed frame rows: valid frames + the 2 padded rows the pitch / energy predictors can see (+ zero halo)
  RowLayout lay;
  RCHECK(make_layout(h, "s2.lay", mlens32, B, T, h->halo_keep, FS2_HALO, &lay, st, 0, Tdev));
  // rows in use, estimated on the host from the batch's frame count (exact up to the 8-row alignment of each utterance);
  // a graph is replayed for other batches of its bucket, so it keeps the allocation's bound
  if (h->halo_keep < T && !Tdev) {
    const long long est = (long long)h->st_frames + (long long)B * (h->halo_keep + FS2_HALO + FS2_ROW_ALIGN / 2);
    lay.rows_hint = (int)(est < lay.R_cap ? est : lay.R_cap);
  }
  const size_t R = (size_t)lay.R_cap;
  WS(float, x, "s2.x", R * D);
  bf16* xb = nullptr;
  {
    const int np = planes_of(h->prec_dec) > planes_of(h->prec_enc) ? planes_of(h->prec_dec) : planes_of(h->prec_enc);
    if (np > 0) {
      WS(bf16, t, "s2.xb", (size_t)np * R * D);
      xb = t;
    }
  }
  // every producer of x writes the bf16 shadow in the precision of its NEXT consumer
  const int first_prec = (pitch_fl || energy_fl) ? h->prec_enc : FS2_PREC_FP32;
  {
    PROF("rows.length_regulate");
    if (mel_mask) HCHECK(rowops_mask(nullptr, mlens32, B, T, mel_mask, st, pitch_fl ? pitch : nullptr, energy_fl ? energy : nullptr, Tdev));
    // modules.py:136 length regulator (hard): gathers encoder rows (stage-1 layout) into frame rows (stage-2 layout)
    if (h->upsampler == FS2_UPSAMPLER_GAUSSIAN)   // modules.py:162-192 in the LengthRegulator's place (README.md:10 of the reference)
      HCHECK(rowops_gaussian_regulate(h->st_enc_out, h->st_lay1.off, h->st_lay1.lens, cum, L, D, lay, x, xb,
                                      planes_of(first_prec), st, h->st_Ldev));
    else
      HCHECK(rowops_length_regulate(h->st_enc_out, h->st_lay1.off, 0, cum, L, D, lay, x, xb, planes_of(first_prec), st,
                                    h->st_Ldev));
  }
  const float* pe = nullptr;
  RCHECK(position_table(h, 1, T, &pe, st));
  // modules.py:139-149 frame-level pitch then energy (energy sees x + pitch embedding)
  if (pitch_fl) {
    RCHECK(run_predictor(h, h->w->pred[1], h->prec_enc, x, xb, lay, pitch, st, mel_mask != nullptr));
    PROF("rows.variance_embed");
    HCHECK(rowops_variance_embed(pitch, p_control, raw_ptr(h, "variance_adaptor.pitch_bins"), d.n_bins,
                                 raw_ptr(h, "variance_adaptor.pitch_embedding.weight"), energy_fl ? nullptr : pe, x, xb,
                                 planes_of(energy_fl ? h->prec_enc : h->prec_dec), lay, D, nullptr, st));
  }
  if (energy_fl) {
    RCHECK(run_predictor(h, h->w->pred[2], h->prec_enc, x, xb, lay, energy, st, mel_mask != nullptr));
    // fused: + energy embedding, + decoder positional encoding (Models.py:231-233), shadow for the decoder
    PROF("rows.variance_embed");
    HCHECK(rowops_variance_embed(energy, e_control, raw_ptr(h, "variance_adaptor.energy_bins"), d.n_bins,
                                 raw_ptr(h, "variance_adaptor.energy_embedding.weight"), pe, x, xb,
                                 planes_of(h->prec_dec), lay, D, nullptr, st));
  }
  if (!pitch_fl && !energy_fl) {  // both phoneme-level: only the decoder's positional add remains
    HCHECK(rowops_add_pe(x, pe, lay, D, st));
    HCHECK(make_shadow(x, R * D, h->prec_dec, xb, st));
  }
  RCHECK(run_fft_stack(h, h->w->dec, 0, d.n_dec_layers, h->prec_dec, x, xb, lay, st));
  RCHECK(run_mel_postnet(h, h->prec_dec, x, xb, lay, mel, mel_post, st));
  return FS2_OK;
}

extern "C" {

int fs2_forward_stage2(fs2_handle* h, int32_t T, float p_control, float e_control, float* mel, float* mel_post,
                       float* pitch, float* energy, uint8_t* mel_mask, void* stream) {
  if (!h) return FS2_ERR_INVALID;
  RCHECK(stage2_check(h, T, mel, mel_post, pitch, energy));
  if (T == 0) return FS2_OK;  // degenerate batch (all durations zero): nothing to write
  return stage2_enqueue(h, T, p_control, e_control, mel, mel_post, pitch, energy, mel_mask, stream);
}

}  // extern "C"

// Runs `enqueue(stream)` (a lambda that launches one stage on the stream it is given) through the handle's graph cache: first sighting of a key =
// plain launches (sizes the workspace), second = stream capture + instantiate + launch, afterwards = one cudaGraphLaunch.
// Returns FS2_OK with *replayed = true when the work went out as a graph launch (no host-side stage code ran).
template <typename F>
static int run_graphed(fs2_handle* h, const std::string& key, cudaStream_t st, bool* replayed, F&& enqueue) {
  // bounded cache: a long-running service with many (B, bucket) combinations starts over instead of growing for ever
  if (h->graphs.size() >= 512 && !h->graphs.count(key)) h->drop_graphs();
  fs2_handle::GraphEntry& e = h->graphs[key];
  *replayed = false;
  if (e.exec && e.gen == h->gen) {
    HCHECK(cudaGraphLaunch(e.exec, st));
    g_fs2_launches += e.kernels;
    g_fs2_plain_next = 1;
    ++h->graph_replays;
    *replayed = true;
    return FS2_OK;
  }
  if (e.exec) { cudaGraphExecDestroy(e.exec); e.exec = nullptr; }   // captured under an older generation
  if (e.seen <= 0 || h->prof_on) {       // first sighting (or tracing on, or capture known to fail): plain launches
    if (e.seen == 0) e.seen = 1;
    return enqueue(st);
  }
  if (!h->capture_stream) HCHECK(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
  cudaStream_t cs = h->capture_stream;
  const unsigned long long gen0 = h->gen;
  const long long l0 = g_fs2_launches.load();
  // relaxed mode: the stage code may still call cudaMalloc / cudaFuncSetAttribute; it must not synchronise `st`
  HCHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
  const int rc = enqueue(cs);
  cudaGraph_t g = nullptr;
  cudaError_t ce = cudaStreamEndCapture(cs, &g);
  cudaGraphExec_t ex = nullptr;
  if (rc == FS2_OK && ce == cudaSuccess && g && h->gen == gen0) ce = cudaGraphInstantiate(&ex, g, 0);
  if (g) cudaGraphDestroy(g);
  if (rc != FS2_OK) { (void)cudaGetLastError(); return rc; }
  if (ce != cudaSuccess || !ex || h->gen != gen0) {
    // nothing has run yet (the launches were only recorded).  A buffer moved during the capture: try again next time;
    // capture / instantiation itself failed: this key stays on plain launches.
    (void)cudaGetLastError();
    if (ex) cudaGraphExecDestroy(ex);
    if (h->gen == gen0) e.seen = -1;
    g_fs2_launches -= g_fs2_launches.load() - l0;
    return enqueue(st);
  }
  e.exec = ex;
  e.gen = gen0;
  e.kernels = g_fs2_launches.load() - l0;
  ++h->graph_captures;
  HCHECK(cudaGraphLaunch(ex, st));
  g_fs2_plain_next = 1;
  return FS2_OK;
}

static int graph_shape_buffers(fs2_handle* h) {
  if (!h->shape_host) HCHECK(cudaMallocHost(reinterpret_cast<void**>(&h->shape_host), 4 * sizeof(int)));
  if (!h->shape_dev) {
    HCHECK(cudaMalloc(reinterpret_cast<void**>(&h->shape_dev), 4 * sizeof(int)));
    ++h->gen;
  }
  return FS2_OK;
}

extern "C" {

int fs2_forward_stage1_graph(fs2_handle* h, const int64_t* texts, const int64_t* src_lens, int32_t B, int32_t L,
                             int32_t L_cap, float p_control, float e_control, float d_control, float* log_d,
                             float* d_rounded, int64_t* mel_lens, uint8_t* src_mask, float* pitch_ph, float* energy_ph,
                             int32_t* T_max_out, void* stream) {
  if (!h) return FS2_ERR_INVALID;
  if (!T_max_out) return h->fail(FS2_ERR_INVALID, "stage1_graph: null T_max_out");
  if (!h->loaded) return h->fail(FS2_ERR_STATE, "fs2_load_weights has not succeeded on this handle");
  if (!texts || !src_lens || B <= 0 || L <= 0 || L_cap < L) return h->fail(FS2_ERR_INVALID, "stage1_graph: bad argument (need 0 < L <= L_cap)");
  // one bucket never straddles max_seq_len: below it the positional table is the checkpoint's, above it the computed one
  if (L <= h->dims.max_seq_len && L_cap > h->dims.max_seq_len) L_cap = h->dims.max_seq_len;
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RCHECK(graph_shape_buffers(h));
  // the caller's input tensors change from call to call, the graph reads fixed staging copies ([B, L] with the TRUE L)
  WS(int64_t, g_texts, "g1.texts", (size_t)B * L_cap);
  WS(int64_t, g_lens, "g1.lens", (size_t)B);
  h->shape_host[0] = L;
  HCHECK(cudaMemcpyAsync(g_texts, texts, sizeof(int64_t) * (size_t)B * L, cudaMemcpyDeviceToDevice, st));
  HCHECK(cudaMemcpyAsync(g_lens, src_lens, sizeof(int64_t) * (size_t)B, cudaMemcpyDeviceToDevice, st));
  HCHECK(cudaMemcpyAsync(h->shape_dev, h->shape_host, sizeof(int), cudaMemcpyHostToDevice, st));
  char key[512];
  snprintf(key, sizeof key, "s1|%d|%d|%p|%p|%p|%p|%p|%p|%a|%a|%a", B, L_cap, (void*)log_d, (void*)d_rounded, (void*)mel_lens,
           (void*)src_mask, (void*)pitch_ph, (void*)energy_ph, p_control, e_control, d_control);
  bool replayed = false;
  h->cur_Ldev = h->shape_dev;
  const int rc = run_graphed(h, key, st, &replayed, [&](cudaStream_t s_) -> int {
    int* tm = nullptr;
    RCHECK(stage1_enqueue(h, g_texts, g_lens, B, L_cap, p_control, e_control, d_control, log_d, d_rounded, mel_lens,
                          src_mask, pitch_ph, energy_ph, &tm, s_));
    HCHECK(cudaMemcpyAsync(h->host_tmax, tm, 2 * sizeof(int), cudaMemcpyDeviceToHost, s_));
    return FS2_OK;
  });
  h->cur_Ldev = nullptr;
  RCHECK(rc);
  fs2_handle::GraphEntry& e = h->graphs[key];
  if (replayed) {
    h->st_B = e.st_B; h->st_L = e.st_L; h->st_enc_out = e.st_enc_out; h->st_lay1 = e.st_lay1; h->st_Ldev = e.st_Ldev;
  } else {
    e.st_B = h->st_B; e.st_L = h->st_L; e.st_enc_out = h->st_enc_out; e.st_lay1 = h->st_lay1; e.st_Ldev = h->st_Ldev;
  }
  HCHECK(cudaStreamSynchronize(st));  // the one data-dependent size of the path
  *T_max_out = h->host_tmax[0];
  h->have_stage1 = true;
  h->st_Tmax = h->host_tmax[0];
  h->st_frames = h->host_tmax[1];
  return FS2_OK;
}

int fs2_forward_stage2_graph(fs2_handle* h, int32_t T, int32_t T_cap, float p_control, float e_control, float* mel,
                             float* mel_post, float* pitch, float* energy, uint8_t* mel_mask, void* stream) {
  if (!h) return FS2_ERR_INVALID;
  RCHECK(stage2_check(h, T, mel, mel_post, pitch, energy));
  if (T == 0) return FS2_OK;
  if (T_cap < T) return h->fail(FS2_ERR_INVALID, "stage2_graph: T_cap smaller than T");
  if (T <= h->dims.max_seq_len && T_cap > h->dims.max_seq_len) T_cap = h->dims.max_seq_len;
  if (T_cap > FS2_MAX_ROWS_PER_UTT) T_cap = FS2_MAX_ROWS_PER_UTT;
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RCHECK(graph_shape_buffers(h));
  h->shape_host[1] = T;
  HCHECK(cudaMemcpyAsync(h->shape_dev + 1, h->shape_host + 1, sizeof(int), cudaMemcpyHostToDevice, st));
  char key[512];
  snprintf(key, sizeof key, "s2|%d|%d|%d|%d|%p|%p|%p|%p|%p|%a|%a", h->st_B, h->st_L, h->st_Ldev ? 1 : 0, T_cap, (void*)mel,
           (void*)mel_post, (void*)pitch, (void*)energy, (void*)mel_mask, p_control, e_control);
  bool replayed = false;
  h->cur_Tdev = h->shape_dev + 1;
  const int rc = run_graphed(h, key, st, &replayed, [&](cudaStream_t s_) -> int {
    return stage2_enqueue(h, T_cap, p_control, e_control, mel, mel_post, pitch, energy, mel_mask, s_);
  });
  h->cur_Tdev = nullptr;
  return rc;
}

/* {graphs held, replays, captures} of this handle */
int fs2_graph_stats(const fs2_handle* h, int64_t* n_graphs, int64_t* replays, int64_t* captures) {
  if (!h) return FS2_ERR_INVALID;
  int64_t n = 0;
  for (const auto& kv : h->graphs) n += kv.second.exec != nullptr;
  if (n_graphs) *n_graphs = n;
  if (replays) *replays = h->graph_replays;
  if (captures) *captures = h->graph_captures;
  return FS2_OK;
}

// ---------------------------------------------------------------------------------------------
// tracing
int fs2_profile_enable(fs2_handle* h, int32_t on) {
  if (!h) return FS2_ERR_INVALID;
  if (on < 0 || on > 2) return h->fail(FS2_ERR_INVALID, "profile_enable: 0 = off, 1 = per kernel class, 2 = per segment");
  h->prof_on = on;
  return FS2_OK;
}

int fs2_profile_reset(fs2_handle* h) {
  if (!h) return FS2_ERR_INVALID;
  HCHECK(cudaSetDevice(h->device));
  HCHECK(cudaDeviceSynchronize());
  for (auto& p : h->prof_pending) { h->prof_pool.push_back(p.a); h->prof_pool.push_back(p.b); }
  h->prof_pending.clear();
  h->prof_names.clear(); h->prof_ms.clear(); h->prof_launches.clear();
  return FS2_OK;
}

int fs2_profile_read(fs2_handle* h, fs2_profile_entry* out, int32_t max_entries, int32_t* n_out) {
  if (!h || !n_out || (max_entries > 0 && !out)) return h ? h->fail(FS2_ERR_INVALID, "profile_read: null argument") : FS2_ERR_INVALID;
  HCHECK(cudaSetDevice(h->device));
  HCHECK(cudaDeviceSynchronize());
  for (auto& p : h->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) h->prof_ms[p.slot] += (double)ms;
    h->prof_pool.push_back(p.a);
    h->prof_pool.push_back(p.b);
  }
  h->prof_pending.clear();
  const int n = (int)h->prof_names.size();
  *n_out = n;
  for (int i = 0; i < n && i < max_entries; ++i) {
    memset(&out[i], 0, sizeof out[i]);
    strncpy(out[i].name, h->prof_names[i].c_str(), sizeof(out[i].name) - 1);
    out[i].launches = h->prof_launches[i];
    out[i].ms = h->prof_ms[i];
  }
  return FS2_OK;
}

// ---------------------------------------------------------------------------------------------
// stand-alone operators
int fs2_round_durations(const float* log_d, int64_t n, float d_control, float* out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!log_d || !out || n < 0) { g_last_error = "null argument"; return FS2_ERR_INVALID; }
  FS2_CUDA_CHECK(rowops_round_durations(log_d, n, d_control, out, reinterpret_cast<cudaStream_t>(stream)));
  return FS2_OK;
}

int fs2_duration_scan(const float* dd, int32_t B, int32_t L, int32_t* cum, int64_t* mel_lens, int32_t* T_max_out,
                      void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!dd || !cum || !mel_lens || !T_max_out || B <= 0 || L <= 0) { g_last_error = "bad argument"; return FS2_ERR_INVALID; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int* tmax = nullptr;
  FS2_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&tmax), 2 * sizeof(int)));   // [0] max, [1] batch total
  cudaError_t e = cudaMemsetAsync(tmax, 0, 2 * sizeof(int), st);
  g_fs2_plain_next = 1;
  if (e == cudaSuccess) e = rowops_duration_scan(dd, B, L, cum, mel_lens, nullptr, tmax, st);
  int host = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&host, tmax, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(tmax);
  if (e != cudaSuccess) return fs2_fail_cuda(e, "fs2_duration_scan");
  *T_max_out = host;
  return FS2_OK;
}

int fs2_length_regulate(const float* x, const int32_t* cum, int32_t B, int32_t L, int32_t D, int32_t T, float* out,
                        void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!x || !cum || !out || B <= 0 || L <= 0 || D <= 0 || D % 4 || T < 0 || B > 65535 || T > FS2_MAX_ROWS_PER_UTT) {
    g_last_error = "bad argument"; return FS2_ERR_INVALID; }
  if (T == 0) return FS2_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // dense layouts on both sides: x [B,L,D] -> out [B,T,D], no halo rows
  TmpLayout tl;
  cudaError_t e = tl.build(nullptr, B, T, T, 0, st);
  if (e == cudaSuccess) e = rowops_length_regulate(x, nullptr, L, cum, L, D, tl.lay, out, nullptr, 0, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fs2_fail_cuda(e, "fs2_length_regulate");
  return FS2_OK;
}

int fs2_gaussian_upsample(const float* x, const float* dd, int32_t B, int32_t L, int32_t D, int32_t T, int32_t T_w,
                          float* out, float* s, float* w, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!x || !dd || !out || B <= 0 || L <= 0 || D <= 0 || D % 4 || T < 0 || T_w < 0 || T_w > T) {
    g_last_error = "bad argument"; return FS2_ERR_INVALID; }
  FS2_CUDA_CHECK(rowops_gaussian_upsample(x, dd, B, L, D, T, T_w, out, s, w, reinterpret_cast<cudaStream_t>(stream)));
  return FS2_OK;
}

int fs2_mask_from_lengths(const int64_t* lens, int32_t B, int32_t max_len, uint8_t* mask, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!lens || !mask || B <= 0 || max_len < 0) { g_last_error = "bad argument"; return FS2_ERR_INVALID; }
  FS2_CUDA_CHECK(rowops_mask(lens, nullptr, B, max_len, mask, reinterpret_cast<cudaStream_t>(stream)));
  return FS2_OK;
}

// ---------------------------------------------------------------------------------------------
// hand-off to the consumers of the result (kernels in fs2_handoff.cu)
int fs2_pack_valid_rows(const float* src, const int64_t* lens, int32_t B, int32_t S, int32_t C, int32_t channel_major,
                        int64_t* offsets, float* dst, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (B == 0) return FS2_OK;
  if (S == 0 && B > 0) {   // no rows at all (degenerate T == 0 batch): every offset is 0; src / dst may be empty tensors (NULL)
    if (offsets) FS2_CUDA_CHECK(cudaMemsetAsync(offsets, 0, sizeof(int64_t) * ((size_t)B + 1), reinterpret_cast<cudaStream_t>(stream)));
    return FS2_OK;
  }
  if (!src || !lens || !dst || B < 0 || B > 65535 || S < 0 || C <= 0 || (channel_major != 0 && channel_major != 1)) {
    char buf[256];
    snprintf(buf, sizeof buf, "fs2_pack_valid_rows: bad argument (src, lens, dst non-null; 0 <= B <= 65535; S >= 0; C > 0): "
             "src=%p lens=%p dst=%p B=%d S=%d C=%d channel_major=%d", (const void*)src, (const void*)lens, (void*)dst, B, S, C,
             channel_major);
    g_last_error = buf;
    return FS2_ERR_INVALID;
  }
  FS2_CUDA_CHECK(handoff_pack_valid_rows(src, lens, B, S, C, channel_major, offsets, dst,
                                         reinterpret_cast<cudaStream_t>(stream)));
  return FS2_OK;
}

int fs2_wav_to_int16(const float* wav, const int64_t* lens, int32_t B, int64_t N, float max_wav_value, int64_t* offsets,
                     int16_t* dst, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (B == 0) return FS2_OK;
  if (N == 0 && B > 0) {   // empty waveforms: every offset is 0
    if (offsets) FS2_CUDA_CHECK(cudaMemsetAsync(offsets, 0, sizeof(int64_t) * ((size_t)B + 1), reinterpret_cast<cudaStream_t>(stream)));
    return FS2_OK;
  }
  if (!wav || !dst || B < 0 || B > 65535 || N < 0) {
    char buf[256];
    snprintf(buf, sizeof buf, "fs2_wav_to_int16: bad argument (wav, dst non-null; 0 <= B <= 65535; N >= 0): wav=%p dst=%p B=%d "
             "N=%lld", (const void*)wav, (void*)dst, B, (long long)N);
    g_last_error = buf;
    return FS2_ERR_INVALID;
  }
  FS2_CUDA_CHECK(handoff_wav_to_int16(wav, lens, B, N, max_wav_value, offsets, dst, reinterpret_cast<cudaStream_t>(stream)));
  return FS2_OK;
}

// ---------------------------------------------------------------------------------------------
// per-operator entry points (unit parity)
int fs2_op_sinusoid_table(fs2_handle* h, int32_t n_pos, float* out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h || !out || n_pos <= 0) return FS2_ERR_INVALID;
  HCHECK(cudaSetDevice(h->device));
  return sinusoid_table_host(h, n_pos, h->dims.d_model, out, reinterpret_cast<cudaStream_t>(stream));
}

int fs2_op_embed_pe(fs2_handle* h, const int64_t* texts, int32_t B, int32_t L, float* out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h || !h->loaded) return FS2_ERR_STATE;
  if (!texts || !out || B <= 0 || L <= 0) return h->fail(FS2_ERR_INVALID, "bad argument");
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* pe = nullptr;
  RCHECK(position_table(h, 0, L, &pe, st));
  RowLayout lay;
  RCHECK(make_layout(h, "op.lay", nullptr, B, L, L, 0, &lay, st));
  HCHECK(rowops_embed_pe(texts, raw_ptr(h, "txt_encoder.src_word_emb.weight"), pe, h->dims.vocab, lay, h->dims.d_model,
                         nullptr, nullptr, 0, out, st));
  return FS2_OK;
}

int fs2_op_fft_stack(fs2_handle* h, int32_t stack, int32_t l0, int32_t l1, int32_t prec, const float* x,
                     const int64_t* lens, int32_t B, int32_t S, float* out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h || !h->loaded) return FS2_ERR_STATE;
  std::vector<FftW>& Ls = stack == 0 ? h->w->enc : h->w->dec;
  if (!x || !lens || !out || B <= 0 || S <= 0 || l0 < 0 || l1 > (int)Ls.size() || l0 > l1 ||
      prec < FS2_PREC_FP32 || prec > FS2_PREC_F16X2)
    return h->fail(FS2_ERR_INVALID, "bad argument");
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = h->dims.d_model;
  WS(int, lens32, "op.lens32", B);
  HCHECK(rowops_lens_to_i32(lens, B, S, lens32, st));
  RowLayout lay;
  RCHECK(make_layout(h, "op.lay", lens32, B, S, h->halo_keep, FS2_HALO, &lay, st));
  const size_t R = (size_t)lay.R_cap;
  WS(float, xg, "op.x", R * D);
  bf16* xb = nullptr;
  if (prec != FS2_PREC_FP32) { WS(bf16, t, "op.xb", (size_t)planes_of(prec) * R * D); xb = t; }
  HCHECK(rowops_to_grid(x, lay, D, xg, D, 0, nullptr, st));
  HCHECK(make_shadow(xg, R * D, prec, xb, st));
  RCHECK(run_fft_stack(h, Ls, l0, l1, prec, xg, xb, lay, st));
  HCHECK(rowops_from_grid(xg, lay, D, out, st));
  return FS2_OK;
}

int fs2_op_variance_predictor(fs2_handle* h, int32_t which, const float* x, const int64_t* lens, int32_t B, int32_t S,
                              float* out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h || !h->loaded) return FS2_ERR_STATE;
  if (!x || !lens || !out || B <= 0 || S <= 0 || which < 0 || which > 2) return h->fail(FS2_ERR_INVALID, "bad argument");
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = h->dims.d_model;
  WS(int, lens32, "op.lens32", B);
  HCHECK(rowops_lens_to_i32(lens, B, S, lens32, st));
  RowLayout lay;
  RCHECK(make_layout(h, "op.lay", lens32, B, S, h->halo_keep, FS2_HALO, &lay, st));
  const size_t R = (size_t)lay.R_cap;
  WS(float, xg, "op.x", R * D);
  bf16* xb = nullptr;
  if (h->prec_enc != FS2_PREC_FP32) { WS(bf16, t, "op.xb", (size_t)planes_of(h->prec_enc) * R * D); xb = t; }
  HCHECK(rowops_to_grid(x, lay, D, xg, D, 0, nullptr, st));
  HCHECK(make_shadow(xg, R * D, h->prec_enc, xb, st));
  return run_predictor(h, h->w->pred[which], h->prec_enc, xg, xb, lay, out, st);
}

int fs2_op_variance_embed(fs2_handle* h, int32_t which, float* pred, float control, float* x, int32_t B, int32_t S,
                          int32_t* idx_out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h || !h->loaded) return FS2_ERR_STATE;
  if (!pred || !x || B <= 0 || S <= 0 || (which != 1 && which != 2)) return h->fail(FS2_ERR_INVALID, "bad argument");
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const char* nm = which == 1 ? "pitch" : "energy";
  RowLayout lay;   // dense user tensor: every row is a grid row, no halo
  RCHECK(make_layout(h, "op.lay", nullptr, B, S, S, 0, &lay, st));
  HCHECK(rowops_variance_embed(pred, control, raw_ptr(h, std::string("variance_adaptor.") + nm + "_bins"), h->dims.n_bins,
                               raw_ptr(h, std::string("variance_adaptor.") + nm + "_embedding.weight"), nullptr, x, nullptr,
                               0, lay, h->dims.d_model, idx_out, st));
  return FS2_OK;
}

int fs2_op_mel_postnet(fs2_handle* h, int32_t prec, const float* dec, int32_t B, int32_t T, float* mel, float* mel_post,
                       void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h || !h->loaded) return FS2_ERR_STATE;
  if (!dec || !mel || !mel_post || B <= 0 || T <= 0 || prec < FS2_PREC_FP32 || prec > FS2_PREC_F16X2)
    return h->fail(FS2_ERR_INVALID, "bad argument");
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = h->dims.d_model;
  RowLayout lay;   // all T rows of every utterance are given: the reference's padded grid
  RCHECK(make_layout(h, "op.lay", nullptr, B, T, T, FS2_HALO, &lay, st));
  const size_t R = (size_t)lay.R_cap;
  WS(float, xg, "op.x", R * D);
  bf16* xb = nullptr;
  if (prec != FS2_PREC_FP32) { WS(bf16, t, "op.xb", (size_t)planes_of(prec) * R * D); xb = t; }
  HCHECK(rowops_to_grid(dec, lay, D, xg, D, 0, nullptr, st));
  HCHECK(make_shadow(xg, R * D, prec, xb, st));
  return run_mel_postnet(h, prec, xg, xb, lay, mel, mel_post, st);
}

int fs2_op_mel_encoder(fs2_handle* h, int32_t prec, const float* src_seq, const float* mels, const int64_t* src_lens,
                       const int64_t* mel_lens, int32_t B, int32_t L, int32_t T, float* out, float* attn, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!h || !h->loaded) return FS2_ERR_STATE;
  if (!src_seq || !mels || !src_lens || !mel_lens || !out || B <= 0 || L <= 0 || T <= 0 || prec < FS2_PREC_FP32 ||
      prec > FS2_PREC_F16X2)
    return h->fail(FS2_ERR_INVALID, "mel_encoder: bad argument");
  if (L > 2048) return h->fail(FS2_ERR_UNSUPPORTED, "mel_encoder: more than 2048 phonemes (the scores of a query tile live in shared memory)");
  if (T > FS2_MAX_ROWS_PER_UTT) return h->fail(FS2_ERR_UNSUPPORTED, "mel_encoder: more than 65535 mel frames per utterance");
  if (!h->w->raw.count("mel_encoder.prenet.w_1.weight"))
    return h->fail(FS2_ERR_MISSING_WEIGHT, "mel_encoder.* was not part of the loaded state_dict");
  HCHECK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RCHECK(build_mel_encoder(h, st));
  g_fs2_plain_next = 1;
  return run_mel_encoder(h, prec, src_seq, mels, src_lens, mel_lens, B, L, T, out, attn, st);
}

int fs2_op_conv_gemm(int32_t prec, const float* A, const float* W, const float* bias, int32_t B, int32_t S, int32_t K,
                     int32_t N, int32_t taps, int32_t act, float* out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!A || !W || !bias || !out || B <= 0 || S <= 0 || K <= 0 || N <= 0 || taps < 1 || taps > 2 * FS2_HALO + 1 ||
      taps % 2 == 0 || K % 16 || N % 4 || act < 0 || act > 2 || prec < FS2_PREC_FP32 || prec > FS2_PREC_F16X2 ||
      B > 65535 || S > FS2_MAX_ROWS_PER_UTT) {
    g_last_error = "bad argument"; return FS2_ERR_INVALID; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float *Ag = nullptr, *Wf = nullptr, *Og = nullptr;
  bf16 *Ab = nullptr, *Wb = nullptr, *Wh = nullptr;
  int rc = FS2_OK;
  cudaError_t e = cudaSuccess;
  TmpLayout tl;
  do {
    if ((e = tl.build(nullptr, B, S, S, FS2_HALO, st)) != cudaSuccess) break;   // the padded grid itself
    const size_t R = (size_t)tl.lay.R_cap;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&Ag), sizeof(float) * R * K)) != cudaSuccess) break;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&Ab), sizeof(bf16) * 3 * R * K)) != cudaSuccess) break;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&Wf), sizeof(float) * (size_t)N * K * taps)) != cudaSuccess) break;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&Wb), sizeof(bf16) * 3 * (size_t)N * K * taps)) != cudaSuccess) break;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&Og), sizeof(float) * R * N)) != cudaSuccess) break;
    if ((e = rowops_to_grid(A, tl.lay, K, Ag, K, 0, nullptr, st)) != cudaSuccess) break;
    if ((e = make_shadow(Ag, R * K, prec, Ab, st)) != cudaSuccess) break;
    if ((e = rowops_pack_weight(W, N, K, taps, nullptr, Wf, Wb, N, 0, st)) != cudaSuccess) break;
    float w_scale = 1.f;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&Wh), sizeof(bf16) * 2 * (size_t)N * K * taps)) != cudaSuccess) break;
    if ((e = rowops_pack_weight_f16x2(Wf, N, K, taps, Wh, &w_scale, st)) != cudaSuccess) break;
    ConvGemmArgs a;
    memset(&a, 0, sizeof a);
    a.A = Ag; a.Ab = Ab; a.K = K; a.Wf = Wf; a.Wb = Wb; a.Wh = Wh; a.bias = bias; a.N = N; a.taps = taps;
    a.acc_scale = 1.f; a.acc_scale_f16x2 = 1.f / (FS2_F16X2_ACT_SCALE * w_scale);
    a.lay = tl.lay; a.epi = act == 0 ? EPI_BIAS : act == 1 ? EPI_RELU : EPI_TANH; a.mask_mode = MASK_GRID;
    a.out = Og; a.ldo = N; a.out_user = out; a.ldu = N;
    rc = run_gemm(nullptr, prec, a, st);
    if (rc != FS2_OK) break;
    e = cudaStreamSynchronize(st);
  } while (0);
  cudaFree(Ag); cudaFree(Ab); cudaFree(Wf); cudaFree(Wb); cudaFree(Wh); cudaFree(Og);
  if (e != cudaSuccess) return fs2_fail_cuda(e, "fs2_op_conv_gemm");
  return rc;
}

int fs2_op_attention(int32_t prec, const float* q, const float* k, const float* v, const int64_t* lens, int32_t B,
                     int32_t S, int32_t H, int32_t dk, float* out, void* stream) {
  g_fs2_plain_next = 1;   // the first launch of an entry point is fully stream-ordered (fs2_common.cuh)
  if (!q || !k || !v || !lens || !out || B <= 0 || S <= 0 || H <= 0 || (dk != 64 && dk != 128) || B > 65535 ||
      S > FS2_MAX_ROWS_PER_UTT) {
    g_last_error = "bad argument"; return FS2_ERR_INVALID; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = H * dk;
  float *qkv = nullptr, *og = nullptr;
  bf16 *qb = nullptr, *kb = nullptr, *vb = nullptr, *vtb = nullptr, *ob = nullptr;
  int* lens32 = nullptr;
  int rc = FS2_OK;
  cudaError_t e = cudaSuccess;
  TmpLayout tl;
  do {
    if ((e = cudaMalloc(reinterpret_cast<void**>(&lens32), sizeof(int) * B)) != cudaSuccess) break;
    if ((e = rowops_lens_to_i32(lens, B, S, lens32, st)) != cudaSuccess) break;
    if ((e = tl.build(lens32, B, S, 2, FS2_HALO, st)) != cudaSuccess) break;   // packed rows, as on the product path
    const size_t R = (size_t)tl.lay.R_cap;
    const int Rv = (tl.lay.R_cap + 7) & ~7;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&og), sizeof(float) * R * D)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(og, 0, sizeof(float) * R * D, st)) != cudaSuccess) break;
    g_fs2_plain_next = 1;
    if (prec == FS2_PREC_FP32) {
      if ((e = cudaMalloc(reinterpret_cast<void**>(&qkv), sizeof(float) * R * 3 * D)) != cudaSuccess) break;
      if ((e = rowops_to_grid(q, tl.lay, D, qkv, 3 * D, 0, nullptr, st)) != cudaSuccess) break;
      if ((e = rowops_to_grid(k, tl.lay, D, qkv, 3 * D, D, nullptr, st)) != cudaSuccess) break;
      if ((e = rowops_to_grid(v, tl.lay, D, qkv, 3 * D, 2 * D, nullptr, st)) != cudaSuccess) break;
      if ((e = simt_attention_launch(qkv, 3 * D, 0, D, 2 * D, tl.lay, H, dk, og, D, st)) != cudaSuccess) break;
    } else {
      if (D != 256 || dk != 128) { g_last_error = "tcgen05 attention is built for H*dk = 256, dk = 128"; rc = FS2_ERR_UNSUPPORTED; break; }
      const size_t np = prec == FS2_PREC_F16X2 ? 2 : 1;   // bf16x3 has no tensor-core attention: tested as bf16
      if ((e = cudaMalloc(reinterpret_cast<void**>(&qb), sizeof(bf16) * np * R * D)) != cudaSuccess) break;
      if ((e = cudaMalloc(reinterpret_cast<void**>(&kb), sizeof(bf16) * np * R * D)) != cudaSuccess) break;
      if ((e = cudaMalloc(reinterpret_cast<void**>(&vb), sizeof(bf16) * np * R * D)) != cudaSuccess) break;
      if ((e = cudaMalloc(reinterpret_cast<void**>(&vtb), sizeof(bf16) * np * (size_t)D * Rv)) != cudaSuccess) break;
      if ((e = cudaMalloc(reinterpret_cast<void**>(&ob), sizeof(bf16) * np * R * D)) != cudaSuccess) break;
      if ((e = cudaMemsetAsync(ob, 0, sizeof(bf16) * np * R * D, st)) != cudaSuccess) break;
      g_fs2_plain_next = 1;
      if (np == 1) {
        if ((e = rowops_to_grid(q, tl.lay, D, nullptr, 0, 0, qb, st)) != cudaSuccess) break;
        if ((e = rowops_to_grid(k, tl.lay, D, nullptr, 0, 0, kb, st)) != cudaSuccess) break;
        if ((e = rowops_to_grid(v, tl.lay, D, nullptr, 0, 0, vb, st)) != cudaSuccess) break;
      } else {   // fp32 grid copy (in og) -> two scaled fp16 planes
        const float* src3[3] = {q, k, v};
        bf16* dst3[3] = {qb, kb, vb};
        bool bad = false;
        for (int i = 0; i < 3 && !bad; ++i) {
          if ((e = rowops_to_grid(src3[i], tl.lay, D, og, D, 0, nullptr, st)) != cudaSuccess) { bad = true; break; }
          if ((e = rowops_split(og, (int64_t)(R * D), 2, dst3[i], (int64_t)(R * D), st)) != cudaSuccess) { bad = true; break; }
        }
        if (bad) break;
      }
      for (size_t pl = 0; pl < np && e == cudaSuccess; ++pl)
        e = rowops_transpose_v(vb + pl * R * D, tl.lay.R_cap, Rv, D, vtb + pl * (size_t)D * Rv, st);
      if (e != cudaSuccess) break;
      rc = tc_attention_launch(qb, kb, vtb, tl.lay, Rv, H, (int)np, ob, st);
      if (rc != FS2_OK) break;
      if (np == 1) e = rowops_bf16_to_f32(ob, (int64_t)(R * D), og, st);
      else e = rowops_unsplit2(ob, (int64_t)(R * D), (int64_t)(R * D), og, st);
      if (e != cudaSuccess) break;
    }
    if ((e = rowops_from_grid(og, tl.lay, D, out, st)) != cudaSuccess) break;
    e = cudaStreamSynchronize(st);
  } while (0);
  cudaFree(qkv); cudaFree(og); cudaFree(qb); cudaFree(kb); cudaFree(vb); cudaFree(vtb); cudaFree(ob); cudaFree(lens32);
  if (e != cudaSuccess) return fs2_fail_cuda(e, "fs2_op_attention");
  return rc;
}

}  // extern "C"
