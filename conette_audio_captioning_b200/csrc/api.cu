// C-ABI of libconette_b200.so (declared in include/conette_b200.h): handle, weight packing, workspace and the
// orchestration of the per-stage kernels.  No torch types anywhere; PyTorch only owns the caller's buffers.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/conette_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace cnb {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

constexpr int kDims[4] = {96, 192, 384, 768};
constexpr int kDepths[4] = {3, 3, 9, 3};
constexpr int kStageW[4] = {56, 28, 14, 7};
constexpr int kMels = 224, kBins = 513, kNfft = 1024, kHop = 320;
constexpr int kD = 256, kFF = 2048, kLayers = 6, kTags = 527;
constexpr int kDefaultMlpSlabMB = 0;  // see encode_chunk: MLP row slabs sized to keep the hidden activations in L2
constexpr int kDefaultChunk = 64;  // measured on B200: larger encoder passes amortise tails; 64 x 10 s needs ~3.5 GB

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

// simple bump arena for packed weights (256-byte aligned)
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  template <typename T> T* take(size_t n) {
    used = (used + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + used);
    used += n * sizeof(T);
    return p;
  }
};

struct BlockW {
  float *dw_w_t, *dw_b, *ln_g, *ln_b, *b1, *b2, *scale, *w1, *w2;
  act16 *w1_bf, *w2_bf;
};
struct DownW {
  float *ln_g, *ln_b, *bias, *w;
  act16* w_bf;
};
struct LayerW {
  float *sa_in_w, *sa_in_b, *sa_out_w, *sa_out_b, *ca_q_w, *ca_q_b, *ca_out_w, *ca_out_b;
  float *l1_w, *l1_b, *l2_w, *l2_b, *n1_g, *n1_b, *n2_g, *n2_b, *n3_g, *n3_b;
};

struct Buffer {
  void* ptr = nullptr;
  size_t bytes = 0;
};
struct DecGraph {
  std::vector<int64_t> key;
  cudaGraphExec_t exec;
  int64_t n_kernels;
};

}  // namespace cnb

using namespace cnb;

struct cnb_handle {
  cnb_config cfg;
  std::map<std::string, HostTensor> staged;
  bool finalized = false;
  bool has_encoder = true;  // false: only the projection + decoder were loaded (BaselinePLM, precomputed embeddings)
  Arena arena;
  // encoder
  FrontendParams fe;
  float *stem_w_t, *stem_b, *stem_ln_g, *stem_ln_b;
  BlockW blocks[18];
  DownW down[3];
  float *head_ln_g, *head_ln_b, *head_w, *head_b;
  // projection + decoder
  float *proj_w, *proj_b, *emb, *pe, *ca_kv_w, *ca_kv_b, *cls_w, *cls_b;
  act16 *proj_w3 = nullptr, *ca_kv_w3 = nullptr;  // [W1 | W1 | W2] rows for the split-precision tensor-core projection (dec_project)
  void* dec_tmaps = nullptr;  // kDecMaps TMA descriptors of the fp16-split decoder weights (cluster decoder)
  LayerW layers[6];
  // workspace (grown on demand)
  std::map<std::string, Buffer> ws;
  size_t ws_bytes = 0;
  int* zero_flag = nullptr;  // device int[4] that stays 0: "done" flag for non-beam callers
  // CUDA-graph replay of the decode loop
  int use_cluster = 1;  // 0 never, 1 when the shape allows it (default), 2 required
  bool use_graphs = true;
  bool dec_compact = false;  // the decode about to be enqueued shares the GPU with the next batch's encoder (streaming API)
  cudaStream_t stream = nullptr;  // library-owned non-blocking stream (graph capture / replay, host-API copies)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<DecGraph> dec_graphs;
  // host-buffer path (cnb_caption_host): sliced H2D on a copy stream overlapping the front-end + stem of earlier slices
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_slice[8] = {};
  cudaStream_t dec_stream = nullptr;     // high-priority stream of the decode half when two batches are in flight
  cudaEvent_t ev_enc[2] = {};            // encoder of the batch in slot i has finished (decode may start)
  cudaEvent_t host_done[2] = {};         // cnb_caption_host_begin/end: completion of the batch that used staging slot i
  bool host_pending[2] = {false, false};
  int host_next = 0;
  cudaEvent_t ev_dev_last = nullptr;     // end of the last device-buffer entry point (cnb_caption / encoder / decode ...)
  bool dev_last_valid = false;
  cudaEvent_t ev_lens = nullptr;         // the last H2D copy out of pin_lens
  int32_t* pin_lens = nullptr;           // pinned staging for the per-clip frame counts
  int pin_lens_cap = 0;
  int enc_narrow_blocks = 0;             // streaming: the first ConvNeXt blocks of this encode share the GPU with a decoder ...
  int enc_narrow_sms = 148;              // ... that leaves this many SMs (set_sm_budget)
  bool pre_stem_done = false;            // encode_chunk: the "logmel" / "xa" workspaces already hold this chunk's stem output
  // optional per-kernel-class timing (cnb_profile_begin/end): CUDA event pairs around every launch
  bool prof_on = false;
  bool prof_timeline = false;            // cnb_profile_timeline_begin: keep the streaming overlap while bracketing (see _end)
  std::vector<cudaEvent_t> prof_events;  // pool, pairs (start, stop)
  std::vector<int> prof_class;           // class of pair i
  size_t prof_used = 0;                  // events used
};

namespace cnb {

// RAII bracket: records an event pair on `st` around the launches issued while it is alive (only in profile mode)
struct Prof {
  cnb_handle* h;
  cudaStream_t st;
  bool on;
  Prof(cnb_handle* h_, int cls, cudaStream_t st_) : h(h_), st(st_), on(h_->prof_on) {
    if (!on) return;
    if (h->prof_used + 2 > h->prof_events.size()) {
      const size_t old = h->prof_events.size();
      h->prof_events.resize(old + 4096);
      for (size_t i = old; i < h->prof_events.size(); ++i) cudaEventCreate(&h->prof_events[i]);
    }
    h->prof_class.push_back(cls);
    cudaEventRecord(h->prof_events[h->prof_used], st);
  }
  ~Prof() {
    if (!on) return;
    cudaEventRecord(h->prof_events[h->prof_used + 1], st);
    h->prof_used += 2;
  }
};

static int ws_get(cnb_handle* h, const char* name, size_t bytes, void** out) {
  Buffer& b = h->ws[name];
  if (b.bytes < bytes) {
    if (b.ptr) {
      CNB_CUDA_OK(cudaDeviceSynchronize());
      CNB_CUDA_OK(cudaFree(b.ptr));
      h->ws_bytes -= b.bytes;
      b.ptr = nullptr;
      b.bytes = 0;
    }
    // any (re)allocation invalidates the buffer addresses baked into cached CUDA graphs
    for (auto& g : h->dec_graphs) cudaGraphExecDestroy(g.exec);
    h->dec_graphs.clear();
    const size_t want = (bytes + 255) & ~size_t(255);
    CNB_CUDA_OK(cudaMalloc(&b.ptr, want));
    CNB_CUDA_OK(cudaMemset(b.ptr, 0, want));
    CNB_CUDA_OK(cudaDeviceSynchronize());  // the memset runs on the legacy stream; our streams are non-blocking
    b.bytes = want;
    h->ws_bytes += want;
  }
  *out = b.ptr;
  return 0;
}
#define WS(h, name, type, count, var)                                                     \
  type* var = nullptr;                                                                    \
  do {                                                                                    \
    void* _p = nullptr;                                                                   \
    if (int _rc = ws_get(h, name, sizeof(type) * (size_t)(count), &_p)) return _rc;       \
    var = reinterpret_cast<type*>(_p);                                                    \
  } while (0)

struct Geometry {
  int t, h[4], tp;
};
static Geometry geometry(int64_t n) {
  Geometry g;
  g.t = (int)(n / kHop) + 1;
  g.h[0] = (g.t + 8 - 4) / 4 + 1;  // Conv2d k=4 s=4 pad=(4,0)  (reference convnext.py:405-408)
  g.h[1] = g.h[0] / 2;
  g.h[2] = g.h[1] / 2;
  g.h[3] = g.h[2] / 2;
  g.tp = g.h[3];
  return g;
}

// ---------------------------------------------------------------------------------------------------------------------
// weight staging helpers
// ---------------------------------------------------------------------------------------------------------------------
static const HostTensor* find(cnb_handle* h, const std::string& name, std::initializer_list<int64_t> shape) {
  auto it = h->staged.find(name);
  if (it == h->staged.end()) {
    set_error("missing weight: " + name);
    return nullptr;
  }
  if (it->second.shape != std::vector<int64_t>(shape)) {
    std::string s = "[";
    for (auto d : it->second.shape) s += std::to_string(d) + ",";
    set_error("unexpected shape for " + name + ": " + s + "]");
    return nullptr;
  }
  return &it->second;
}

template <typename T> static int upload(T* dst, const std::vector<T>& src) {
  CNB_CUDA_OK(cudaMemcpy(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
static int put_f32(cnb_handle* h, const std::vector<float>& v, float** out) {
  *out = h->arena.take<float>(v.size());
  return upload(*out, v);
}
static int put_f16(cnb_handle* h, const std::vector<float>& v, act16** out) {
  std::vector<act16> t(v.size());
  for (size_t i = 0; i < v.size(); ++i) t[i] = float2act(v[i]);
  *out = h->arena.take<act16>(v.size());
  CNB_CUDA_OK(cudaMemcpy(*out, t.data(), t.size() * sizeof(act16), cudaMemcpyHostToDevice));
  return 0;
}
#define GET(var, name, ...)                                  \
  const HostTensor* var = find(h, name, {__VA_ARGS__});      \
  if (!var) return -4
#define PUT(dst, vec) \
  if (int _rc = put_f32(h, vec, &(dst))) return _rc
#define PUT_BF(dst, vec) \
  if (int _rc = put_f16(h, vec, &(dst))) return _rc
// an already converted fp16 vector
#define PUT_BF16V(dst, vec)                                                                          \
  do {                                                                                               \
    const std::vector<act16> _v = (vec);                                                             \
    (dst) = h->arena.take<act16>(_v.size());                                                         \
    CNB_CUDA_OK(cudaMemcpy((dst), _v.data(), _v.size() * sizeof(act16), cudaMemcpyHostToDevice));    \
  } while (0)

// fp16 hi/lo split of a decoder weight matrix for the cluster decoder (decoder_cluster.cu): W = W1 + 2^-11 W2 with
// W1 = fp16(W), W2 = fp16((W - W1) * 2048); `out` = [W1 (n*k) | W2 (n*k)].  `head_pack`: columns regrouped per attention head,
// out row (h * n + r) = W[r][32 h .. 32 h + 31] (the K-split operand of the attention output projections).
static std::vector<act16> split_f16(const std::vector<float>& w, int n, int k, bool head_pack) {
  const size_t sz = (size_t)n * k;
  std::vector<act16> out(2 * sz);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < k; ++c) {
      const float v = w[(size_t)r * k + c];
      const act16 h1 = float2act(v);
      const act16 h2 = float2act((v - act2float(h1)) * 2048.f);
      const size_t o = head_pack ? ((size_t)(c / 32) * n + r) * 32 + (c % 32) : (size_t)r * k + c;
      out[o] = h1;
      out[sz + o] = h2;
    }
  return out;
}

// Weights of a GEMM that runs on the tensor cores at fp32-level accuracy through ONE fp16 GEMM of triple depth (dec_project):
// row r = [W1 | W1 | W2] with W1 = fp16(W), W2 = fp16(W - W1), against activation rows [A1 | A2 | A1]:
// A1 W1 + A2 W1 + A1 W2 = A W - A2 W2.  The low parts are stored unscaled: they fall into fp16's subnormal range, whose fixed
// 2^-24 step still resolves them to 2^-13 of the high part's unit in the last place (|W| < 0.25, |A| < 64 here), i.e. ~1e-6
// relative on the product sums -- the level of fp32 accumulation order effects.
static std::vector<act16> concat3_f16(const std::vector<float>& w, int n, int k) {
  std::vector<act16> out((size_t)n * 3 * k);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < k; ++c) {
      const float v = w[(size_t)r * k + c];
      const act16 h1 = float2act(v);
      act16* o = &out[(size_t)r * 3 * k];
      o[c] = h1;
      o[k + c] = h1;
      o[2 * k + c] = float2act(v - act2float(h1));
    }
  return out;
}

static int finalize(cnb_handle* h) {
  const std::string E = "preprocessor.encoder.", M = "model.", D = "model.decoder.";
  const int V = h->cfg.vocab_size;
  CNB_CUDA_OK(cudaSetDevice(h->cfg.device));
  if (int rc = gemm_tc_init()) return rc;

  size_t total = 0;
  for (auto& kv : h->staged) total += kv.second.data.size();
  h->arena.cap = total * 10 + (64u << 20);  // f32 + fp16 copies + fp16-split decoder copies + slack
  CNB_CUDA_OK(cudaMalloc(&h->arena.base, h->arena.cap));
  CNB_CUDA_OK(cudaMemset(h->arena.base, 0, h->arena.cap));

  // A state dict without the audio encoder (reference BaselinePLM, pl_modules/baseline.py:35: FrameIdentEncoder + projection +
  // decoder on precomputed frame embeddings) loads the decoder half only; the waveform entry points then refuse to run.
  h->has_encoder = h->staged.count(E + "bn0.weight") != 0;
  // ---- front-end: analytic twiddles; the checkpoint's DFT basis must be the Hann-windowed DFT
  if (h->has_encoder) {
    GET(cr, E + "spectrogram_extractor.stft.conv_real.weight", kBins, 1, kNfft);
    GET(ci, E + "spectrogram_extractor.stft.conv_imag.weight", kBins, 1, kNfft);
    double max_err = 0;
    for (int k : {0, 1, 7, 100, 255, 256, 511, 512})
      for (int n = 0; n < kNfft; n += 3) {
        const double w = 0.5 - 0.5 * std::cos(2.0 * M_PI * n / kNfft);
        const double ang = 2.0 * M_PI * (double)((int64_t)k * n % kNfft) / kNfft;
        max_err = std::max(max_err, std::fabs(cr->data[(size_t)k * kNfft + n] - w * std::cos(ang)));
        max_err = std::max(max_err, std::fabs(ci->data[(size_t)k * kNfft + n] + w * std::sin(ang)));
      }
    if (max_err > 1e-5) {
      set_error("spectrogram_extractor.stft.conv_{real,imag}.weight is not the periodic-Hann DFT basis (max err " +
                std::to_string(max_err) + "); the FFT front-end cannot represent it");
      return -4;
    }
    std::vector<float> tw(2 * kNfft);
    for (int j = 0; j < kNfft; ++j) {
      tw[2 * j] = (float)std::cos(2.0 * M_PI * j / kNfft);
      tw[2 * j + 1] = (float)(-std::sin(2.0 * M_PI * j / kNfft));
    }
    float* twd;
    PUT(twd, tw);
    h->fe.twiddle = reinterpret_cast<const float2*>(twd);
    // sparse mel: per filter the contiguous range of non-zero FFT bins (melW is data, never regenerated)
    GET(mel, E + "logmel_extractor.melW", kBins, kMels);
    std::vector<int> lo(kMels), cnt(kMels), off(kMels);
    std::vector<float> wts;
    for (int m = 0; m < kMels; ++m) {
      int first = -1, last = -1;
      for (int k = 0; k < kBins; ++k)
        if (mel->data[(size_t)k * kMels + m] != 0.f) {
          if (first < 0) first = k;
          last = k;
        }
      lo[m] = first < 0 ? 0 : first;
      cnt[m] = first < 0 ? 0 : last - first + 1;
      off[m] = (int)wts.size();
      for (int k = 0; k < cnt[m]; ++k) wts.push_back(mel->data[(size_t)(lo[m] + k) * kMels + m]);
    }
    if (wts.empty()) wts.push_back(0.f);
    int *dlo = h->arena.take<int>(kMels), *dcnt = h->arena.take<int>(kMels), *doff = h->arena.take<int>(kMels);
    if (int rc = upload(dlo, lo)) return rc;
    if (int rc = upload(dcnt, cnt)) return rc;
    if (int rc = upload(doff, off)) return rc;
    float* dw;
    PUT(dw, wts);
    h->fe.mel_lo = dlo; h->fe.mel_cnt = dcnt; h->fe.mel_off = doff; h->fe.mel_w = dw;
    {
      // Lane-balanced schedule for the warp-per-frame-pair kernel: the filters are dealt to the 32 lanes longest first, each
      // to the lane with the least work so far; a lane walks its filters as ONE flat list of (weight, bin | end | filter)
      // entries (bins in ascending order, so the sum order of every filter is unchanged), padded with zero weights to the
      // longest lane.  Entry i of lane l sits at [i][l].
      std::vector<int> order(kMels), load(32, 0);
      for (int m = 0; m < kMels; ++m) order[m] = m;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
      std::vector<std::vector<int>> mine(32);
      for (int m : order) {
        int best = 0;
        for (int l = 1; l < 32; ++l)
          if (load[l] < load[best]) best = l;
        mine[best].push_back(m);
        load[best] += std::max(cnt[m], 1);
      }
      int len = 0;
      for (int l = 0; l < 32; ++l) len = std::max(len, load[l]);
      if (len > kMelSchedMax) {
        set_error("logmel_extractor.melW: a lane of the balanced mel schedule needs " + std::to_string(len) + " entries (max " +
                  std::to_string(kMelSchedMax) + ")");
        return -4;
      }
      std::vector<float> sched((size_t)len * 64, 0.f);
      for (int l = 0; l < 32; ++l) {
        int i = 0;
        for (int m : mine[l]) {
          const int n = std::max(cnt[m], 1);   // an all-zero filter still emits one (zero-weight) entry that ends it
          for (int k = 0; k < n; ++k, ++i) {
            const float w = cnt[m] ? wts[(size_t)off[m] + k] : 0.f;
            const int code = (lo[m] + k) | (k == n - 1 ? 1 << 10 : 0) | (m << 11);
            float cf;
            std::memcpy(&cf, &code, 4);
            sched[((size_t)i * 32 + l) * 2] = w;
            sched[((size_t)i * 32 + l) * 2 + 1] = cf;
          }
        }
      }
      float* dsched;
      PUT(dsched, sched);
      h->fe.mel_sched = reinterpret_cast<const float2*>(dsched);
      h->fe.mel_sched_len = len;
    }
    GET(g, E + "bn0.weight", kMels);
    GET(b, E + "bn0.bias", kMels);
    GET(mu, E + "bn0.running_mean", kMels);
    GET(var, E + "bn0.running_var", kMels);
    std::vector<float> sc(kMels), sh(kMels), ones(kMels, 1.f), zeros(kMels, 0.f);
    for (int m = 0; m < kMels; ++m) {
      sc[m] = g->data[m] / std::sqrt(var->data[m] + 1e-5f);
      sh[m] = b->data[m] - mu->data[m] * sc[m];
    }
    float *dsc, *dsh, *dones, *dzeros;
    PUT(dsc, sc); PUT(dsh, sh); PUT(dones, ones); PUT(dzeros, zeros);
    h->fe.bn_scale = dsc; h->fe.bn_shift = dsh; h->fe.ones = dones; h->fe.zeros = dzeros;
  }
  // ---- stem
  if (h->has_encoder) {
    GET(w, E + "downsample_layers.0.0.weight", 96, 1, 4, 4);
    GET(b, E + "downsample_layers.0.0.bias", 96);
    GET(g, E + "downsample_layers.0.1.weight", 96);
    GET(be, E + "downsample_layers.0.1.bias", 96);
    std::vector<float> wt(16 * 96);
    for (int c = 0; c < 96; ++c)
      for (int i = 0; i < 16; ++i) wt[i * 96 + c] = w->data[c * 16 + i];
    PUT(h->stem_w_t, wt); PUT(h->stem_b, b->data); PUT(h->stem_ln_g, g->data); PUT(h->stem_ln_b, be->data);
  }
  // ---- downsample layers 1..3: LN(cf) + Conv2d(k=2,s=2); weight reordered to (Cout, kh, kw, Cin)
  for (int i = 1; i < 4 && h->has_encoder; ++i) {
    const int cin = kDims[i - 1], cout = kDims[i];
    const std::string p = E + "downsample_layers." + std::to_string(i) + ".";
    GET(g, p + "0.weight", cin);
    GET(be, p + "0.bias", cin);
    GET(w, p + "1.weight", cout, cin, 2, 2);
    GET(b, p + "1.bias", cout);
    std::vector<float> wr((size_t)cout * 4 * cin);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int kh = 0; kh < 2; ++kh)
          for (int kw = 0; kw < 2; ++kw)
            wr[((size_t)o * 4 + kh * 2 + kw) * cin + c] = w->data[(((size_t)o * cin + c) * 2 + kh) * 2 + kw];
    DownW& d = h->down[i - 1];
    PUT(d.ln_g, g->data); PUT(d.ln_b, be->data); PUT(d.bias, b->data); PUT(d.w, wr); PUT_BF(d.w_bf, wr);
  }
  // ---- blocks
  if (h->has_encoder) {
    int bi = 0;
    for (int s = 0; s < 4; ++s)
      for (int j = 0; j < kDepths[s]; ++j, ++bi) {
        const int c = kDims[s];
        const std::string p = E + "stages." + std::to_string(s) + "." + std::to_string(j) + ".";
        GET(sc, p + "scale_layer", c);
        GET(dw, p + "dwconv.weight", c, 1, 7, 7);
        GET(db, p + "dwconv.bias", c);
        GET(g, p + "norm.weight", c);
        GET(be, p + "norm.bias", c);
        GET(w1, p + "pwconv1.weight", 4 * c, c);
        GET(b1, p + "pwconv1.bias", 4 * c);
        GET(w2, p + "pwconv2.weight", c, 4 * c);
        GET(b2, p + "pwconv2.bias", c);
        std::vector<float> wt((size_t)49 * c);
        for (int ch = 0; ch < c; ++ch)
          for (int t = 0; t < 49; ++t) wt[(size_t)t * c + ch] = dw->data[(size_t)ch * 49 + t];
        BlockW& b = h->blocks[bi];
        PUT(b.dw_w_t, wt); PUT(b.dw_b, db->data); PUT(b.ln_g, g->data); PUT(b.ln_b, be->data);
        PUT(b.b1, b1->data); PUT(b.b2, b2->data); PUT(b.scale, sc->data);
        PUT(b.w1, w1->data); PUT(b.w2, w2->data); PUT_BF(b.w1_bf, w1->data); PUT_BF(b.w2_bf, w2->data);
      }
  }
  // ---- clip head
  if (h->has_encoder) {
    GET(g, E + "norm.weight", 768);
    GET(be, E + "norm.bias", 768);
    GET(w, E + "head_audioset.weight", kTags, 768);
    GET(b, E + "head_audioset.bias", kTags);
    PUT(h->head_ln_g, g->data); PUT(h->head_ln_b, be->data); PUT(h->head_w, w->data); PUT(h->head_b, b->data);
  }
  // ---- projection + decoder
  {
    GET(pw, M + "projection.2.weight", kD, 768);
    GET(pb, M + "projection.2.bias", kD);
    GET(emb, D + "emb_layer.weight", V, kD);
    GET(cw, D + "classifier.weight", V, kD);
    GET(cb, D + "classifier.bias", V);
    PUT(h->proj_w, pw->data); PUT(h->proj_b, pb->data); PUT(h->emb, emb->data); PUT(h->cls_w, cw->data);
    PUT_BF16V(h->proj_w3, concat3_f16(pw->data, kD, 768));
    PUT(h->cls_b, cb->data);
    // fp16-split decoder weights + their TMA descriptors (cluster decoder): map index 12 l + 2 j + half, classifier 72 + half
    std::vector<uint8_t> maps((size_t)kDecMaps * 128);
    auto put_split = [&](int idx, const std::vector<float>& w, int n, int k, bool head_pack, int box_rows, int box_cols) -> int {
      const std::vector<act16> sp = split_f16(w, n, k, head_pack);
      act16* d = h->arena.take<act16>(sp.size());
      CNB_CUDA_OK(cudaMemcpy(d, sp.data(), sp.size() * sizeof(act16), cudaMemcpyHostToDevice));
      const int64_t rows = head_pack ? (int64_t)n * (k / 32) : n, cols = head_pack ? 32 : k;
      for (int half = 0; half < 2; ++half) {
        alignas(64) uint8_t tmp[128];
        if (int rc = tc_make_map_f16_box(tmp, d + (size_t)half * n * k, rows, cols, box_rows, box_cols)) return rc;
        memcpy(&maps[(size_t)(idx + half) * 128], tmp, 128);
      }
      return 0;
    };
    if (int rc = put_split(kLayers * kDecMapsPerLayer, cw->data, V, kD, false, 128, 64)) return rc;
    auto it = h->staged.find(D + "pos_encoding.pos_embedding");
    if (it == h->staged.end() || it->second.shape.size() != 3 || it->second.shape[2] != kD) {
      set_error("missing or malformed weight: " + D + "pos_encoding.pos_embedding");
      return -4;
    }
    PUT(h->pe, it->second.data);
    std::vector<float> kvw((size_t)kLayers * 2 * kD * kD), kvb((size_t)kLayers * 2 * kD);
    for (int l = 0; l < kLayers; ++l) {
      const std::string p = D + "layers." + std::to_string(l) + ".";
      LayerW& L = h->layers[l];
      GET(saw, p + "self_attn.in_proj_weight", 3 * kD, kD);
      GET(sab, p + "self_attn.in_proj_bias", 3 * kD);
      GET(sow, p + "self_attn.out_proj.weight", kD, kD);
      GET(sob, p + "self_attn.out_proj.bias", kD);
      GET(caw, p + "multihead_attn.in_proj_weight", 3 * kD, kD);
      GET(cab, p + "multihead_attn.in_proj_bias", 3 * kD);
      GET(cow, p + "multihead_attn.out_proj.weight", kD, kD);
      GET(cob, p + "multihead_attn.out_proj.bias", kD);
      GET(l1w, p + "linear1.weight", kFF, kD);
      GET(l1b, p + "linear1.bias", kFF);
      GET(l2w, p + "linear2.weight", kD, kFF);
      GET(l2b, p + "linear2.bias", kD);
      PUT(L.sa_in_w, saw->data); PUT(L.sa_in_b, sab->data); PUT(L.sa_out_w, sow->data); PUT(L.sa_out_b, sob->data);
      std::vector<float> qw(caw->data.begin(), caw->data.begin() + (size_t)kD * kD);
      std::vector<float> qb(cab->data.begin(), cab->data.begin() + kD);
      PUT(L.ca_q_w, qw); PUT(L.ca_q_b, qb); PUT(L.ca_out_w, cow->data); PUT(L.ca_out_b, cob->data);
      std::copy(caw->data.begin() + (size_t)kD * kD, caw->data.end(), kvw.begin() + (size_t)l * 2 * kD * kD);
      std::copy(cab->data.begin() + kD, cab->data.end(), kvb.begin() + (size_t)l * 2 * kD);
      PUT(L.l1_w, l1w->data); PUT(L.l1_b, l1b->data); PUT(L.l2_w, l2w->data); PUT(L.l2_b, l2b->data);
      const int mb = kDecMapsPerLayer * l;
      if (int rc = put_split(mb + 0, saw->data, 3 * kD, kD, false, 32, 64)) return rc;
      if (int rc = put_split(mb + 2, sow->data, kD, kD, true, 128, 32)) return rc;
      if (int rc = put_split(mb + 4, qw, kD, kD, false, 32, 64)) return rc;
      if (int rc = put_split(mb + 6, cow->data, kD, kD, true, 128, 32)) return rc;
      if (int rc = put_split(mb + 8, l1w->data, kFF, kD, false, 128, 64)) return rc;
      if (int rc = put_split(mb + 10, l2w->data, kD, kFF, false, 128, 64)) return rc;
      float** gs[3] = {&L.n1_g, &L.n2_g, &L.n3_g};
      float** bs[3] = {&L.n1_b, &L.n2_b, &L.n3_b};
      for (int n = 0; n < 3; ++n) {
        GET(g, p + "norm" + std::to_string(n + 1) + ".weight", kD);
        GET(b, p + "norm" + std::to_string(n + 1) + ".bias", kD);
        PUT(*gs[n], g->data); PUT(*bs[n], b->data);
      }
    }
    PUT(h->ca_kv_w, kvw); PUT(h->ca_kv_b, kvb);
    PUT_BF16V(h->ca_kv_w3, concat3_f16(kvw, kLayers * 2 * kD, kD));
    CNB_CUDA_OK(cudaMalloc(&h->dec_tmaps, maps.size()));
    CNB_CUDA_OK(cudaMemcpy(h->dec_tmaps, maps.data(), maps.size(), cudaMemcpyHostToDevice));
  }
  CNB_CUDA_OK(cudaMalloc(&h->zero_flag, 4 * sizeof(int)));
  CNB_CUDA_OK(cudaMemset(h->zero_flag, 0, 4 * sizeof(int)));
  h->staged.clear();
  h->finalized = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// encoder orchestration
// ---------------------------------------------------------------------------------------------------------------------
struct Tap {
  int kind = -1, stage = 0, block = 0;
  float* out = nullptr;
  int64_t cap = 0;
  bool hit = false;
};

static int copy_tap(Tap* tap, const void* src, int64_t elems, bool src_is_f16, cudaStream_t st);

template <typename ActT>
static int mlp_gemm(cnb_handle* h, const ActT* a, const float* w32, const act16* wbf, int m, int n, int k, Epilogue epi,
                    const EpiParams& ep, void* out, bool out_act, int64_t ldo, cudaStream_t st);

template <>
int mlp_gemm<float>(cnb_handle* h, const float* a, const float* w32, const act16*, int m, int n, int k, Epilogue epi,
                    const EpiParams& ep, void* out, bool, int64_t ldo, cudaStream_t st) {
  return launch_gemm_f32<float>(a, k, w32, m, n, k, epi, ep, reinterpret_cast<float*>(out), ldo, st);
}
template <>
int mlp_gemm<act16>(cnb_handle* h, const act16* a, const float*, const act16* wbf, int m, int n, int k,
                            Epilogue epi, const EpiParams& ep, void* out, bool out_act, int64_t ldo, cudaStream_t st) {
  if (out_act)
    return launch_gemm_tc<act16>(a, wbf, m, n, k, epi, ep, reinterpret_cast<act16*>(out), ldo, st);
  return launch_gemm_tc<float>(a, wbf, m, n, k, epi, ep, reinterpret_cast<float*>(out), ldo, st);
}

// bytes of MLP hidden activations per row slab (0 = whole chunk at once); CNB_MLP_SLAB_MB overrides the default
static int64_t mlp_slab_bytes() {
  static const int64_t v = [] {
    const char* e = getenv("CNB_MLP_SLAB_MB");
    return (int64_t)(e ? atoi(e) : kDefaultMlpSlabMB) << 20;
  }();
  return v;
}

// encode `nb` clips (nb <= chunk) starting at wav; writes frame_embs (nb, T', 768) and optionally clip probs
template <typename ActT>
static int encode_chunk(cnb_handle* h, const float* wav, int nb, int64_t n, float* frame_embs, float* clip_probs, Tap* tap,
                        cudaStream_t st) {
  const Geometry g = geometry(n);
  const int64_t p0 = (int64_t)g.h[0] * kStageW[0];
  WS(h, "logmel", float, (size_t)nb * g.t * kMels, logmel);
  WS(h, "xa", float, (size_t)nb * p0 * kDims[0], xa);
  WS(h, "xb", float, (size_t)nb * p0 * kDims[0] / 2, xb);
  WS(h, "y", ActT, (size_t)nb * p0 * kDims[0], y);
  WS(h, "hid", ActT, (size_t)nb * p0 * kDims[0] * 4, hid);

  if (!h->pre_stem_done) {
    { Prof _p(h, CNB_K_FRONTEND, st); if (int rc = launch_frontend(wav, nb, n, h->fe, true, logmel, st)) return rc; }
    if (tap && tap->kind == CNB_TAP_LOGMEL_BN) return copy_tap(tap, logmel, (int64_t)nb * g.t * kMels, false, st);
    { Prof _p(h, CNB_K_STEM, st); if (int rc = launch_stem(logmel, nb, g.t, g.h[0], h->stem_w_t, h->stem_b, h->stem_ln_g, h->stem_ln_b, xa, st)) return rc; }
  }
  if (tap && tap->kind == CNB_TAP_STEM) return copy_tap(tap, xa, (int64_t)nb * p0 * kDims[0], false, st);

  float* x = xa;
  float* x_other = xb;
  int bi = 0;
  for (int s = 0; s < 4; ++s) {
    const int c = kDims[s], hh = g.h[s], ww = kStageW[s];
    if (s > 0) {
      // downsample: LN(channels_first) + 2x2/s2 conv as pack + GEMM (K = 4*Cin)
      const int cin = kDims[s - 1], hin = g.h[s - 1], win = kStageW[s - 1];
      const DownW& d = h->down[s - 1];
      set_sm_budget(bi < h->enc_narrow_blocks ? h->enc_narrow_sms : kNumSMs);
      { Prof _p(h, CNB_K_DS_PACK, st); if (int rc = launch_ln_pack2x2<ActT>(x, nb, hin, win, cin, d.ln_g, d.ln_b, y, st)) return rc; }
      const int m = nb * hh * ww;
      EpiParams ep;
      ep.bias = d.bias;
      { Prof _p(h, CNB_K_DS_GEMM, st); if (int rc = mlp_gemm<ActT>(h, y, d.w, d.w_bf, m, c, 4 * cin, EPI_BIAS, ep, x_other, false, c, st)) return rc; }
      std::swap(x, x_other);
      if (tap && tap->kind == CNB_TAP_DOWN && tap->stage == s) return copy_tap(tap, x, (int64_t)m * c, false, st);
    }
    const int m = nb * hh * ww;
    for (int j = 0; j < kDepths[s]; ++j, ++bi) {
      const BlockW& b = h->blocks[bi];
      set_sm_budget(bi < h->enc_narrow_blocks ? h->enc_narrow_sms : kNumSMs);
      { Prof _p(h, CNB_K_DWLN_S0 + s, st); if (int rc = launch_dwconv_ln<ActT>(x, nb, hh, ww, c, b.dw_w_t, b.dw_b, b.ln_g, b.ln_b, y, st)) return rc; }
      if (tap && tap->kind == CNB_TAP_DWLN && tap->stage == s && tap->block == j)
        return copy_tap(tap, y, (int64_t)m * c, sizeof(ActT) == 2, st);
      // The MLP runs in row slabs whose hidden activations (slab x 4C) fit the L2: pw1 writes them, pw2 reads them back and
      // every slab reuses the SAME hidden buffer, so the dirty lines are overwritten in L2 instead of travelling to HBM and
      // back (stage 1: 693 MB written + 693 MB read per block otherwise).  Slabs are multiples of the 128-row GEMM tile.
      int64_t slab = m;
      if (sizeof(ActT) == 2 && mlp_slab_bytes() > 0) {
        const int64_t rows_fit = mlp_slab_bytes() / ((int64_t)4 * c * sizeof(ActT));
        const int64_t n_slab = ceil_div(m, rows_fit > 128 ? rows_fit : 128);
        slab = ceil_div(ceil_div(m, n_slab), 128) * 128;
      }
      // stage 1, fp16 operands: one kernel for the whole MLP, the hidden tile never leaves the SM (mlp_fused.cu)
      static const bool fused_ok = getenv("CNB_NO_MLP_FUSED") == nullptr;
      static const bool fused192_ok = fused_ok && getenv("CNB_NO_MLP_FUSED192") == nullptr;
      static const bool fused384_ok = fused_ok && getenv("CNB_NO_MLP_FUSED384") == nullptr;
      if (sizeof(ActT) == 2 && c == 96 && fused_ok) {
        Prof _p(h, CNB_K_GEMM_PW1_S0 + s, st);
        if (int rc = launch_mlp_fused_c96(reinterpret_cast<const act16*>(y), b.w1_bf, b.w2_bf, b.b1, b.b2, b.scale, x, m, st))
          return rc;
        slab = m;  // skip the two-GEMM path below
      } else if (sizeof(ActT) == 2 && ((c == 192 && fused192_ok) || (c == 384 && fused384_ok))) {
        // stages 2-3: the same fusion over CTA pairs with the weights streamed and the hidden tile walked in chunks (mlp_fused_pair.cu)
        Prof _p(h, CNB_K_GEMM_PW1_S0 + s, st);
        if (int rc = launch_mlp_fused_pair(c, reinterpret_cast<const act16*>(y), b.w1_bf, b.w2_bf, b.b1, b.b2, b.scale, x, m, st))
          return rc;
        slab = m;
      } else
      for (int64_t m0 = 0; m0 < m; m0 += slab) {
        const int mm = (int)std::min<int64_t>(slab, m - m0);
        EpiParams e1;
        e1.bias = b.b1;
        { Prof _p(h, CNB_K_GEMM_PW1_S0 + s, st); if (int rc = mlp_gemm<ActT>(h, y + m0 * c, b.w1, b.w1_bf, mm, 4 * c, c, EPI_BIAS_GELU, e1, hid, true, 4 * c, st)) return rc; }
        EpiParams e2;
        e2.bias = b.b2;
        e2.scale = b.scale;
        e2.resid = x + m0 * c;
        { Prof _p(h, CNB_K_GEMM_PW2_S0 + s, st); if (int rc = mlp_gemm<ActT>(h, hid, b.w2, b.w2_bf, mm, c, 4 * c, EPI_SCALE_RESID, e2, x + m0 * c, false, c, st)) return rc; }
      }
      if (tap && tap->kind == CNB_TAP_BLOCK && tap->stage == s && tap->block == j)
        return copy_tap(tap, x, (int64_t)m * c, false, st);
    }
  }
  set_sm_budget(kNumSMs);
  { Prof _p(h, CNB_K_HEAD, st); if (int rc = launch_freq_mean(x, nb, g.tp, kStageW[3], 768, frame_embs, st)) return rc; }
  if (clip_probs) {
    Prof _p(h, CNB_K_HEAD, st);
    if (int rc = launch_clip_head(frame_embs, nb, g.tp, h->head_ln_g, h->head_ln_b, h->head_w, h->head_b, kTags, clip_probs,
                                  st))
      return rc;
  }
  return 0;
}

__global__ void act16_to_f32_kernel(const act16* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = act2float(in[i]);
}

static int copy_tap(Tap* tap, const void* src, int64_t elems, bool src_is_f16, cudaStream_t st) {
  if (elems > tap->cap) {
    set_error("encoder tap needs " + std::to_string(elems) + " elements but the output holds " + std::to_string(tap->cap));
    return -1;
  }
  if (src_is_f16) {
    act16_to_f32_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, st>>>(reinterpret_cast<const act16*>(src), tap->out,
                                                                        elems);
    CNB_LAUNCH_OK();
  } else {
    CNB_CUDA_OK(cudaMemcpyAsync(tap->out, src, elems * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  tap->hit = true;
  return 0;
}

static int chunk_size(const cnb_handle* h) { return h->cfg.enc_chunk > 0 ? h->cfg.enc_chunk : kDefaultChunk; }

static int encode(cnb_handle* h, const float* wav, int batch, int64_t n, float* frame_embs, float* clip_probs,
                  cudaStream_t st) {
  const Geometry g = geometry(n);
  const int chunk = chunk_size(h);
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int nb = std::min(chunk, batch - b0);
    float* fe = frame_embs + (int64_t)b0 * g.tp * 768;
    float* cp = clip_probs ? clip_probs + (int64_t)b0 * kTags : nullptr;
    int rc = (h->cfg.precision == CNB_PRECISION_PARITY)
                 ? encode_chunk<float>(h, wav + (int64_t)b0 * n, nb, n, fe, cp, nullptr, st)
                 : encode_chunk<act16>(h, wav + (int64_t)b0 * n, nb, n, fe, cp, nullptr, st);
    if (rc) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// decoder orchestration
// ---------------------------------------------------------------------------------------------------------------------
struct DecWs {
  float *mem, *ckv, *x, *qkv, *attn, *tmp, *ff, *logits, *kc, *vc, *part;
};
constexpr int kFf2Splits = 8;  // split-K slices of the 2048-deep FF2 GEMM

static int dec_prepare(cnb_handle* h, int batch, int tp, int rows, int max_len, DecWs* w) {
  const int V = h->cfg.vocab_size;
  WS(h, "mem", float, (size_t)batch * tp * kD, mem);
  WS(h, "ckv", float, (size_t)batch * tp * kLayers * 2 * kD, ckv);
  WS(h, "dx", float, (size_t)rows * kD, x);
  WS(h, "dqkv", float, (size_t)rows * 3 * kD, qkv);
  WS(h, "dattn", float, (size_t)rows * kD, attn);
  WS(h, "dtmp", float, (size_t)rows * kD, tmp);
  WS(h, "dff", float, (size_t)rows * kFF, ff);
  WS(h, "dlogits", float, (size_t)rows * V, logits);
  WS(h, "kc", float, (size_t)kLayers * rows * max_len * kD, kc);
  WS(h, "vc", float, (size_t)kLayers * rows * max_len * kD, vc);
  WS(h, "dpart", float, (size_t)kFf2Splits * rows * kD, part);
  *w = DecWs{mem, ckv, x, qkv, attn, tmp, ff, logits, kc, vc, part};
  return 0;
}

// projection: Linear(768,256) + ReLU (reference common.py:71-78); cross-attention K|V of all 6 layers in one GEMM
// rows [A1 | A2 | A1] of the triple-depth GEMM (see concat3_f16): A1 = fp16(a), A2 = fp16(a - A1)
__global__ void split3_act16_kernel(const float* __restrict__ in, act16* __restrict__ out, int64_t rows, int k) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per pair of columns
  const int kp = k >> 1;
  if (i >= rows * kp) return;
  const int64_t r = i / kp;
  const int c = 2 * (int)(i - r * kp);
  const float2 v = *reinterpret_cast<const float2*>(in + r * k + c);
  const act16x2 h1 = floats2act2(v.x, v.y);
  const float2 f1 = __half22float2(h1);
  const act16x2 h2 = floats2act2(v.x - f1.x, v.y - f1.y);
  act16* o = out + r * 3 * k + c;
  *reinterpret_cast<act16x2*>(o) = h1;
  *reinterpret_cast<act16x2*>(o + k) = h2;
  *reinterpret_cast<act16x2*>(o + 2 * k) = h1;
}

static int dec_project(cnb_handle* h, const float* frame_embs, int batch, int tp, const DecWs& w, cudaStream_t st) {
  Prof _p(h, CNB_K_PROJ_KV, st);
  EpiParams ep, ek;
  ep.bias = h->proj_b;
  ek.bias = h->ca_kv_b;
  const int m = batch * tp;
  static const bool simt = getenv("CNB_PROJ_SIMT") != nullptr;
  if (h->cfg.precision == CNB_PRECISION_FAST && !simt && m > 0) {
    // fast mode: both GEMMs on the tensor cores at fp32-level accuracy (split operands, one fp16 GEMM of triple depth each)
    WS(h, "mem3", act16, (size_t)m * 3 * 768, a3);
    split3_act16_kernel<<<(unsigned)ceil_div((int64_t)m * 384, 256), 256, 0, st>>>(frame_embs, a3, m, 768);
    CNB_LAUNCH_OK();
    if (int rc = launch_gemm_tc<float>(a3, h->proj_w3, m, kD, 3 * 768, EPI_BIAS_RELU, ep, w.mem, kD, st)) return rc;
    split3_act16_kernel<<<(unsigned)ceil_div((int64_t)m * (kD / 2), 256), 256, 0, st>>>(w.mem, a3, m, kD);
    CNB_LAUNCH_OK();
    return launch_gemm_tc<float>(a3, h->ca_kv_w3, m, kLayers * 2 * kD, 3 * kD, EPI_BIAS, ek, w.ckv, kLayers * 2 * kD, st);
  }
  if (int rc = launch_gemm_f32<float>(frame_embs, 768, h->proj_w, m, kD, 768, EPI_BIAS_RELU, ep, w.mem, kD, st)) return rc;
  return launch_gemm_f32<float>(w.mem, kD, h->ca_kv_w, m, kLayers * 2 * kD, kD, EPI_BIAS, ek, w.ckv, kLayers * 2 * kD, st);
}

// one decoder step for position `pos`: tokens[r][pos] -> logits (R, V)
static int dec_step(cnb_handle* h, const DecWs& w, const int* tokens, const int* src_row, const int* lens, int pos,
                    const DecoderDims& dd, const int* done, cudaStream_t st) {
  const int R = dd.rows;
  const int64_t cache_l = (int64_t)R * dd.max_len * kD;
  const int64_t kv_stride = kLayers * 2 * kD;
  { Prof _p(h, CNB_K_DEC_ATTN, st); if (int rc = launch_embed(tokens, pos, h->emb, h->pe, w.x, dd, done, st)) return rc; }
  for (int l = 0; l < kLayers; ++l) {
    const LayerW& L = h->layers[l];
    EpiParams e;
    e.bias = L.sa_in_b;
    { Prof _p(h, CNB_K_DEC_GEMM, st);     if (int rc = launch_gemm_f32_panel(w.x, kD, L.sa_in_w, R, 3 * kD, kD, EPI_BIAS, e, w.qkv, 3 * kD, st)) return rc; }
    { Prof _p(h, CNB_K_DEC_ATTN, st);     if (int rc = launch_self_attn(w.qkv, w.kc + l * cache_l, w.vc + l * cache_l, src_row, pos, w.attn, dd, done, st)) return rc; }
    e.bias = L.sa_out_b;
    { Prof _p(h, CNB_K_DEC_GEMM, st);     if (int rc = launch_gemm_f32_panel(w.attn, kD, L.sa_out_w, R, kD, kD, EPI_BIAS, e, w.tmp, kD, st)) return rc; }
    { Prof _p(h, CNB_K_DEC_ATTN, st);     if (int rc = launch_add_ln(w.x, w.tmp, 1, nullptr, L.n1_g, L.n1_b, R, done, st)) return rc; }
    e.bias = L.ca_q_b;
    { Prof _p(h, CNB_K_DEC_GEMM, st);     if (int rc = launch_gemm_f32_panel(w.x, kD, L.ca_q_w, R, kD, kD, EPI_BIAS, e, w.tmp, kD, st)) return rc; }
    {
      Prof _p(h, CNB_K_DEC_ATTN, st);
      if (int rc = launch_cross_attn(w.tmp, w.ckv + (int64_t)l * 2 * kD, w.ckv + (int64_t)l * 2 * kD + kD, kv_stride, lens,
                                     w.attn, dd, done, st))
        return rc;
    }
    e.bias = L.ca_out_b;
    { Prof _p(h, CNB_K_DEC_GEMM, st);     if (int rc = launch_gemm_f32_panel(w.attn, kD, L.ca_out_w, R, kD, kD, EPI_BIAS, e, w.tmp, kD, st)) return rc; }
    { Prof _p(h, CNB_K_DEC_ATTN, st);     if (int rc = launch_add_ln(w.x, w.tmp, 1, nullptr, L.n2_g, L.n2_b, R, done, st)) return rc; }
    e.bias = L.l1_b;
    { Prof _p(h, CNB_K_DEC_GEMM, st);     if (int rc = launch_gemm_f32_panel(w.x, kD, L.l1_w, R, kFF, kD, EPI_BIAS_GELU, e, w.ff, kFF, st)) return rc; }
    e.bias = nullptr;  // FF2 is split-K: bias + residual + LayerNorm happen in the reducing add_ln
    { Prof _p(h, CNB_K_DEC_GEMM, st); if (int rc = launch_gemm_f32_panel(w.ff, kFF, L.l2_w, R, kD, kFF, EPI_BIAS, e, w.part, kD, st)) return rc; }
    { Prof _p(h, CNB_K_DEC_ATTN, st); if (int rc = launch_add_ln(w.x, w.part, kFf2Splits, L.l2_b, L.n3_g, L.n3_b, R, done, st)) return rc; }
  }
  EpiParams e;
  e.bias = h->cls_b;
  Prof _p(h, CNB_K_DEC_CLS, st);
  return launch_gemm_f32_panel(w.x, kD, h->cls_w, R, dd.vocab, kD, EPI_BIAS, e, w.logits, dd.vocab, st);
}

// cross-attention keys of every clip regrouped for the cluster decoder: ckt[clip][layer][head][dim][tpad] = ckv[clip, t][layer, 0:256]
// (frames contiguous per (head, dim): lane = frame reads coalesced words), zero beyond T'
__global__ void cross_k_transpose_kernel(const float* __restrict__ ckv, float* __restrict__ ckt, int batch, int tp, int tpad) {
  const int64_t total = (int64_t)batch * kLayers * kD * tpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % tpad);
    const int64_t r = i / tpad;          // (clip * 6 + layer) * 256 + head * 32 + dim
    const int c = (int)(r % kD), l = (int)((r / kD) % kLayers);
    const int64_t clip = r / (kD * kLayers);
    ckt[i] = t < tp ? ckv[(clip * tp + t) * (kLayers * 2 * kD) + l * 2 * kD + c] : 0.f;
  }
}

__global__ void gather_mult_kernel(BeamState st, int64_t* __restrict__ mult_preds, float* __restrict__ mult_lp,
                                   const int* __restrict__ best_len, int* __restrict__ info, int rows, int max_len,
                                   int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * max_len) mult_preds[i] = st.out_preds[i];
  if (i < rows) mult_lp[i] = st.out_lp[i];
  if (i < batch) info[2 + i] = best_len[i];
  if (i == 0) {
    info[0] = st.done[1];
    info[1] = st.done[0];
  }
}

__global__ void tokens_i64_to_i32_kernel(const int64_t* __restrict__ in, int* __restrict__ out, int rows, int steps,
                                         int out_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * steps) out[(i / steps) * out_stride + (i % steps)] = (int)in[i];
}
__global__ void iota_rows_kernel(int* __restrict__ src_row, int rows, int max_len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * max_len) src_row[i] = i / max_len;
}
__global__ void copy_logits_kernel(const float* __restrict__ logits, float* __restrict__ out, int rows, int vocab, int step,
                                   int steps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)rows * vocab) {
    const int r = (int)(i / vocab), v = (int)(i % vocab);
    out[((int64_t)r * steps + step) * vocab + v] = logits[i];
  }
}

// Teacher-forced scoring (reference pl_modules/conette.py:293-318 + nn/modules/ce_mean.py): log-softmax of one decode
// position evaluated at the given next token, one CTA per caption row; pad targets (ignore_index = pad_id 0) score 0.
__global__ void __launch_bounds__(256) score_token_kernel(const float* __restrict__ logits, const int* __restrict__ tokens,
                                                          float* __restrict__ token_lprobs, int vocab, int pos, int steps) {
  __shared__ float red_m[8], red_s[8];
  const int r = blockIdx.x;
  const float* row = logits + (int64_t)r * vocab;
  const int tgt = tokens[(int64_t)r * (steps + 1) + pos + 1];
  float m = -INFINITY, s = 0.f;
  for (int v = threadIdx.x; v < vocab; v += 256) {
    const float x = row[v];
    if (x > m) { s = s * __expf(m - x) + 1.f; m = x; } else { s += __expf(x - m); }
  }
  for (int o = 16; o; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mm = fmaxf(m, m2);
    s = (mm == -INFINITY) ? 0.f : s * __expf(m - mm) + s2 * __expf(m2 - mm);
    m = mm;
  }
  if ((threadIdx.x & 31) == 0) { red_m[threadIdx.x >> 5] = m; red_s[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red_m[0];
    for (int w = 1; w < 8; ++w) mm = fmaxf(mm, red_m[w]);
    float ss = 0.f;
    for (int w = 0; w < 8; ++w) ss += (red_m[w] == -INFINITY) ? 0.f : red_s[w] * expf(red_m[w] - mm);
    token_lprobs[(int64_t)r * steps + pos] = (tgt == 0) ? 0.f : row[tgt] - mm - logf(ss);
  }
}
// losses[r] = -mean over non-pad targets of token_lprobs[r, :] (CrossEntropyLossMean(ignore_index=pad, dim=1); count clamped to >= 1)
__global__ void score_reduce_kernel(const float* __restrict__ token_lprobs, const int* __restrict__ tokens,
                                    float* __restrict__ losses, int rows, int steps) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  int n = 0;
  for (int p = 0; p < steps; ++p)
    if (tokens[(int64_t)r * (steps + 1) + p + 1] != 0) { s -= token_lprobs[(int64_t)r * steps + p]; ++n; }
  losses[r] = s / (float)max(n, 1);
}

// everything of one decode call that runs on the device, in launch order (this is what gets captured into a CUDA graph)
static int decode_body(cnb_handle* h, const DecWs& w, BeamState bs, const float* frame_embs, const int32_t* lens,
                       const int64_t* bos_ids, const uint8_t* forbid, int batch, int min_len, const DecoderDims& dd,
                       int64_t* preds, float* lprobs, int64_t* mult_preds, float* mult_lprobs, int32_t* info, int* best_len,
                       cudaStream_t st) {
  if (int rc = dec_project(h, frame_embs, batch, dd.tp, w, st)) return rc;
  if (int rc = launch_beam_init(bos_ids, bs, dd, st)) return rc;
  int cur = 0;
  for (int i = 0; i < dd.max_len; ++i) {
    if (int rc = dec_step(h, w, bs.tokens[cur], bs.src_row[cur], lens, i, dd, bs.done, st)) return rc;
    { Prof _p(h, CNB_K_BEAM, st); if (int rc = launch_beam_step(w.logits, forbid, bs, i, cur, min_len, dd, st)) return rc; }
    cur ^= 1;
  }
  if (int rc = launch_beam_finalize(bs, preds, lprobs, best_len, dd, st)) return rc;
  gather_mult_kernel<<<(dd.rows * dd.max_len + 255) / 256, 256, 0, st>>>(bs, mult_preds, mult_lprobs, best_len, info, dd.rows,
                                                                        dd.max_len, batch);
  CNB_LAUNCH_OK();
  return 0;
}

// Two decoder implementations:
//   * cluster (default of precision "fast"): the whole decode in ONE launch, fp16 hi/lo split tensor-core GEMMs with fp32-level
//     accuracy (decoder_cluster.cu);
//   * graph (precision "parity", or decoder="graph"): the fp32 CUDA-core KV-cached step of dec_step, ~1400 small dependent
//     launches captured once per (shape, buffer) signature into a CUDA graph and replayed on the handle's own stream (stream
//     capture is illegal on the legacy default stream callers often pass); fork/join events order it against the caller's
//     stream.  Profiling mode runs the same launches eagerly (event brackets).
// `tap` (tests only, cluster decoder): raw per-step logits (max_len, rows, V).
static int decode(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* bos_ids, const uint8_t* forbid,
                  int batch, int tp, int beam, int min_len, int max_len, int64_t* preds, float* lprobs, int64_t* mult_preds,
                  float* mult_lprobs, int32_t* info, cudaStream_t st, float* tap = nullptr) {
  const int rows = batch * beam;
  DecoderDims dd{rows, beam, tp, max_len, h->cfg.vocab_size};
  DecWs w;
  if (int rc = dec_prepare(h, batch, tp, rows, max_len, &w)) return rc;
  BeamState bs;
  WS(h, "tok0", int, (size_t)rows * (max_len + 1), tok0);
  WS(h, "tok1", int, (size_t)rows * (max_len + 1), tok1);
  WS(h, "src0", int, (size_t)rows * max_len, src0);
  WS(h, "src1", int, (size_t)rows * max_len, src1);
  WS(h, "sum_lp", float, rows, sum_lp);
  WS(h, "live", uint8_t, rows, live);
  WS(h, "out_preds", int64_t, (size_t)rows * max_len, out_preds);
  WS(h, "out_lp", float, rows, out_lp);
  WS(h, "done", int, 4, done);
  WS(h, "best_len", int, batch, best_len);
  bs.tokens[0] = tok0; bs.tokens[1] = tok1; bs.src_row[0] = src0; bs.src_row[1] = src1;
  bs.sum_lp = sum_lp; bs.live = live; bs.out_preds = out_preds; bs.out_lp = out_lp; bs.done = done;

  ClusterArgs ca;
  for (int l = 0; l < kLayers; ++l) {
    const LayerW& L = h->layers[l];
    ca.layers[l] = ClusterLayer{L.sa_in_b, L.sa_out_b, L.ca_q_b, L.ca_out_b, L.l1_b, L.l2_b, L.n1_g, L.n1_b, L.n2_g, L.n2_b,
                                L.n3_g, L.n3_b};
  }
  ca.emb = h->emb; ca.pe = h->pe; ca.cls_b = h->cls_b; ca.tmaps = h->dec_tmaps;
  ca.ckv = w.ckv; ca.lens = lens; ca.bos_ids = bos_ids; ca.forbid = forbid; ca.kc = w.kc; ca.vc = w.vc; ca.bs = bs;
  ca.tap = tap; ca.trace = nullptr; ca.ckt = nullptr; ca.tpad = 0;
  ca.rows = rows; ca.beam = beam; ca.tp = tp; ca.max_len = max_len; ca.vocab = h->cfg.vocab_size; ca.min_len = min_len;
  ca.batch = batch; ca.compact = h->dec_compact ? 1 : 0;
  {
    // L2 eviction priority of the weight-ring loads.  A decode step touches the 43 MB of weights once per cluster and then not for
    // another ~190 us, during which the cross-attention K/V (25 MB at 64 clips) and the self-attention caches (up to 47 MB) are
    // re-read by latency-critical loads.  With the default policy the weight stream pushes those out of L2 -- how badly depends
    // on where the workspaces happen to be mapped (decoder 3.94 .. 4.13 ms between processes) -- and still comes from DRAM every
    // step itself (ncu: 37 MB of DRAM reads per step).  evict_first on the stream keeps the attention data resident:
    // 3.85 ms at every placement (profiles/r2_decoder_l2_policy.txt).  CNB_DEC_L2=none|normal|last|first overrides.
    const char* e = getenv("CNB_DEC_L2");
    ca.w_policy = !e ? kL2EvictFirst
                     : !strcmp(e, "last") ? kL2EvictLast : !strcmp(e, "first") ? kL2EvictFirst : !strcmp(e, "normal") ? kL2EvictNormal : 0ull;
    const char* k = getenv("CNB_DEC_KV_L2");
    ca.kv_policy = k && !strcmp(k, "last") ? kL2EvictLast : kL2EvictNormal;
  }

  const bool want_cluster = h->use_cluster == 2 || (h->use_cluster == 1 && h->cfg.precision == CNB_PRECISION_FAST);
  if (want_cluster && (h->use_cluster == 2 || decoder_cluster_supported(ca))) {
    // one cluster of 8 CTAs per group of clips decodes start to finish: 2 GEMMs + key regrouping + 1 kernel + 2 gathers
    if (int rc = dec_project(h, frame_embs, batch, tp, w, st)) return rc;
    const int tpad = (tp + 31) / 32 * 32;
    WS(h, "ckt", float, (size_t)batch * kLayers * kD * tpad, ckt);
    {
      Prof _p(h, CNB_K_PROJ_KV, st);
      const int64_t total = (int64_t)batch * kLayers * kD * tpad;
      cross_k_transpose_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 8), 256, 0, st>>>(w.ckv, ckt, batch, tp, tpad);
      CNB_LAUNCH_OK();
    }
    ca.ckt = ckt;
    ca.tpad = tpad;
    // debug: time between phase marks as seen by thread 0 of the first CTA, summed over steps and layers.  CNB_DEC_TRACE=1
    // reads it back right after the launch (serialises the host); =2 prints the PREVIOUS launch's marks before the next one,
    // so the traced decode still overlapped whatever was enqueued after it (the streaming API's next encoder).
    const char* trace_env = getenv("CNB_DEC_TRACE");
    const int trace_mode = trace_env ? (trace_env[0] == '2' ? 2 : 1) : 0;
    auto print_trace = [&](unsigned long long* dev) -> int {
      unsigned long long t[24];
      CNB_CUDA_OK(cudaStreamSynchronize(st));
      CNB_CUDA_OK(cudaMemcpy(t, dev, sizeof(t), cudaMemcpyDeviceToHost));
      static const char* names[] = {"qkv gemm", "self attn", "sa_out gemm (K-split)", "ln1", "ca_q gemm",
                                    "cross attn", "ca_out gemm (K-split)", "ln2", "ff1 gemm", "ff2 gemm",
                                    "ln3", "cls gemm (rounds)", "cls scan (rounds)", "beam exchange", "beam merge",
                                    "3x wait reduce-scatter", "3x sum + all-gather push", "3x wait all-gather"};
      double tot = 0;
      for (int i = 0; i < 18; ++i) tot += (double)t[i];
      if (tot == 0) return 0;
      for (int i = 0; i < 18; ++i)
        fprintf(stderr, "[dec cluster trace] %-24s %9.1f us  (%4.1f %%)\n", names[i], (double)t[i] / 1e3, 100.0 * t[i] / tot);
      const double mhz = t[19] ? 1e3 * (double)t[18] / (double)t[19] : 0.0;
      fprintf(stderr, "[dec cluster trace] total %.1f us; SM clock of the traced CTA %.0f MHz; its MMA issuer waited %.1f us for weight chunks\n",
              tot / 1e3, mhz, mhz > 0 ? (double)t[20] / mhz : 0.0);
      return 0;
    };
    if (trace_mode) {
      WS(h, "dtrace_cl", unsigned long long, 32, tr);
      static bool primed = false;  // the first launch has no predecessor (and the fresh buffer is not zeroed)
      if (trace_mode == 2 && primed)
        if (int rc = print_trace(tr)) return rc;
      primed = true;
      CNB_CUDA_OK(cudaMemsetAsync(tr, 0, 32 * sizeof(unsigned long long), st));
      ca.trace = tr;
    }
    {
      Prof _p(h, CNB_K_DEC_GEMM, st);
      if (int rc = launch_decoder_cluster(ca, st)) return rc;
    }
    if (trace_mode == 1)
      if (int rc = print_trace(ca.trace)) return rc;
    if (int rc = launch_beam_finalize(bs, preds, lprobs, best_len, dd, st)) return rc;
    gather_mult_kernel<<<(rows * max_len + 255) / 256, 256, 0, st>>>(bs, mult_preds, mult_lprobs, best_len, info, rows, max_len,
                                                                    batch);
    CNB_LAUNCH_OK();
    return 0;
  }
  CNB_REQUIRE(tap == nullptr, "the logits tap belongs to the cluster decoder (cnb_decoder_logits taps the fp32 step)");
  if (h->prof_on || !h->use_graphs)
    return decode_body(h, w, bs, frame_embs, lens, bos_ids, forbid, batch, min_len, dd, preds, lprobs, mult_preds, mult_lprobs,
                       info, best_len, st);

  const std::vector<int64_t> key = {batch, tp, beam, min_len, max_len, (int64_t)frame_embs, (int64_t)lens, (int64_t)bos_ids,
                                    (int64_t)forbid, (int64_t)preds, (int64_t)lprobs, (int64_t)mult_preds,
                                    (int64_t)mult_lprobs, (int64_t)info, (int64_t)w.logits, (int64_t)w.kc, (int64_t)tok0};
  DecGraph* dg = nullptr;
  for (auto& g : h->dec_graphs)
    if (g.key == key) dg = &g;
  if (!dg) {
    const int64_t before = g_launches.load();
    CNB_CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = decode_body(h, w, bs, frame_embs, lens, bos_ids, forbid, batch, min_len, dd, preds, lprobs, mult_preds,
                               mult_lprobs, info, best_len, h->stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    const int64_t n_kernels = g_launches.load() - before;
    g_launches.fetch_sub(n_kernels);  // captured, not executed
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    CNB_CUDA_OK(ce);
    cudaGraphExec_t exec = nullptr;
    CNB_CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
    CNB_CUDA_OK(cudaGraphDestroy(graph));
    if (h->dec_graphs.size() >= 8) {  // bounded cache: drop the oldest signature
      cudaGraphExecDestroy(h->dec_graphs.front().exec);
      h->dec_graphs.erase(h->dec_graphs.begin());
    }
    h->dec_graphs.push_back(DecGraph{key, exec, n_kernels});
    dg = &h->dec_graphs.back();
  }
  CNB_CUDA_OK(cudaEventRecord(h->ev_fork, st));
  CNB_CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_fork, 0));
  CNB_CUDA_OK(cudaGraphLaunch(dg->exec, h->stream));
  g_launches.fetch_add(dg->n_kernels);
  CNB_CUDA_OK(cudaEventRecord(h->ev_join, h->stream));
  CNB_CUDA_OK(cudaStreamWaitEvent(st, h->ev_join, 0));
  return 0;
}

static void frame_lens_host(const int64_t* x_lens, int batch, int64_t n, std::vector<int32_t>* out) {
  const Geometry g = geometry(n);
  const int64_t red = n / g.tp;  // reference convnext.py:313
  out->resize(batch);
  for (int b = 0; b < batch; ++b) {
    const int64_t len = x_lens ? x_lens[b] : n;
    // torch: float32 division then round-half-to-even (convnext.py:315)
    const float q = (float)len / (float)red;
    (*out)[b] = (int32_t)std::nearbyintf(q);
  }
}

// The device-buffer entry points (caller's stream) and the split-phase host API (the handle's own streams) share the
// un-suffixed workspaces (log-mel, activations, decoder buffers ...).  Ordering between the two families:
//   dev_enter: the caller's stream waits for every host batch still in flight;
//   dev_leave: records the end of the device-path work, which the next cnb_caption_host_begin makes its streams wait for.
static int dev_enter(cnb_handle* h, cudaStream_t st) {
  for (int slot = 0; slot < 2; ++slot)
    if (h->host_pending[slot]) CNB_CUDA_OK(cudaStreamWaitEvent(st, h->host_done[slot], 0));
  return 0;
}
static int dev_leave(cnb_handle* h, cudaStream_t st) {
  if (!h->ev_dev_last) CNB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_dev_last, cudaEventDisableTiming));
  CNB_CUDA_OK(cudaEventRecord(h->ev_dev_last, st));
  h->dev_last_valid = true;
  return 0;
}
#define DEV_SCOPE(h, st, call)                      \
  do {                                              \
    if (int _rc = dev_enter(h, st)) return _rc;     \
    if (int _rc = (call)) return _rc;               \
    return dev_leave(h, st);                        \
  } while (0)

}  // namespace cnb

// =====================================================================================================================
// extern "C"
// =====================================================================================================================
#define CHECK_HANDLE(h)                                       \
  CNB_REQUIRE((h) != nullptr, "null handle");                 \
  CNB_CUDA_OK(cudaSetDevice((h)->cfg.device))
#define CHECK_READY(h) \
  CHECK_HANDLE(h);     \
  CNB_REQUIRE((h)->finalized, "weights not finalized (call cnb_finalize_weights)")

extern "C" {

const char* cnb_last_error(void) { return g_error.c_str(); }
int cnb_abi_version(void) { return CNB_ABI_VERSION; }

int cnb_create(const cnb_config* cfg, cnb_handle** out) {
  CNB_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  CNB_REQUIRE(cfg->abi_version == CNB_ABI_VERSION, "ABI version mismatch");
  CNB_REQUIRE(cfg->vocab_size > 4, "vocab_size must cover the 4 special tokens");
  CNB_REQUIRE(cfg->precision == CNB_PRECISION_FAST || cfg->precision == CNB_PRECISION_PARITY, "unknown precision mode");
  int n_dev = 0;
  CNB_CUDA_OK(cudaGetDeviceCount(&n_dev));
  CNB_REQUIRE(cfg->device >= 0 && cfg->device < n_dev, "no such CUDA device");
  cudaDeviceProp prop;
  CNB_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    set_error(std::string("conette_b200 needs an sm_100-class GPU (B200); found ") + prop.name + " (sm_" +
              std::to_string(prop.major) + std::to_string(prop.minor) + "); there is no fallback path");
    return -5;
  }
  cnb_handle* h = new cnb_handle();
  h->cfg = *cfg;
  CNB_CUDA_OK(cudaSetDevice(cfg->device));
  CNB_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CNB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CNB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  // reserved[0] = decoder implementation: 0 default (cluster kernel in precision "fast" when the shape allows it, else the fp32
  // graph), 1 fp32 CUDA graph, 2 fp32 eager launches (debugging), 3 cluster kernel (error if the shape does not fit)
  const int dmode = cfg->reserved[0];
  if (dmode < 0 || dmode > 3) {
    set_error("cnb_create: reserved[0] (decoder mode) must be 0..3");
    delete h;
    return -1;
  }
  h->use_cluster = dmode == 0 ? 1 : (dmode == 3 ? 2 : 0);
  h->use_graphs = dmode != 2;
  *out = h;
  return 0;
}

int cnb_destroy(cnb_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (auto& kv : h->ws)
    if (kv.second.ptr) cudaFree(kv.second.ptr);
  if (h->arena.base) cudaFree(h->arena.base);
  if (h->zero_flag) cudaFree(h->zero_flag);
  if (h->dec_tmaps) cudaFree(h->dec_tmaps);
  for (auto e : h->prof_events) cudaEventDestroy(e);
  for (auto& g : h->dec_graphs) cudaGraphExecDestroy(g.exec);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (auto& e : h->ev_slice)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->host_done)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_enc)
    if (e) cudaEventDestroy(e);
  if (h->dec_stream) cudaStreamDestroy(h->dec_stream);
  if (h->pin_lens) cudaFreeHost(h->pin_lens);
  if (h->ev_lens) cudaEventDestroy(h->ev_lens);
  if (h->ev_dev_last) cudaEventDestroy(h->ev_dev_last);
  delete h;
  return 0;
}

int cnb_load_weight(cnb_handle* h, const char* name, const void* host_ptr, int32_t dtype, int32_t ndim, const int64_t* shape) {
  CNB_REQUIRE(h != nullptr && name != nullptr, "null argument");
  CNB_REQUIRE(!h->finalized, "weights already finalized");
  if (dtype != CNB_DTYPE_F32) return 0;  // integer / bool tensors are host-side state (task ids, forbid mask, ...)
  CNB_REQUIRE(host_ptr != nullptr && ndim >= 0 && ndim <= 8, "bad tensor");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(reinterpret_cast<const float*>(host_ptr), reinterpret_cast<const float*>(host_ptr) + n);
  h->staged[name] = std::move(t);
  return 0;
}

int cnb_finalize_weights(cnb_handle* h) {
  CNB_REQUIRE(h != nullptr, "null handle");
  CNB_REQUIRE(!h->finalized, "weights already finalized");
  return finalize(h);
}

int cnb_geometry(int64_t n_samples, int32_t* n_stft_frames, int32_t stage_heights[4], int32_t* n_out_frames) {
  CNB_REQUIRE(n_samples > 0, "n_samples must be positive");
  const Geometry g = geometry(n_samples);
  if (n_stft_frames) *n_stft_frames = g.t;
  if (stage_heights)
    for (int i = 0; i < 4; ++i) stage_heights[i] = g.h[i];
  if (n_out_frames) *n_out_frames = g.tp;
  return 0;
}

#define CHECK_ENCODER(h) \
  CNB_REQUIRE((h)->has_encoder, "this handle holds no audio encoder weights (decoder-only state dict): use cnb_decode / cnb_score_captions")

static int check_audio(int32_t batch, int64_t n) {
  CNB_REQUIRE(batch > 0, "empty batch");
  CNB_REQUIRE(n > 512, "clips must be longer than the 512-sample reflect padding");
  CNB_REQUIRE(geometry(n).h[0] >= 8, "padded batch shorter than 7360 samples: the last 2x2 downsample has no input row "
                                     "(the reference raises RuntimeError here too)");
  return 0;
}

int cnb_resample(cnb_handle* h, const float* wav_in, const int64_t* lens_in, int32_t batch, int64_t n_in, const float* taps,
                 const int32_t* tap_lo, int32_t orig_freq, int32_t new_freq, int32_t n_taps, int32_t width, float* wav_out,
                 int64_t n_out, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(wav_in && taps && tap_lo && wav_out, "null buffer");
  CNB_REQUIRE(batch > 0 && n_in > 0 && n_out > 0, "bad audio shape");
  return launch_resample(wav_in, batch, n_in, lens_in, taps, tap_lo, orig_freq, new_freq, n_taps, width, wav_out, n_out,
                         (cudaStream_t)stream);
}

int cnb_frontend(cnb_handle* h, const float* wav, int32_t batch, int64_t n, int32_t apply_bn, float* out, void* stream) {
  CHECK_READY(h);
  CHECK_ENCODER(h);
  CNB_REQUIRE(wav && out, "null buffer");
  CNB_REQUIRE(batch > 0 && n > 512, "bad audio shape");
  return launch_frontend(wav, batch, n, h->fe, apply_bn != 0, out, (cudaStream_t)stream);
}

int cnb_encoder(cnb_handle* h, const float* wav, int32_t batch, int64_t n, float* frame_embs_out, float* clip_probs_out,
                void* stream) {
  CHECK_READY(h);
  CHECK_ENCODER(h);
  CNB_REQUIRE(wav && frame_embs_out, "null buffer");
  if (int rc = check_audio(batch, n)) return rc;
  DEV_SCOPE(h, (cudaStream_t)stream, encode(h, wav, batch, n, frame_embs_out, clip_probs_out, (cudaStream_t)stream));
}

int cnb_encoder_tap(cnb_handle* h, const float* wav, int32_t batch, int64_t n, int32_t tap_kind, int32_t stage, int32_t block,
                    float* out, int64_t cap, void* stream) {
  CHECK_READY(h);
  CHECK_ENCODER(h);
  CNB_REQUIRE(wav && out, "null buffer");
  if (int rc = check_audio(batch, n)) return rc;
  CNB_REQUIRE(batch <= chunk_size(h), "tap batch must fit one encoder chunk");
  Tap tap;
  tap.kind = tap_kind; tap.stage = stage; tap.block = block; tap.out = out; tap.cap = cap;
  const Geometry g = geometry(n);
  WS(h, "tap_fe", float, (size_t)batch * g.tp * 768, fe);
  if (int rc = dev_enter(h, (cudaStream_t)stream)) return rc;
  int rc = (h->cfg.precision == CNB_PRECISION_PARITY)
               ? encode_chunk<float>(h, wav, batch, n, fe, nullptr, &tap, (cudaStream_t)stream)
               : encode_chunk<act16>(h, wav, batch, n, fe, nullptr, &tap, (cudaStream_t)stream);
  if (rc) return rc;
  CNB_REQUIRE(tap.hit, "no such tap point");
  return dev_leave(h, (cudaStream_t)stream);
}

static int check_decode(cnb_handle* h, int32_t batch, int32_t tp, int32_t beam, int32_t min_len, int32_t max_len) {
  CNB_REQUIRE(batch > 0 && tp > 0, "empty batch");
  CNB_REQUIRE(beam > 0 && beam <= CNB_MAX_BEAM, "beam_size must be in [1, 8] (CNB_MAX_BEAM)");
  CNB_REQUIRE(min_len >= 0, "min_pred_size must be >= 0");
  CNB_REQUIRE(max_len > 0 && max_len <= CNB_MAX_PRED_SIZE, "max_pred_size must be in [1, 64] (CNB_MAX_PRED_SIZE)");
  return 0;
}

int cnb_decode(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* bos_ids, const uint8_t* forbid,
               int32_t batch, int32_t tp, int32_t beam, int32_t min_len, int32_t max_len, int64_t* preds, float* lprobs,
               int64_t* mult_preds, float* mult_lprobs, int32_t* info, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(frame_embs && lens && bos_ids && preds && lprobs && mult_preds && mult_lprobs && info, "null buffer");
  if (int rc = check_decode(h, batch, tp, beam, min_len, max_len)) return rc;
  DEV_SCOPE(h, (cudaStream_t)stream, decode(h, frame_embs, lens, bos_ids, forbid, batch, tp, beam, min_len, max_len, preds, lprobs,
                                            mult_preds, mult_lprobs, info, (cudaStream_t)stream));
}

int cnb_decode_tap(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* bos_ids, const uint8_t* forbid,
                   int32_t batch, int32_t tp, int32_t beam, int32_t min_len, int32_t max_len, int64_t* preds, float* lprobs,
                   int64_t* mult_preds, float* mult_lprobs, int32_t* info, float* logits_out, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(frame_embs && lens && bos_ids && preds && lprobs && mult_preds && mult_lprobs && info && logits_out, "null buffer");
  if (int rc = check_decode(h, batch, tp, beam, min_len, max_len)) return rc;
  const int keep = h->use_cluster;
  h->use_cluster = 2;  // the tap lives in the cluster kernel: fail instead of falling back
  if (int rc = dev_enter(h, (cudaStream_t)stream)) return rc;
  const int rc = decode(h, frame_embs, lens, bos_ids, forbid, batch, tp, beam, min_len, max_len, preds, lprobs, mult_preds,
                        mult_lprobs, info, (cudaStream_t)stream, logits_out);
  h->use_cluster = keep;
  return rc ? rc : dev_leave(h, (cudaStream_t)stream);
}

int cnb_decoder_logits(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* tokens, int32_t batch,
                       int32_t tp, int32_t steps, float* logits_out, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(frame_embs && lens && tokens && logits_out, "null buffer");
  if (int rc = check_decode(h, batch, tp, 1, 0, steps)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = dev_enter(h, st)) return rc;
  DecoderDims dd{batch, 1, tp, steps, h->cfg.vocab_size};
  DecWs w;
  if (int rc = dec_prepare(h, batch, tp, batch, steps, &w)) return rc;
  if (int rc = dec_project(h, frame_embs, batch, tp, w, st)) return rc;
  WS(h, "tok0", int, (size_t)batch * (steps + 1), tok);
  WS(h, "src0", int, (size_t)batch * steps, src);
  tokens_i64_to_i32_kernel<<<(batch * steps + 255) / 256, 256, 0, st>>>(tokens, tok, batch, steps, steps + 1);
  CNB_LAUNCH_OK();
  iota_rows_kernel<<<(batch * steps + 255) / 256, 256, 0, st>>>(src, batch, steps);
  CNB_LAUNCH_OK();
  for (int i = 0; i < steps; ++i) {
    if (int rc = dec_step(h, w, tok, src, lens, i, dd, h->zero_flag, st)) return rc;
    copy_logits_kernel<<<(unsigned)ceil_div((int64_t)batch * dd.vocab, 256), 256, 0, st>>>(w.logits, logits_out, batch,
                                                                                          dd.vocab, i, steps);
    CNB_LAUNCH_OK();
  }
  return dev_leave(h, st);
}

int cnb_score_captions(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* captions, int32_t batch,
                       int32_t tp, int32_t n_caps, int32_t cap_len, float* token_lprobs_out, float* losses_out, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(frame_embs && lens && captions && token_lprobs_out && losses_out, "null buffer");
  CNB_REQUIRE(n_caps > 0 && n_caps <= CNB_MAX_BEAM, "n_caps must be in [1, 8] (CNB_MAX_BEAM)");
  CNB_REQUIRE(cap_len >= 2, "captions need at least BOS + one target token");
  const int steps = cap_len - 1;
  if (int rc = check_decode(h, batch, tp, n_caps, 0, steps)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = dev_enter(h, st)) return rc;
  const int rows = batch * n_caps;
  DecoderDims dd{rows, n_caps, tp, steps, h->cfg.vocab_size};
  DecWs w;
  if (int rc = dec_prepare(h, batch, tp, rows, steps, &w)) return rc;
  if (int rc = dec_project(h, frame_embs, batch, tp, w, st)) return rc;
  WS(h, "tok0", int, (size_t)rows * cap_len, tok);
  WS(h, "src0", int, (size_t)rows * steps, src);
  tokens_i64_to_i32_kernel<<<(rows * cap_len + 255) / 256, 256, 0, st>>>(captions, tok, rows, cap_len, cap_len);
  CNB_LAUNCH_OK();
  iota_rows_kernel<<<(rows * steps + 255) / 256, 256, 0, st>>>(src, rows, steps);
  CNB_LAUNCH_OK();
  for (int i = 0; i < steps; ++i) {
    if (int rc = dec_step(h, w, tok, src, lens, i, dd, h->zero_flag, st)) return rc;
    score_token_kernel<<<rows, 256, 0, st>>>(w.logits, tok, token_lprobs_out, dd.vocab, i, steps);
    CNB_LAUNCH_OK();
  }
  score_reduce_kernel<<<(rows + 127) / 128, 128, 0, st>>>(token_lprobs_out, tok, losses_out, rows, steps);
  CNB_LAUNCH_OK();
  return dev_leave(h, st);
}

// Encoder on `st`, projection + beam search on `st_dec` (the same stream, or the handle's high-priority decode stream when the
// split-phase host API keeps two batches in flight: then batch i decodes -- a latency-bound kernel that leaves most of every
// SM idle and 44 SMs untouched -- while batch i+1 is already being encoded).  `slot` selects the frame-embedding buffers,
// which must outlive the encoder of the following batch in that case.
static int caption_impl(cnb_handle* h, const float* wav, const int64_t* x_lens_host, const int64_t* bos_ids, const uint8_t* forbid,
                        int32_t batch, int64_t n, int32_t beam, int32_t min_len, int32_t max_len, int64_t* preds, float* lprobs,
                        int64_t* mult_preds, float* mult_lprobs, int32_t* info, float* clip_probs, cudaStream_t st,
                        cudaStream_t st_dec, int slot) {
  CHECK_READY(h);
  CHECK_ENCODER(h);
  CNB_REQUIRE(wav && bos_ids && preds && lprobs && mult_preds && mult_lprobs && info, "null buffer");
  if (int rc = check_audio(batch, n)) return rc;
  const Geometry g = geometry(n);
  if (int rc = check_decode(h, batch, g.tp, beam, min_len, max_len)) return rc;
  WS(h, slot ? "frame_embs.1" : "frame_embs", float, (size_t)batch * g.tp * 768, fe);
  WS(h, slot ? "lens.1" : "lens", int32_t, batch, lens);
  // frame counts go through a handle-owned pinned buffer: no host synchronisation between the copy and the launches.  The
  // buffer is reused by the next call, whose first write happens after this call's work has been enqueued on `st`; calls
  // that do not end with a synchronisation (device-buffer API) therefore wait for the previous copy first.
  if (h->pin_lens_cap < batch) {
    CNB_CUDA_OK(cudaStreamSynchronize(st));
    if (h->pin_lens) cudaFreeHost(h->pin_lens);
    CNB_CUDA_OK(cudaMallocHost(&h->pin_lens, sizeof(int32_t) * (size_t)batch));
    h->pin_lens_cap = batch;
  }
  if (!h->ev_lens) CNB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_lens, cudaEventDisableTiming));
  CNB_CUDA_OK(cudaEventSynchronize(h->ev_lens));  // returns at once when nothing was recorded yet
  {
    std::vector<int32_t> lens_h;
    frame_lens_host(x_lens_host, batch, n, &lens_h);
    memcpy(h->pin_lens, lens_h.data(), sizeof(int32_t) * (size_t)batch);
  }
  CNB_CUDA_OK(cudaMemcpyAsync(lens, h->pin_lens, batch * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CNB_CUDA_OK(cudaEventRecord(h->ev_lens, st));
  if (int rc = encode(h, wav, batch, n, fe, clip_probs, st)) return rc;
  if (st_dec != st) {
    if (!h->ev_enc[slot]) CNB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_enc[slot], cudaEventDisableTiming));
    CNB_CUDA_OK(cudaEventRecord(h->ev_enc[slot], st));
    CNB_CUDA_OK(cudaStreamWaitEvent(st_dec, h->ev_enc[slot], 0));
  }
  return decode(h, fe, lens, bos_ids, forbid, batch, g.tp, beam, min_len, max_len, preds, lprobs, mult_preds, mult_lprobs,
                info, st_dec);
}

int cnb_caption(cnb_handle* h, const float* wav, const int64_t* x_lens_host, const int64_t* bos_ids, const uint8_t* forbid,
                int32_t batch, int64_t n, int32_t beam, int32_t min_len, int32_t max_len, int64_t* preds, float* lprobs,
                int64_t* mult_preds, float* mult_lprobs, int32_t* info, float* clip_probs, void* stream) {
  CHECK_READY(h);
  DEV_SCOPE(h, (cudaStream_t)stream, caption_impl(h, wav, x_lens_host, bos_ids, forbid, batch, n, beam, min_len, max_len, preds,
                                                  lprobs, mult_preds, mult_lprobs, info, clip_probs, (cudaStream_t)stream,
                                                  (cudaStream_t)stream, 0));
}

int cnb_caption_host_begin(cnb_handle* h, const float* wav_host, const int64_t* x_lens_host, const int64_t* bos_ids_host,
                           const uint8_t* forbid_host, int32_t batch, int64_t n, int32_t beam, int32_t min_len, int32_t max_len,
                           int64_t* preds_host, float* lprobs_host, int64_t* mult_preds_host, float* mult_lprobs_host,
                           int32_t* info_host, float* clip_probs_host, int32_t* ticket_out) {
  CHECK_READY(h);
  CHECK_ENCODER(h);
  CNB_REQUIRE(wav_host && bos_ids_host && preds_host && lprobs_host && mult_preds_host && mult_lprobs_host && info_host &&
                  ticket_out,
              "null buffer");
  if (int rc = check_audio(batch, n)) return rc;
  const int V = h->cfg.vocab_size;
  const int rows = batch * beam;
  cudaStream_t st = h->stream;
  // two sets of staging buffers: the H2D copy of batch i+1 (copy stream) runs while batch i computes on `st`
  const int slot = h->host_next;
  h->host_next ^= 1;
  if (!h->host_done[slot]) CNB_CUDA_OK(cudaEventCreateWithFlags(&h->host_done[slot], cudaEventDisableTiming));
  if (h->host_pending[slot]) {  // the caller never collected the batch that used this slot: its buffers are still in use
    CNB_CUDA_OK(cudaEventSynchronize(h->host_done[slot]));
    h->host_pending[slot] = false;
  }
  if (h->dev_last_valid) {  // a device-buffer call may still be using the shared workspaces on the caller's stream
    CNB_CUDA_OK(cudaStreamWaitEvent(st, h->ev_dev_last, 0));
    h->dev_last_valid = false;
  }
  const std::string sfx = slot ? ".1" : ".0";
  WS(h, ("io_wav" + sfx).c_str(), float, (size_t)batch * n, wav);
  WS(h, ("io_bos" + sfx).c_str(), int64_t, batch, bos);
  WS(h, ("io_forbid" + sfx).c_str(), uint8_t, V, forbid);
  WS(h, ("io_preds" + sfx).c_str(), int64_t, (size_t)batch * max_len, preds);
  WS(h, ("io_lprobs" + sfx).c_str(), float, batch, lprobs);
  WS(h, ("io_mpreds" + sfx).c_str(), int64_t, (size_t)rows * max_len, mpreds);
  WS(h, ("io_mlprobs" + sfx).c_str(), float, rows, mlprobs);
  WS(h, ("io_info" + sfx).c_str(), int32_t, 2 + batch, info);
  WS(h, ("io_clip" + sfx).c_str(), float, (size_t)batch * kTags, clip);
  CNB_CUDA_OK(cudaMemcpyAsync(bos, bos_ids_host, batch * sizeof(int64_t), cudaMemcpyDefault, st));
  if (forbid_host) CNB_CUDA_OK(cudaMemcpyAsync(forbid, forbid_host, V, cudaMemcpyDefault, st));
  // `wav_host` may also be DEVICE memory (batches that are already resident, e.g. produced by cnb_resample): it is then used in
  // place and must stay untouched until _end.
  bool wav_on_device = false;
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, wav_host) == cudaSuccess) wav_on_device = attr.type == cudaMemoryTypeDevice;
    else cudaGetLastError();  // plain pageable host memory on old drivers: not an error
  }
  if (wav_on_device) wav = const_cast<float*>(wav_host);
  // Waveforms travel in up to 8 slices on a copy stream; the front-end and the stem of slice i run while slice i+1 is still
  // on the PCIe bus (both are per-clip kernels).  The copy stream does not wait for `st`: this slot's previous user has
  // completed (event above), so with two batches in flight the whole copy hides behind the other batch's compute.
  const int n_slices = (batch <= chunk_size(h) && batch >= 16 && (!h->prof_on || h->prof_timeline) && !wav_on_device) ? 8 : 1;
  if (n_slices > 1) {
    const Geometry g = geometry(n);
    if (!h->copy_stream) CNB_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& e : h->ev_slice)
      if (!e) CNB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    WS(h, "logmel", float, (size_t)batch * g.t * kMels, logmel);
    WS(h, "xa", float, (size_t)batch * g.h[0] * kStageW[0] * kDims[0], xa);
    const int per = (int)ceil_div(batch, n_slices);
    for (int i = 0, b0 = 0; b0 < batch; ++i, b0 += per) {
      const int nb = std::min(per, batch - b0);
      CNB_CUDA_OK(cudaMemcpyAsync(wav + (int64_t)b0 * n, wav_host + (int64_t)b0 * n, (size_t)nb * n * sizeof(float),
                                  cudaMemcpyHostToDevice, h->copy_stream));
      CNB_CUDA_OK(cudaEventRecord(h->ev_slice[i], h->copy_stream));
      CNB_CUDA_OK(cudaStreamWaitEvent(st, h->ev_slice[i], 0));
      float* lm = logmel + (int64_t)b0 * g.t * kMels;
      { Prof _p(h, CNB_K_FRONTEND, st); if (int rc = launch_frontend(wav + (int64_t)b0 * n, nb, n, h->fe, true, lm, st)) return rc; }
      Prof _p(h, CNB_K_STEM, st);
      if (int rc = launch_stem(lm, nb, g.t, g.h[0], h->stem_w_t, h->stem_b, h->stem_ln_g, h->stem_ln_b,
                               xa + (int64_t)b0 * g.h[0] * kStageW[0] * kDims[0], st))
        return rc;
    }
    h->pre_stem_done = true;
  } else if (!wav_on_device) {
    CNB_CUDA_OK(cudaMemcpyAsync(wav, wav_host, (size_t)batch * n * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  // decode on its own high-priority stream (cluster decoder only: the graph decoder replays on h->stream anyway)
  cudaStream_t sd = st;
  static const bool overlap_dec = getenv("CNB_NO_DEC_OVERLAP") == nullptr;
  if (overlap_dec && h->use_cluster != 0 && h->cfg.precision == CNB_PRECISION_FAST && (!h->prof_on || h->prof_timeline)) {
    if (!h->dec_stream) {
      int lo = 0, hi = 0;
      CNB_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CNB_CUDA_OK(cudaStreamCreateWithPriority(&h->dec_stream, cudaStreamNonBlocking, hi));
    }
    sd = h->dec_stream;
  }
  // another batch in flight = this batch's encoder will share the GPU with that batch's decoder
  dwconv_set_overlap_hint(sd != st && h->host_pending[slot ^ 1]);
  {
    static const int nb_env = [] { const char* e = getenv("CNB_ENC_NARROW_BLOCKS"); return e ? atoi(e) : 0; }();
    static const int ns_env = [] { const char* e = getenv("CNB_ENC_NARROW_SMS"); return e ? atoi(e) : 92; }();
    h->enc_narrow_blocks = sd != st && h->host_pending[slot ^ 1] ? nb_env : 0;
    h->enc_narrow_sms = ns_env;
  }
  h->dec_compact = sd != st && h->host_pending[slot ^ 1];  // steady-state streaming: this decode will overlap the next encoder
  const int rc_cap = caption_impl(h, wav, x_lens_host, bos, forbid_host ? forbid : nullptr, batch, n, beam, min_len, max_len,
                                  preds, lprobs, mpreds, mlprobs, info, clip_probs_host ? clip : nullptr, st, sd, slot);
  dwconv_set_overlap_hint(false);
  set_sm_budget(kNumSMs);
  h->enc_narrow_blocks = 0;
  h->dec_compact = false;
  h->pre_stem_done = false;
  if (rc_cap) return rc_cap;
  CNB_CUDA_OK(cudaMemcpyAsync(preds_host, preds, (size_t)batch * max_len * sizeof(int64_t), cudaMemcpyDeviceToHost, sd));
  CNB_CUDA_OK(cudaMemcpyAsync(lprobs_host, lprobs, batch * sizeof(float), cudaMemcpyDeviceToHost, sd));
  CNB_CUDA_OK(cudaMemcpyAsync(mult_preds_host, mpreds, (size_t)rows * max_len * sizeof(int64_t), cudaMemcpyDeviceToHost, sd));
  CNB_CUDA_OK(cudaMemcpyAsync(mult_lprobs_host, mlprobs, rows * sizeof(float), cudaMemcpyDeviceToHost, sd));
  CNB_CUDA_OK(cudaMemcpyAsync(info_host, info, (2 + batch) * sizeof(int32_t), cudaMemcpyDeviceToHost, sd));
  if (clip_probs_host)
    CNB_CUDA_OK(cudaMemcpyAsync(clip_probs_host, clip, (size_t)batch * kTags * sizeof(float), cudaMemcpyDeviceToHost, sd));
  CNB_CUDA_OK(cudaEventRecord(h->host_done[slot], sd));
  h->host_pending[slot] = true;
  *ticket_out = slot;
  return 0;
}

int cnb_caption_host_end(cnb_handle* h, int32_t ticket) {
  CHECK_READY(h);
  CNB_REQUIRE(ticket == 0 || ticket == 1, "unknown ticket");
  CNB_REQUIRE(h->host_pending[ticket], "no batch in flight for this ticket");
  CNB_CUDA_OK(cudaEventSynchronize(h->host_done[ticket]));
  h->host_pending[ticket] = false;
  return 0;
}

int cnb_caption_host(cnb_handle* h, const float* wav_host, const int64_t* x_lens_host, const int64_t* bos_ids_host,
                     const uint8_t* forbid_host, int32_t batch, int64_t n, int32_t beam, int32_t min_len, int32_t max_len,
                     int64_t* preds_host, float* lprobs_host, int64_t* mult_preds_host, float* mult_lprobs_host,
                     int32_t* info_host, float* clip_probs_host) {
  int32_t ticket = -1;
  if (int rc = cnb_caption_host_begin(h, wav_host, x_lens_host, bos_ids_host, forbid_host, batch, n, beam, min_len, max_len,
                                      preds_host, lprobs_host, mult_preds_host, mult_lprobs_host, info_host, clip_probs_host,
                                      &ticket))
    return rc;
  return cnb_caption_host_end(h, ticket);
}

__global__ void f32_to_act16_kernel(const float* __restrict__ in, act16* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = float2act(in[i]);
}

int cnb_debug_gemm(cnb_handle* h, const float* a, const float* w, const float* bias, const float* scale, const float* resid,
                   int32_t m, int32_t n, int32_t k, int32_t epi, int32_t use_tc, int32_t out_bf16, float* out, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(a && w && bias && out, "null buffer");
  CNB_REQUIRE(epi >= 0 && epi <= 3, "unknown epilogue");
  cudaStream_t st = (cudaStream_t)stream;
  EpiParams ep;
  ep.bias = bias; ep.scale = scale; ep.resid = resid;
  if (!use_tc) return launch_gemm_f32<float>(a, k, w, m, n, k, (Epilogue)epi, ep, out, n, st);
  WS(h, "dbg_a", act16, (size_t)m * k, a_bf);
  WS(h, "dbg_w", act16, (size_t)n * k, w_bf);
  f32_to_act16_kernel<<<(unsigned)ceil_div((int64_t)m * k, 256), 256, 0, st>>>(a, a_bf, (int64_t)m * k);
  CNB_LAUNCH_OK();
  f32_to_act16_kernel<<<(unsigned)ceil_div((int64_t)n * k, 256), 256, 0, st>>>(w, w_bf, (int64_t)n * k);
  CNB_LAUNCH_OK();
  if (!out_bf16) return launch_gemm_tc<float>(a_bf, w_bf, m, n, k, (Epilogue)epi, ep, out, n, st);
  CNB_REQUIRE(epi != EPI_SCALE_RESID, "the residual epilogue writes fp32");
  WS(h, "dbg_o", act16, (size_t)m * n, o_bf);
  if (int rc = launch_gemm_tc<act16>(a_bf, w_bf, m, n, k, (Epilogue)epi, ep, o_bf, n, st)) return rc;
  act16_to_f32_kernel<<<(unsigned)ceil_div((int64_t)m * n, 256), 256, 0, st>>>(o_bf, out, (int64_t)m * n);
  CNB_LAUNCH_OK();
  return 0;
}

int cnb_debug_mlp_fused(cnb_handle* h, const float* y, const float* w1, const float* b1, const float* w2, const float* b2,
                        const float* scale, float* x, int32_t m, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(y && w1 && b1 && w2 && b2 && scale && x, "null buffer");
  CNB_REQUIRE(m >= 0, "negative row count");
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) return 0;
  WS(h, "dbg_y", act16, (size_t)m * 96, y_bf);
  WS(h, "dbg_w1", act16, (size_t)384 * 96, w1_bf);
  WS(h, "dbg_w2", act16, (size_t)96 * 384, w2_bf);
  f32_to_act16_kernel<<<(unsigned)ceil_div((int64_t)m * 96, 256), 256, 0, st>>>(y, y_bf, (int64_t)m * 96);
  CNB_LAUNCH_OK();
  f32_to_act16_kernel<<<(unsigned)ceil_div((int64_t)384 * 96, 256), 256, 0, st>>>(w1, w1_bf, (int64_t)384 * 96);
  CNB_LAUNCH_OK();
  f32_to_act16_kernel<<<(unsigned)ceil_div((int64_t)384 * 96, 256), 256, 0, st>>>(w2, w2_bf, (int64_t)384 * 96);
  CNB_LAUNCH_OK();
  return launch_mlp_fused_c96(y_bf, w1_bf, w2_bf, b1, b2, scale, x, m, st);
}

int cnb_debug_mlp_fused_pair(cnb_handle* h, int32_t c, const float* y, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* scale, float* x, int32_t m, void* stream) {
  CHECK_READY(h);
  CNB_REQUIRE(y && w1 && b1 && w2 && b2 && scale && x, "null buffer");
  CNB_REQUIRE(c == 192 || c == 384, "C must be 192 or 384");
  CNB_REQUIRE(m >= 0, "negative row count");
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) return 0;
  const int64_t nw = (int64_t)4 * c * c;
  WS(h, "dbg_y", act16, (size_t)m * c, y_bf);
  WS(h, "dbg_w1", act16, (size_t)nw, w1_bf);
  WS(h, "dbg_w2", act16, (size_t)nw, w2_bf);
  f32_to_act16_kernel<<<(unsigned)ceil_div((int64_t)m * c, 256), 256, 0, st>>>(y, y_bf, (int64_t)m * c);
  CNB_LAUNCH_OK();
  f32_to_act16_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, st>>>(w1, w1_bf, nw);
  CNB_LAUNCH_OK();
  f32_to_act16_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, st>>>(w2, w2_bf, nw);
  CNB_LAUNCH_OK();
  return launch_mlp_fused_pair(c, y_bf, w1_bf, w2_bf, b1, b2, scale, x, m, st);
}

int cnb_profile_begin(cnb_handle* h) {
  CHECK_READY(h);
  h->prof_on = true;
  h->prof_used = 0;
  h->prof_class.clear();
  return 0;
}

int cnb_profile_end(cnb_handle* h, float* ms_per_class, int64_t* brackets_per_class, int32_t n_classes) {
  CHECK_READY(h);
  CNB_REQUIRE(ms_per_class && brackets_per_class && n_classes >= CNB_K_COUNT, "output arrays too small");
  h->prof_on = false;
  CNB_CUDA_OK(cudaDeviceSynchronize());
  for (int i = 0; i < n_classes; ++i) {
    ms_per_class[i] = 0.f;
    brackets_per_class[i] = 0;
  }
  for (size_t i = 0; i < h->prof_class.size(); ++i) {
    float ms = 0.f;
    CNB_CUDA_OK(cudaEventElapsedTime(&ms, h->prof_events[2 * i], h->prof_events[2 * i + 1]));
    ms_per_class[h->prof_class[i]] += ms;
    brackets_per_class[h->prof_class[i]] += 1;
  }
  h->prof_used = 0;
  h->prof_class.clear();
  return 0;
}

int cnb_profile_timeline_begin(cnb_handle* h) {
  CHECK_READY(h);
  h->prof_on = h->prof_timeline = true;
  h->prof_used = 0;
  h->prof_class.clear();
  return 0;
}

int cnb_profile_timeline_end(cnb_handle* h, int32_t* cls, float* t_begin_ms, float* t_end_ms, int32_t cap, int32_t* n_out) {
  CHECK_READY(h);
  CNB_REQUIRE(cls && t_begin_ms && t_end_ms && n_out && cap >= 0, "null buffer");
  h->prof_on = h->prof_timeline = false;
  CNB_CUDA_OK(cudaDeviceSynchronize());
  const int n = (int)std::min<size_t>(h->prof_class.size(), (size_t)cap);
  for (int i = 0; i < n; ++i) {
    cls[i] = h->prof_class[i];
    CNB_CUDA_OK(cudaEventElapsedTime(&t_begin_ms[i], h->prof_events[0], h->prof_events[2 * i]));
    CNB_CUDA_OK(cudaEventElapsedTime(&t_end_ms[i], h->prof_events[0], h->prof_events[2 * i + 1]));
  }
  *n_out = (int32_t)h->prof_class.size();
  h->prof_used = 0;
  h->prof_class.clear();
  return 0;
}

int64_t cnb_launch_count(const cnb_handle*) { return g_launches.load(); }
int64_t cnb_device_bytes(const cnb_handle* h) { return h ? (int64_t)(h->arena.cap + h->ws_bytes) : 0; }

}  // extern "C"
