// C ABI + layer-program executor (see include/yololite_b200.h).
//
// yl_engine_plan sizes everything once per input shape (activation arena, output-level buffers, postprocess scratch); per op the
// launch record of the tcgen05 kernels (kernel parameters + tensor maps, launch.cuh) is cached against the pointers it was
// built for, so a steady-state yl_forward is a list of kernel launches: no planning, no cuTensorMapEncodeTiled, no allocation.
// Kernels are launched with programmatic dependent launch (their prologue overlaps the previous kernel's tail) and, with the
// "graph" option, the whole call (forward, or forward + postprocess for yl_engine_detect) is captured once into a CUDA graph
// and replayed.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "launch.cuh"

namespace yl {

static thread_local std::string g_err;
std::atomic<long long> g_tc_launches{0}, g_simt_launches{0}, g_post_launches{0}, g_graph_launches{0}, g_graph_captures{0},
    g_prepares{0};
void set_error(const std::string& msg) { g_err = msg; }

unsigned long long* debug_words() {
  static unsigned long long* w = [] {
    unsigned long long* h = nullptr;
    if (cudaHostAlloc(&h, 32 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return (unsigned long long*)nullptr; }
    memset(h, 0, 32 * sizeof(unsigned long long));
    return h;
  }();
  return w;
}

struct BufShape {
  int H = 0, W = 0, C = 0;
  size_t off = 0, bytes = 0;
};

// the caller's current device is restored on return: the engine never changes it behind torch's back
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  int enter(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return -2; }
    if (prev != dev) {
      if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return -2; }
      changed = true;
    }
    return 0;
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

constexpr int MAX_FEATS = 8, MAX_LEVELS = 8;

// launch record of one op, valid for exactly these pointers
struct OpRec {
  const void *in = nullptr, *in_u8 = nullptr, *out = nullptr, *res = nullptr, *up = nullptr;
  int kind = 0;                 // 1 tcgen05 conv, 2 fused stem
  TcLaunch tc;
  Stem2Launch s2;
};
struct OpCache {
  std::vector<OpRec> recs;
  size_t next = 0;
};

// identity of one yl_forward / yl_engine_detect call: the captured graph bakes all of it in
struct CallKey {
  const void* x = nullptr; const void* x_u8 = nullptr;
  const void* feats[MAX_FEATS] = {};
  const void* levels[MAX_LEVELS] = {};
  const void* det[6] = {};       // boxes, scores, classes, anchor_idx, counts, packed
  int B = 0, H = 0, W = 0, detect = 0, img_size = 0, max_det = 0, cap = 0, use_tc = 0, pdl = 0;
  float conf = 0.f;
  double iou = 0.0;
};
struct GraphRec {
  CallKey key;
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  unsigned long long last_use = 0;
};

}  // namespace yl

struct yl_engine {
  int device = 0;
  std::vector<yl_op> ops;
  float* d_blob = nullptr;
  size_t blob_floats = 0;
  int n_buffers = 0, n_levels = 0, n_feats = 0;
  // plan
  int B = 0, H = 0, W = 0;
  std::vector<int> feat_dims;    // n_feats * 3 (H, W, C) of the planned shape
  std::vector<yl::BufShape> bufs;
  std::vector<int> level_shape;  // n_levels * 4
  std::vector<size_t> level_off; // engine-owned level buffers (yl_engine_detect) inside the arena
  std::vector<int> op_hout, op_wout, op_hin, op_win, op_hu, op_wu;
  size_t stem_tmp_off = 0, stem_mid_off = 0;   // unfused stem fallback: stem activation / conv2 output (only when the fused kernel cannot run)
  bool has_stem_tmp = false;
  size_t post_scratch_off = 0, post_scratch_bytes = 0;
  long long n_anchors = 0;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  int use_tc = 1, pdl = 1, use_graph = 0;
  int sm_count = 148;
  std::vector<yl::OpCache> cache;
  std::vector<yl::GraphRec> graphs;
  unsigned long long tick = 0;
  cudaStream_t cap_stream = nullptr;
};

namespace yl {

int launch_tc_conv(const ConvParams& c, const float* wimg, int mode, int sm_count, cudaStream_t st);
bool tc_supported(int K, int N, int anchors, int mode, int dw_ks, int Hout, int Wout, int dw_stride);
bool stem2_supported(const ConvParams& c);
int post_run(const float* const* level_logits, const int32_t* level_dims, int n_levels, int B, int D, int img_size, float conf, double iou,
             int max_det_per_class, int cap, float* boxes, float* scores, int64_t* classes, int64_t* anchor_idx, int32_t* counts,
             float* packed, void* scratch, size_t scratch_bytes, cudaStream_t st);

static void fill_params(ConvParams& p, const yl_op& op, const float* blob, const float* in, const float* res, const float* up, float* out,
                        int B, int hin, int win, int hout, int wout, int hu, int wu, const unsigned char* in_u8) {
  p = ConvParams{};
  p.in = in;
  p.in_u8 = in_u8;
  p.w = blob + op.w_off;
  p.bias = op.b_off >= 0 ? blob + op.b_off : nullptr;
  p.res = res; p.up = up;
  p.w2 = op.w2_off >= 0 ? blob + op.w2_off : nullptr;
  p.out = out;
  p.B = B; p.Hin = hin; p.Win = win; p.Cin = op.cin;
  p.Hout = hout; p.Wout = wout; p.Cout = op.cout;
  p.KS = op.k; p.stride = op.stride; p.pad = op.k / 2;
  p.Hu = hu; p.Wu = wu;
  p.act = op.act; p.anchors = op.anchors;
  p.wt_layout = op.wt_layout;
  if (op.kind == YL_OP_DWPW) {                                        // geometry / epilogue of the depthwise stage
    p.KS = op.k2; p.pad = op.k2 / 2; p.stride = op.stride2 > 1 ? op.stride2 : 1;
    p.b2 = op.b2_off >= 0 ? blob + op.b2_off : nullptr;
    p.act2 = op.act2;
  }
  if (op.kind == YL_OP_STEM2) {
    p.Cin = op.k2;                                // K of the second conv = 9 * stem channels
    p.b2 = op.b2_off >= 0 ? blob + op.b2_off : nullptr;     // fused pointwise conv after conv2 ([cout][cout] + cout biases)
    p.act2 = op.act2;
  }
}

// which tcgen05 mode (0 pointwise, 1 dense KxK, 2 depthwise -> pointwise) runs this op, or -1 for the fp32 SIMT kernels
static int tc_mode_for(const yl_op& op, const ConvParams& p, int use_tc, int hout, int wout) {
  if (!(use_tc && op.wt_off >= 0 && (op.kind == YL_OP_CONV || op.kind == YL_OP_DWPW))) return -1;
  const int mode = op.kind == YL_OP_DWPW ? 2 : (op.k == 1 && op.stride == 1) ? 0 : 1;
  const int K = (mode == 0 || mode == 2) ? op.cin : op.k * op.k * (op.wt_layout == 1 ? (op.cin + 31) / 32 * 32 : op.cin);
  // the per-tap padded image only exists for the TMA path of dense stride-1 convs; anything else with that layout runs on SIMT
  if (op.wt_layout == 1 && !(mode == 1 && op.stride == 1 && !p.up && op.anchors <= 1 && (op.cout & 3) == 0 &&
                             (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0))
    return -1;
  // small layers (K or N < 32) are per-tile-overhead bound on the tensor-core pipeline and already stream at ~2 TB/s on the
  // SIMT kernel: keep them there unless the caller forces the tensor path (use_tc == 2).  Rows of N % 4 != 0 floats are not
  // 16-byte aligned: the tensor-core epilogues read residual / upsample sources with 16-byte loads, so that (unused by the
  // reference models) combination stays on the SIMT kernel, as does a fused depthwise stage with N % 4 != 0.
  const bool unaligned = (op.cout & 3) != 0 && (p.res || p.up || mode == 2);
  const bool big = ((K >= 32 && op.cout >= 32) || use_tc == 2) && !unaligned;
  if (big && (op.cin & 3) == 0 &&
      tc_supported(K, op.cout, op.anchors, mode, op.kind == YL_OP_DWPW ? op.k2 : 0, hout, wout, op.kind == YL_OP_DWPW && op.stride2 > 1 ? op.stride2 : 1))
    return mode;
  return -1;
}

static int launch_simt(const yl_op& op, const ConvParams& p, cudaStream_t st) {
  ++g_simt_launches;
  switch (op.kind) {
    case YL_OP_STEM: return launch_stem(p, st);
    case YL_OP_CONV: return launch_conv_gemm(p, st);
    case YL_OP_DW: return launch_dw(p, st);
    case YL_OP_DWPW: return launch_dwpw(p, st);
    default: YL_REQUIRE(false, "unknown op kind");
  }
  return 0;
}

static const int g_old_stem = [] { const char* e = getenv("YL_OLD_STEM"); return e ? atoi(e) : 0; }();
static const int g_sync_each = [] { const char* e = getenv("YL_SYNC_EACH"); return e ? atoi(e) : 0; }();      // diagnostics: sync + check after every op
static const int g_pdl_env = [] { const char* e = getenv("YL_PDL"); return e ? atoi(e) : 1; }();      // YL_PDL=0: diagnostics, never use PDL

// YL_OP_STEM2 when the fused bf16 kernel cannot take the shape (W % 4 != 0, more than 32 conv2 channels ...): conv_stem on the
// fp32 SIMT kernel into `tmp_stem` [B,Hs,Ws,32], the 3x3 s2 conv as an ordinary dense conv, then the fused pointwise conv (timm
// blocks.0.1) as its own launch through `tmp_mid` [B,Ho,Wo,cout].
static int run_stem2_unfused(const yl_op& op, const ConvParams& p, const float* blob, float* tmp_stem, float* tmp_mid, int use_tc,
                             int sm_count, cudaStream_t st) {
  YL_REQUIRE(!p.in_u8, "uint8 image input needs the fused bf16 stem kernel (16/32-channel second conv, even H, W % 16 == 0)");
  YL_REQUIRE(tmp_stem && (!p.b2 || tmp_mid), "unfused stem fallback needs its intermediate buffers (plan the engine for this shape first)");
  const int sc = op.k2;                                           // stem channels
  const int Hs = (p.Hin + 2 - 3) / 2 + 1, Ws = (p.Win + 2 - 3) / 2 + 1;
  ConvParams s{};
  s.in = p.in; s.w = blob + op.w2_off; s.bias = s.w + 27 * sc; s.out = tmp_stem;
  s.B = p.B; s.Hin = p.Hin; s.Win = p.Win; s.Cin = 3; s.Hout = Hs; s.Wout = Ws; s.Cout = sc;
  s.KS = 3; s.stride = 2; s.pad = 1; s.act = YL_ACT_RELU;
  ++g_simt_launches;
  if (int rc = launch_stem(s, st)) return rc;
  ConvParams c2{};
  c2.in = tmp_stem; c2.w = p.w; c2.bias = p.bias; c2.out = p.b2 ? tmp_mid : p.out;
  c2.B = p.B; c2.Hin = Hs; c2.Win = Ws; c2.Cin = sc; c2.Hout = p.Hout; c2.Wout = p.Wout; c2.Cout = op.cout;
  c2.KS = op.k; c2.stride = op.stride; c2.pad = op.k / 2; c2.act = p.act;
  if (use_tc && op.wt_off >= 0 && op.cout >= 32 && (op.cout & 3) == 0 && tc_supported(op.k * op.k * sc, op.cout, 0, 1, 0, p.Hout, p.Wout, 1)) {
    ++g_tc_launches;
    if (int rc = launch_tc_conv(c2, blob + op.wt_off, 1, sm_count, st)) return rc;
  } else {
    ++g_simt_launches;
    if (int rc = launch_conv_gemm(c2, st)) return rc;
  }
  if (p.b2) {
    ConvParams r{};
    r.in = tmp_mid; r.w = p.b2; r.bias = p.b2 + (size_t)op.cout * op.cout; r.out = p.out;
    r.B = p.B; r.Hin = p.Hout; r.Win = p.Wout; r.Cin = op.cout; r.Hout = p.Hout; r.Wout = p.Wout; r.Cout = op.cout;
    r.KS = 1; r.stride = 1; r.pad = 0; r.act = p.act2;
    ++g_simt_launches;
    return launch_conv_gemm(r, st);
  }
  return 0;
}

// One op, no caching (yl_run_op and the profile path).
static int run_op(const yl_op& op, const float* blob, const float* in, const float* res, const float* up, float* out, int B,
                  int hin, int win, int hout, int wout, int hu, int wu, int use_tc, int sm_count, cudaStream_t st,
                  const unsigned char* in_u8 = nullptr) {
  ConvParams p;
  fill_params(p, op, blob, in, res, up, out, B, hin, win, hout, wout, hu, wu, in_u8);
  if (op.kind == YL_OP_STEM2) {
    if (op.w3_off >= 0 && !g_old_stem && stem2_supported(p)) {
      ++g_tc_launches;
      Stem2Launch L;
      if (int rc = stem2_prepare(p, blob + op.w3_off, sm_count, &L)) return rc;
      return stem2_launch(L, st, 0);
    }
    // single-op API: stream-ordered temporaries
    const size_t hs = (hin + 2 - 3) / 2 + 1, ws = (win + 2 - 3) / 2 + 1;
    float *t1 = nullptr, *t2 = nullptr;
    YL_CHECK_CUDA(cudaMallocAsync(&t1, (size_t)B * hs * ws * op.k2 * sizeof(float), st));
    if (op.b2_off >= 0) YL_CHECK_CUDA(cudaMallocAsync(&t2, (size_t)B * hout * wout * op.cout * sizeof(float), st));
    const int rc = run_stem2_unfused(op, p, blob, t1, t2, use_tc, sm_count, st);
    cudaFreeAsync(t1, st);
    if (t2) cudaFreeAsync(t2, st);
    return rc;
  }
  const int mode = tc_mode_for(op, p, use_tc, hout, wout);
  if (mode >= 0) { ++g_tc_launches; return launch_tc_conv(p, blob + op.wt_off, mode, sm_count, st); }
  return launch_simt(op, p, st);
}

static void drop_graphs(yl_engine* e) {
  for (auto& g : e->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  e->graphs.clear();
}

static int plan(yl_engine* e, int B, int H, int W, const int32_t* feat_dims) {
  const bool same_feats = e->n_feats == 0 || (feat_dims && std::equal(e->feat_dims.begin(), e->feat_dims.end(), feat_dims));
  if (e->B == B && e->H == H && e->W == W && e->arena && same_feats) return 0;
  YL_REQUIRE(B >= 1, "B must be positive");
  YL_REQUIRE(e->n_feats > 0 || (H >= 1 && W >= 1), "H,W must be positive");
  YL_REQUIRE(e->n_feats == 0 || feat_dims, "this layer program reads backbone features: pass their dimensions");
  const int nops = (int)e->ops.size();
  std::vector<BufShape> bufs(e->n_buffers);
  std::vector<int> lvl(e->n_levels * 4, 0);
  e->op_hout.assign(nops, 0); e->op_wout.assign(nops, 0); e->op_hin.assign(nops, 0); e->op_win.assign(nops, 0);
  e->op_hu.assign(nops, 0); e->op_wu.assign(nops, 0);
  size_t stem_tmp_bytes = 0, stem_mid_bytes = 0;
  for (int i = 0; i < nops; ++i) {
    const yl_op& op = e->ops[i];
    int hin, win, cin;
    if (op.src == YL_SRC_INPUT) { hin = H; win = W; cin = 3; }
    else if (op.src <= YL_SRC_FEATURE(0)) {
      const int f = YL_FEATURE_INDEX(op.src);
      YL_REQUIRE(f < e->n_feats, "feature input index out of range");
      hin = feat_dims[f * 3]; win = feat_dims[f * 3 + 1]; cin = feat_dims[f * 3 + 2];
      YL_REQUIRE(hin >= 1 && win >= 1, "feature dimensions must be positive");
    } else {
      YL_REQUIRE(op.src >= 0 && op.src < e->n_buffers, "op.src out of range");
      hin = bufs[op.src].H; win = bufs[op.src].W; cin = bufs[op.src].C;
      YL_REQUIRE(hin > 0, "op reads a buffer that was never written");
    }
    YL_REQUIRE(cin == op.cin, "op.cin does not match the source tensor");
    const int pad = op.k / 2;
    if (op.kind == YL_OP_STEM2) {                 // two stacked 3x3 s2 convs: size after the stem first
      hin = (hin + 2 - 3) / 2 + 1; win = (win + 2 - 3) / 2 + 1;
      YL_REQUIRE(hin >= 1 && win >= 1, "input too small for the network");
    }
    int hout = (hin + 2 * pad - op.k) / op.stride + 1;
    int wout = (win + 2 * pad - op.k) / op.stride + 1;
    if (op.kind == YL_OP_DWPW) {                   // the depthwise stage sets the output size
      const int s2 = op.stride2 > 1 ? op.stride2 : 1;
      hout = (hin + 2 * (op.k2 / 2) - op.k2) / s2 + 1;
      wout = (win + 2 * (op.k2 / 2) - op.k2) / s2 + 1;
    }
    YL_REQUIRE(hout >= 1 && wout >= 1, "input too small for the network");
    if (op.kind == YL_OP_STEM2) {
      hin = H; win = W;
      // the fused kernel needs W % 4 == 0 (TMA row pitch) and <= 32 conv2 channels; otherwise the unfused fallback runs and a
      // fused pointwise conv needs an intermediate tensor
      const bool fused_ok = op.w3_off >= 0 && !g_old_stem && (W & 3) == 0 && op.cout <= 32 && (op.cout & 3) == 0;
      if (!fused_ok) {
        stem_tmp_bytes = std::max(stem_tmp_bytes, (size_t)B * ((H + 2 - 3) / 2 + 1) * ((W + 2 - 3) / 2 + 1) * op.k2 * sizeof(float));
        if (op.b2_off >= 0) stem_mid_bytes = std::max(stem_mid_bytes, (size_t)B * hout * wout * op.cout * sizeof(float));
      }
    }
    e->op_hin[i] = hin; e->op_win[i] = win; e->op_hout[i] = hout; e->op_wout[i] = wout;
    if (op.dst >= 0) {
      YL_REQUIRE(op.dst < e->n_buffers, "op.dst out of range");
      BufShape& d = bufs[op.dst];
      d.H = hout; d.W = wout; d.C = op.cout;
      d.bytes = std::max(d.bytes, (size_t)B * hout * wout * op.cout * sizeof(float));
    } else {
      const int l = -op.dst - 1;
      YL_REQUIRE(l < e->n_levels && op.anchors >= 1 && op.cout % op.anchors == 0, "bad level output op");
      lvl[l * 4 + 0] = op.anchors; lvl[l * 4 + 1] = hout; lvl[l * 4 + 2] = wout; lvl[l * 4 + 3] = op.cout / op.anchors;
    }
    if (op.res >= 0) {
      YL_REQUIRE(op.res < e->n_buffers && bufs[op.res].H == hout && bufs[op.res].W == wout && bufs[op.res].C == op.cout,
                 "residual shape mismatch");
    }
    if (op.up >= 0) {
      YL_REQUIRE(op.up < e->n_buffers && bufs[op.up].C == op.cout && bufs[op.up].H > 0, "upsample source mismatch");
      e->op_hu[i] = bufs[op.up].H; e->op_wu[i] = bufs[op.up].W;
    }
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  for (auto& b : bufs) b.off = take(b.bytes);
  e->stem_tmp_off = take(stem_tmp_bytes);
  e->stem_mid_off = take(stem_mid_bytes);
  e->has_stem_tmp = stem_tmp_bytes > 0;
  // engine-owned output levels + postprocess scratch (yl_engine_detect)
  std::vector<size_t> loff(e->n_levels);
  long long N = 0;
  for (int l = 0; l < e->n_levels; ++l) {
    const long long n = (long long)lvl[l * 4] * lvl[l * 4 + 1] * lvl[l * 4 + 2];
    N += n;
    loff[l] = take((size_t)B * n * lvl[l * 4 + 3] * sizeof(float));
  }
  const size_t sbytes = yl_postprocess_scratch_bytes(B, N);
  const size_t soff = take(sbytes);
  if (off > e->arena_bytes) {
    if (e->arena) YL_CHECK_CUDA(cudaFree(e->arena));
    e->arena = nullptr; e->arena_bytes = 0;
    YL_CHECK_CUDA(cudaMalloc(&e->arena, off));
    e->arena_bytes = off;
  }
  e->bufs = bufs; e->level_shape = lvl; e->level_off = loff; e->n_anchors = N;
  e->post_scratch_off = soff; e->post_scratch_bytes = sbytes;
  e->B = B; e->H = H; e->W = W;
  if (e->n_feats) e->feat_dims.assign(feat_dims, feat_dims + 3 * e->n_feats);
  for (auto& c : e->cache) { c.recs.clear(); c.next = 0; }      // launch records hold the old shapes / pointers
  drop_graphs(e);
  return 0;
}

// Launch op i through its cached record (building it on first use for these pointers).
static int launch_cached(yl_engine* e, int i, const float* in, const unsigned char* in_u8, const float* res, const float* up, float* out,
                         cudaStream_t st, int pdl) {
  const yl_op& op = e->ops[i];
  OpCache& oc = e->cache[i];
  for (const OpRec& r : oc.recs)
    if (r.in == in && r.in_u8 == in_u8 && r.out == out && r.res == res && r.up == up) {
      ++g_tc_launches;
      return r.kind == 1 ? tc_launch(r.tc, st, pdl) : stem2_launch(r.s2, st, pdl);
    }
  ConvParams p;
  fill_params(p, op, e->d_blob, in, res, up, out, e->B, e->op_hin[i], e->op_win[i], e->op_hout[i], e->op_wout[i], e->op_hu[i], e->op_wu[i],
              in_u8);
  OpRec rec;
  rec.in = in; rec.in_u8 = in_u8; rec.out = out; rec.res = res; rec.up = up;
  if (op.kind == YL_OP_STEM2) {
    if (!(op.w3_off >= 0 && !g_old_stem && stem2_supported(p)))
      return run_stem2_unfused(op, p, e->d_blob, e->has_stem_tmp ? reinterpret_cast<float*>(e->arena + e->stem_tmp_off) : nullptr,
                               e->has_stem_tmp ? reinterpret_cast<float*>(e->arena + e->stem_mid_off) : nullptr, e->use_tc, e->sm_count, st);
    rec.kind = 2;
    if (int rc = stem2_prepare(p, e->d_blob + op.w3_off, e->sm_count, &rec.s2)) return rc;
  } else {
    const int mode = tc_mode_for(op, p, e->use_tc, e->op_hout[i], e->op_wout[i]);
    if (mode < 0) return launch_simt(op, p, st);
    rec.kind = 1;
    if (int rc = tc_prepare(p, e->d_blob + op.wt_off, mode, e->sm_count, &rec.tc)) return rc;
  }
  ++g_prepares;
  if (oc.recs.size() < 4) oc.recs.push_back(rec);
  else { oc.recs[oc.next] = rec; oc.next = (oc.next + 1) % 4; }
  ++g_tc_launches;
  return rec.kind == 1 ? tc_launch(rec.tc, st, pdl) : stem2_launch(rec.s2, st, pdl);
}

struct CallArgs {
  const float* x = nullptr;
  const unsigned char* x_u8 = nullptr;
  const float* const* feats = nullptr;
  float* levels[MAX_LEVELS] = {};
  // detect
  int detect = 0, img_size = 0, max_det = 0, cap = 0;
  float conf = 0.f;
  double iou = 0.0;
  float* boxes = nullptr; float* scores = nullptr; int64_t* classes = nullptr; int64_t* anchor_idx = nullptr;
  int32_t* counts = nullptr; float* packed = nullptr;
};

static int enqueue_ops(yl_engine* e, const CallArgs& a, cudaStream_t st, cudaEvent_t* ev) {
  if (ev) YL_CHECK_CUDA(cudaEventRecord(ev[0], st));
  auto bufptr = [&](int id) -> float* { return reinterpret_cast<float*>(e->arena + e->bufs[id].off); };
  const int pdl = (ev || !g_pdl_env) ? 0 : e->pdl;
  for (size_t i = 0; i < e->ops.size(); ++i) {
    const yl_op& op = e->ops[i];
    const float* in = op.src == YL_SRC_INPUT ? a.x : op.src <= YL_SRC_FEATURE(0) ? a.feats[YL_FEATURE_INDEX(op.src)] : bufptr(op.src);
    float* outp = op.dst >= 0 ? bufptr(op.dst) : a.levels[-op.dst - 1];
    YL_REQUIRE(outp, "null level output pointer");
    YL_REQUIRE(in || (op.src == YL_SRC_INPUT && a.x_u8), "null input pointer");
    if (int rc = launch_cached(e, (int)i, in, op.src == YL_SRC_INPUT ? a.x_u8 : nullptr, op.res >= 0 ? bufptr(op.res) : nullptr,
                               op.up >= 0 ? bufptr(op.up) : nullptr, outp, st, pdl))
      return rc;
    if (g_sync_each) {
      const cudaError_t se = cudaStreamSynchronize(st);
      if (se != cudaSuccess) {
        set_error("op " + std::to_string(i) + " (kind " + std::to_string(op.kind) + ", " + std::to_string(op.cin) + "->" + std::to_string(op.cout) +
                  ", k" + std::to_string(op.k) + "/" + std::to_string(op.k2) + ") failed: " + cudaGetErrorString(se));
        return -2;
      }
    }
    if (ev) YL_CHECK_CUDA(cudaEventRecord(ev[i + 1], st));
  }
  if (a.detect) {
    int32_t dims[MAX_LEVELS * 3];
    const float* lv[MAX_LEVELS];
    for (int l = 0; l < e->n_levels; ++l) {
      dims[l * 3] = e->level_shape[l * 4]; dims[l * 3 + 1] = e->level_shape[l * 4 + 1]; dims[l * 3 + 2] = e->level_shape[l * 4 + 2];
      lv[l] = a.levels[l];
    }
    return post_run(lv, dims, e->n_levels, e->B, e->level_shape[3], a.img_size, a.conf, a.iou, a.max_det, a.cap, a.boxes, a.scores,
                    a.classes, a.anchor_idx, a.counts, a.packed, e->arena + e->post_scratch_off, e->post_scratch_bytes, st);
  }
  return 0;
}

static int run_call(yl_engine* e, const CallArgs& a, int B, int H, int W, const int32_t* feat_dims, cudaStream_t st, cudaEvent_t* ev) {
  YL_REQUIRE(!a.x_u8 || (e->ops[0].kind == YL_OP_STEM2 && e->ops[0].src == YL_SRC_INPUT && e->ops[0].w3_off >= 0 && e->use_tc),
             "this layer program has no uint8 image entry (fused stem kernel required)");
  if (int rc = plan(e, B, H, W, feat_dims)) return rc;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (st != nullptr) { if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; } }
  if (!e->use_graph || ev || cs != cudaStreamCaptureStatusNone) return enqueue_ops(e, a, st, ev);

  // ---- CUDA graph: one capture per distinct call, then replays
  CallKey k;
  memset(&k, 0, sizeof(k));
  k.x = a.x; k.x_u8 = a.x_u8;
  for (int f = 0; f < e->n_feats; ++f) k.feats[f] = a.feats[f];
  for (int l = 0; l < e->n_levels; ++l) k.levels[l] = a.levels[l];
  k.det[0] = a.boxes; k.det[1] = a.scores; k.det[2] = a.classes; k.det[3] = a.anchor_idx; k.det[4] = a.counts; k.det[5] = a.packed;
  k.B = B; k.H = H; k.W = W; k.detect = a.detect; k.img_size = a.img_size; k.max_det = a.max_det; k.cap = a.cap;
  k.use_tc = e->use_tc; k.pdl = e->pdl; k.conf = a.conf; k.iou = a.iou;
  ++e->tick;
  GraphRec* slot = nullptr;
  for (auto& g : e->graphs)
    if (memcmp(&g.key, &k, sizeof(k)) == 0) { slot = &g; break; }
  if (slot && slot->exec) {
    slot->last_use = e->tick;
    ++g_graph_launches;
    YL_CHECK_CUDA(cudaGraphLaunch(slot->exec, st));
    return 0;
  }
  if (!slot) {
    // first call with these pointers: run eagerly (this builds the launch records and sets the kernel attributes outside
    // of any capture); the second call captures
    GraphRec g;
    g.key = k; g.last_use = e->tick;
    if (e->graphs.size() >= 8) {          // evict the least recently used
      size_t v = 0;
      for (size_t i = 1; i < e->graphs.size(); ++i) if (e->graphs[i].last_use < e->graphs[v].last_use) v = i;
      if (e->graphs[v].exec) cudaGraphExecDestroy(e->graphs[v].exec);
      if (e->graphs[v].graph) cudaGraphDestroy(e->graphs[v].graph);
      e->graphs[v] = g;
    } else e->graphs.push_back(g);
    return enqueue_ops(e, a, st, nullptr);
  }
  if (!e->cap_stream) YL_CHECK_CUDA(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  // capture on an engine-owned stream (the caller's may be the legacy default stream, which cannot be captured)
  YL_CHECK_CUDA(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_ops(e, a, e->cap_stream, nullptr);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
  YL_CHECK_CUDA(ce);
  slot->graph = graph; slot->last_use = e->tick;
  YL_CHECK_CUDA(cudaGraphInstantiate(&slot->exec, graph, 0));
  ++g_graph_captures;
  ++g_graph_launches;
  YL_CHECK_CUDA(cudaGraphLaunch(slot->exec, st));
  return 0;
}

}  // namespace yl

extern "C" {

const char* yl_last_error(void) { return yl::g_err.c_str(); }
long long yl_stat(const char* key) {
  if (!key) return -1;
  if (!std::strcmp(key, "tc_launches")) return yl::g_tc_launches;
  if (!std::strcmp(key, "simt_launches")) return yl::g_simt_launches;
  if (!std::strcmp(key, "post_launches")) return yl::g_post_launches;
  if (!std::strcmp(key, "graph_launches")) return yl::g_graph_launches;
  if (!std::strcmp(key, "graph_captures")) return yl::g_graph_captures;
  if (!std::strcmp(key, "prepares")) return yl::g_prepares;
  if (!std::strncmp(key, "stem_row_pixel", 14)) return yl::stem2_row_pixel(atoi(key + 14));
  if (!std::strcmp(key, "debug_reset")) { unsigned long long* w = yl::debug_words(); if (w) memset(w, 0, 32 * sizeof(unsigned long long)); return 0; }
  if (!std::strncmp(key, "trap_word", 9)) { unsigned long long* w = yl::debug_words(); const int i = atoi(key + 9) & 31; return w ? (long long)w[i] : 0; }
  return -1;
}
int yl_abi_version(void) { return YL_ABI_VERSION; }
int yl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int yl_engine_create(const yl_op* ops, int32_t n_ops, const float* blob_host, size_t blob_floats, int32_t n_buffers,
                     int32_t n_levels, int32_t device, yl_engine** out) {
  using namespace yl;
  YL_REQUIRE(ops && n_ops > 0 && blob_host && blob_floats > 0 && out, "null/empty arguments");
  YL_REQUIRE(n_buffers >= 1 && n_levels >= 1 && n_levels <= MAX_LEVELS, "n_buffers >= 1, 1 <= n_levels <= 8");
  int ndev = 0;
  YL_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  YL_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this engine has no CPU fallback)");
  DeviceGuard dg;
  YL_REQUIRE(dg.enter(device) == 0, "cannot select the CUDA device");
  cudaDeviceProp prop;
  YL_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  YL_REQUIRE(prop.major == 10, "yololite_b200 is built for sm_100a (B200) only");
  int n_feats = 0;
  for (int i = 0; i < n_ops; ++i) {
    const yl_op& op = ops[i];
    YL_REQUIRE(op.kind >= YL_OP_STEM && op.kind <= YL_OP_STEM2, "unknown op kind");
    YL_REQUIRE(op.kind != YL_OP_STEM2 || (op.wt_off >= 0 && op.w2_off >= 0 && op.k2 == 32 && op.k == 3 && op.stride == 2),
               "YL_OP_STEM2 needs the tcgen05 weight image, stem weights, 32 stem channels and a 3x3 s2 second conv");
    YL_REQUIRE(op.k >= 1 && (op.k & 1) && op.stride >= 1 && op.cin >= 1 && op.cout >= 1, "bad conv geometry");
    YL_REQUIRE(op.kind != YL_OP_DWPW || ((op.k2 == 3 || op.k2 == 5) && op.k == 1 && op.stride == 1 && op.w2_off >= 0 &&
                                         op.b2_off < (int64_t)blob_floats && op.act2 >= YL_ACT_NONE && op.act2 <= YL_ACT_SILU &&
                                         op.stride2 >= 0 && op.stride2 <= 2),
               "YL_OP_DWPW: depthwise 3x3 or 5x5 (stride 1 or 2) followed by a pointwise conv");
    YL_REQUIRE(op.w_off >= 0 && (size_t)op.w_off < blob_floats, "w_off out of range");
    YL_REQUIRE(op.b_off < (int64_t)blob_floats, "b_off out of range");
    YL_REQUIRE((op.w_off & 3) == 0 && (op.b_off < 0 || (op.b_off & 3) == 0), "blob offsets must be 16-byte aligned");
    YL_REQUIRE(op.wt_layout == 0 || (op.wt_layout == 1 && op.kind == YL_OP_CONV && op.k > 1 && op.stride == 1), "wt_layout 1: dense k x k stride-1 convs");
    if (op.src <= YL_SRC_FEATURE(0)) {
      YL_REQUIRE(YL_FEATURE_INDEX(op.src) < MAX_FEATS && op.kind == YL_OP_CONV, "feature inputs: at most 8, read by YL_OP_CONV ops");
      n_feats = std::max(n_feats, YL_FEATURE_INDEX(op.src) + 1);
    }
  }
  yl_engine* e = new yl_engine();
  e->device = device;
  e->ops.assign(ops, ops + n_ops);
  e->cache.resize(n_ops);
  e->n_buffers = n_buffers; e->n_levels = n_levels; e->blob_floats = blob_floats; e->n_feats = n_feats;
  e->sm_count = prop.multiProcessorCount;
  if (cudaMalloc(&e->d_blob, blob_floats * sizeof(float)) != cudaSuccess ||
      cudaMemcpy(e->d_blob, blob_host, blob_floats * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error(std::string("weight upload failed: ") + cudaGetErrorString(cudaGetLastError()));
    if (e->d_blob) cudaFree(e->d_blob);
    delete e;
    return -2;
  }
  *out = e;
  return 0;
}

int yl_engine_destroy(yl_engine* e) {
  if (!e) return 0;
  yl::DeviceGuard dg;
  dg.enter(e->device);
  yl::drop_graphs(e);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->arena) cudaFree(e->arena);
  if (e->d_blob) cudaFree(e->d_blob);
  delete e;
  return 0;
}

int yl_engine_set_option(yl_engine* e, const char* key, int32_t value) {
  using namespace yl;
  YL_REQUIRE(e && key, "null argument");
  if (std::strcmp(key, "tensor_cores") == 0) {
    const int v = value ? 1 : 0;
    if (v != e->use_tc) { for (auto& c : e->cache) { c.recs.clear(); c.next = 0; } }
    e->use_tc = v;
    return 0;
  }
  if (std::strcmp(key, "pdl") == 0) { e->pdl = value ? 1 : 0; return 0; }
  if (std::strcmp(key, "graph") == 0) { e->use_graph = value ? 1 : 0; return 0; }
  YL_REQUIRE(false, "unknown engine option");
  return -1;
}

int yl_run_op(const yl_op* op, const float* blob_dev, const float* in, const float* res, const float* up, float* out,
              int32_t B, int32_t Hin, int32_t Win, int32_t Hu, int32_t Wu, int32_t use_tensor_cores, void* stream) {
  using namespace yl;
  YL_REQUIRE(op && blob_dev && in && out, "null argument");
  YL_REQUIRE(op->k >= 1 && (op->k & 1) && op->stride >= 1, "bad conv geometry");
  int dev = 0;
  YL_CHECK_CUDA(cudaGetDevice(&dev));
  int sms = 0;
  YL_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int pad = op->k / 2;
  int hs = Hin, ws = Win;
  if (op->kind == YL_OP_STEM2) { hs = (Hin + 2 - 3) / 2 + 1; ws = (Win + 2 - 3) / 2 + 1; }
  int hout = (hs + 2 * pad - op->k) / op->stride + 1, wout = (ws + 2 * pad - op->k) / op->stride + 1;
  if (op->kind == YL_OP_DWPW) {
    const int s2 = op->stride2 > 1 ? op->stride2 : 1;
    hout = (Hin + 2 * (op->k2 / 2) - op->k2) / s2 + 1;
    wout = (Win + 2 * (op->k2 / 2) - op->k2) / s2 + 1;
  }
  return run_op(*op, blob_dev, in, res, up, out, B, Hin, Win, hout, wout, Hu, Wu, use_tensor_cores, sms,
                reinterpret_cast<cudaStream_t>(stream));
}

int yl_engine_plan(yl_engine* e, int32_t B, int32_t H, int32_t W, int32_t* shapes) {
  using namespace yl;
  YL_REQUIRE(e, "null engine");
  YL_REQUIRE(e->n_feats == 0, "this layer program reads backbone features: use yl_engine_plan_features");
  DeviceGuard dg;
  YL_REQUIRE(dg.enter(e->device) == 0, "cannot select the engine's device");
  if (int rc = plan(e, B, H, W, nullptr)) return rc;
  if (shapes) std::memcpy(shapes, e->level_shape.data(), e->level_shape.size() * sizeof(int));
  return 0;
}

int yl_engine_plan_features(yl_engine* e, int32_t B, const int32_t* feat_dims, int32_t n_feats, int32_t* shapes) {
  using namespace yl;
  YL_REQUIRE(e && feat_dims, "null argument");
  YL_REQUIRE(e->n_feats > 0 && n_feats == e->n_feats, "feature count does not match the layer program");
  DeviceGuard dg;
  YL_REQUIRE(dg.enter(e->device) == 0, "cannot select the engine's device");
  if (int rc = plan(e, B, 0, 0, feat_dims)) return rc;
  if (shapes) std::memcpy(shapes, e->level_shape.data(), e->level_shape.size() * sizeof(int));
  return 0;
}

static int forward_entry(yl_engine* e, const float* x, const uint8_t* x_u8, int32_t B, int32_t H, int32_t W, float* const* level_out,
                         void* stream, cudaEvent_t* ev) {
  using namespace yl;
  YL_REQUIRE(e && (x || x_u8) && level_out, "null argument");
  YL_REQUIRE(e->n_feats == 0, "this layer program reads backbone features: use yl_forward_features");
  DeviceGuard dg;
  YL_REQUIRE(dg.enter(e->device) == 0, "cannot select the engine's device");
  CallArgs a;
  a.x = x; a.x_u8 = x_u8;
  for (int l = 0; l < e->n_levels; ++l) { a.levels[l] = level_out[l]; YL_REQUIRE(a.levels[l], "null level output pointer"); }
  return run_call(e, a, B, H, W, nullptr, reinterpret_cast<cudaStream_t>(stream), ev);
}

int yl_forward(yl_engine* e, const float* x, int32_t B, int32_t H, int32_t W, float* const* level_out, void* stream) {
  return forward_entry(e, x, nullptr, B, H, W, level_out, stream, nullptr);
}

int yl_forward_u8(yl_engine* e, const uint8_t* images_bgr, int32_t B, int32_t H, int32_t W, float* const* level_out, void* stream) {
  return forward_entry(e, nullptr, images_bgr, B, H, W, level_out, stream, nullptr);
}

int yl_forward_features(yl_engine* e, const float* const* feats, const int32_t* feat_dims, int32_t n_feats, int32_t B,
                        float* const* level_out, void* stream) {
  using namespace yl;
  YL_REQUIRE(e && feats && feat_dims && level_out, "null argument");
  YL_REQUIRE(e->n_feats > 0 && n_feats == e->n_feats, "feature count does not match the layer program");
  DeviceGuard dg;
  YL_REQUIRE(dg.enter(e->device) == 0, "cannot select the engine's device");
  CallArgs a;
  a.feats = feats;
  for (int f = 0; f < n_feats; ++f) YL_REQUIRE(feats[f] && (reinterpret_cast<uintptr_t>(feats[f]) & 15) == 0, "feature pointers must be non-null and 16-byte aligned");
  for (int l = 0; l < e->n_levels; ++l) { a.levels[l] = level_out[l]; YL_REQUIRE(a.levels[l], "null level output pointer"); }
  return run_call(e, a, B, 0, 0, feat_dims, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int yl_engine_detect(yl_engine* e, const float* x, const uint8_t* images_bgr, int32_t B, int32_t H, int32_t W, int32_t img_size,
                     float conf, double iou, int32_t max_det_per_class, int32_t cap, float* boxes, float* scores, int64_t* classes,
                     int64_t* anchor_idx, int32_t* counts, float* packed, void* stream) {
  using namespace yl;
  YL_REQUIRE(e && ((x != nullptr) != (images_bgr != nullptr)), "pass exactly one of x (fp32 NCHW) and images_bgr (uint8 HWC)");
  YL_REQUIRE(e->n_feats == 0, "this layer program reads backbone features");
  YL_REQUIRE(cap >= 1 && (packed || (boxes && scores && classes && anchor_idx && counts)), "null output pointer");
  DeviceGuard dg;
  YL_REQUIRE(dg.enter(e->device) == 0, "cannot select the engine's device");
  if (int rc = plan(e, B, H, W, nullptr)) return rc;
  CallArgs a;
  a.x = x; a.x_u8 = images_bgr;
  for (int l = 0; l < e->n_levels; ++l) a.levels[l] = reinterpret_cast<float*>(e->arena + e->level_off[l]);
  a.detect = 1; a.img_size = img_size; a.conf = conf; a.iou = iou; a.max_det = max_det_per_class; a.cap = cap;
  a.boxes = boxes; a.scores = scores; a.classes = classes; a.anchor_idx = anchor_idx; a.counts = counts; a.packed = packed;
  return run_call(e, a, B, H, W, nullptr, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int yl_engine_levels(yl_engine* e, float** level_ptrs, int32_t* shapes) {
  using namespace yl;
  YL_REQUIRE(e && e->arena, "engine has no plan yet");
  for (int l = 0; l < e->n_levels; ++l) if (level_ptrs) level_ptrs[l] = reinterpret_cast<float*>(e->arena + e->level_off[l]);
  if (shapes) std::memcpy(shapes, e->level_shape.data(), e->level_shape.size() * sizeof(int));
  return 0;
}

int yl_forward_profile(yl_engine* e, const float* x, int32_t B, int32_t H, int32_t W, float* const* level_out,
                       void* stream, float* op_ms, int32_t n_ops) {
  using namespace yl;
  YL_REQUIRE(e && op_ms && n_ops == (int)e->ops.size(), "op_ms must hold one float per op");
  DeviceGuard dg;
  YL_REQUIRE(dg.enter(e->device) == 0, "cannot select the engine's device");
  std::vector<cudaEvent_t> ev(n_ops + 1);
  for (auto& v : ev) YL_CHECK_CUDA(cudaEventCreate(&v));
  int rc = forward_entry(e, x, nullptr, B, H, W, level_out, stream, ev.data());
  if (rc == 0) {
    if (cudaEventSynchronize(ev[n_ops]) != cudaSuccess) { set_error("event sync failed"); rc = -2; }
    for (int i = 0; i < n_ops && rc == 0; ++i)
      if (cudaEventElapsedTime(&op_ms[i], ev[i], ev[i + 1]) != cudaSuccess) { set_error("elapsed time failed"); rc = -2; }
  }
  for (auto& v : ev) cudaEventDestroy(v);
  return rc;
}

int yl_engine_read_buffer(yl_engine* e, int32_t buf, float* dst, int32_t* dims, void* stream) {
  using namespace yl;
  YL_REQUIRE(e && e->arena, "engine has no plan yet (call yl_forward first)");
  YL_REQUIRE(buf >= 0 && buf < e->n_buffers, "buffer id out of range");
  const BufShape& b = e->bufs[buf];
  if (dims) { dims[0] = b.H; dims[1] = b.W; dims[2] = b.C; }
  if (dst) {
    DeviceGuard dg;
    YL_REQUIRE(dg.enter(e->device) == 0, "cannot select the engine's device");
    YL_CHECK_CUDA(cudaMemcpyAsync(dst, e->arena + b.off, (size_t)e->B * b.H * b.W * b.C * sizeof(float),
                                  cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  }
  return 0;
}

}  // extern "C"
