// C ABI + layer-program executor (see include/yololite_b200.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace yl {

static thread_local std::string g_err;
long long g_tc_launches = 0, g_simt_launches = 0, g_post_launches = 0;
void set_error(const std::string& msg) { g_err = msg; }

struct BufShape {
  int H = 0, W = 0, C = 0;
  size_t off = 0, bytes = 0;
};

}  // namespace yl

struct yl_engine {
  int device = 0;
  std::vector<yl_op> ops;
  float* d_blob = nullptr;
  size_t blob_floats = 0;
  int n_buffers = 0, n_levels = 0;
  // plan
  int B = 0, H = 0, W = 0;
  std::vector<yl::BufShape> bufs;
  std::vector<int> level_shape;  // n_levels * 4
  std::vector<int> op_hout, op_wout, op_hin, op_win, op_hu, op_wu;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  int use_tc = 1;
  int sm_count = 148;
};

namespace yl {

int launch_tc_conv(const ConvParams& c, const float* wimg, int mode, int sm_count, cudaStream_t st);
bool tc_supported(int K, int N, int anchors, int mode, int dw_ks, int Hout, int Wout, int dw_stride);
bool stem2_supported(const ConvParams& c);
int launch_stem2(const ConvParams& c, const float* wimg, int sm_count, cudaStream_t st);

static int run_op(const yl_op& op, const float* blob, const float* in, const float* res, const float* up, float* out, int B,
                  int hin, int win, int hout, int wout, int hu, int wu, int use_tc, int sm_count, cudaStream_t st,
                  const unsigned char* in_u8 = nullptr) {
  ConvParams p{};
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
  if (op.kind == YL_OP_DWPW) {                                        // geometry / epilogue of the depthwise stage
    p.KS = op.k2; p.pad = op.k2 / 2; p.stride = op.stride2 > 1 ? op.stride2 : 1;
    p.b2 = op.b2_off >= 0 ? blob + op.b2_off : nullptr;
    p.act2 = op.act2;
  }
  if (op.kind == YL_OP_STEM2) {
    p.Cin = op.k2;                                // K of the second conv = 9 * stem channels
    p.b2 = op.b2_off >= 0 ? blob + op.b2_off : nullptr;     // fused pointwise conv after conv2 ([cout][cout] + cout biases)
    p.act2 = op.act2;
    ++g_tc_launches;
    static const int old_stem = [] { const char* e = getenv("YL_OLD_STEM"); return e ? atoi(e) : 0; }();
    if (op.w3_off >= 0 && !old_stem && stem2_supported(p)) return launch_stem2(p, blob + op.w3_off, sm_count, st);
    YL_REQUIRE(!in_u8, "uint8 image input needs the fused bf16 stem kernel (16/32-channel second conv, even H, W % 16 == 0)");
    YL_REQUIRE(!p.b2, "YL_OP_STEM2 with a fused pointwise conv needs the bf16-triple kernel (w3_off, 16 channels, W % 4 == 0)");
    return launch_tc_conv(p, blob + op.wt_off, 3, sm_count, st);
  }
  if (use_tc && op.wt_off >= 0 && (op.kind == YL_OP_CONV || op.kind == YL_OP_DWPW)) {
    const int mode = op.kind == YL_OP_DWPW ? 2 : (op.k == 1 && op.stride == 1) ? 0 : 1;
    const int K = (mode == 0 || mode == 2) ? op.cin : op.k * op.k * op.cin;
    // small layers (K or N < 32) are per-tile-overhead bound on the tensor-core pipeline and already stream at
    // ~2 TB/s on the SIMT kernel: keep them there unless the caller forces the tensor path (use_tc == 2)
    // rows of N % 4 != 0 floats are not 16-byte aligned: the tensor-core epilogues read residual / upsample sources with
    // 16-byte loads, so that (unused by the reference models) combination stays on the SIMT kernel
    const bool unaligned_addend = (op.cout & 3) != 0 && (res || up);
    const bool big = ((K >= 32 && op.cout >= 32) || use_tc == 2) && !unaligned_addend;
    if (big && (op.cin & 3) == 0 && tc_supported(K, op.cout, op.anchors, mode, op.kind == YL_OP_DWPW ? op.k2 : 0, hout, wout, op.kind == YL_OP_DWPW && op.stride2 > 1 ? op.stride2 : 1))
    { ++g_tc_launches; return launch_tc_conv(p, blob + op.wt_off, mode, sm_count, st); }
  }
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

static int plan(yl_engine* e, int B, int H, int W) {
  if (e->B == B && e->H == H && e->W == W && e->arena) return 0;
  YL_REQUIRE(B >= 1 && H >= 1 && W >= 1, "B,H,W must be positive");
  const int nops = (int)e->ops.size();
  std::vector<BufShape> bufs(e->n_buffers);
  std::vector<int> lvl(e->n_levels * 4, 0);
  e->op_hout.assign(nops, 0); e->op_wout.assign(nops, 0); e->op_hin.assign(nops, 0); e->op_win.assign(nops, 0);
  e->op_hu.assign(nops, 0); e->op_wu.assign(nops, 0);
  for (int i = 0; i < nops; ++i) {
    const yl_op& op = e->ops[i];
    int hin, win, cin;
    if (op.src == YL_SRC_INPUT) { hin = H; win = W; cin = 3; }
    else {
      YL_REQUIRE(op.src >= 0 && op.src < e->n_buffers, "op.src out of range");
      hin = bufs[op.src].H; win = bufs[op.src].W; cin = bufs[op.src].C;
      YL_REQUIRE(hin > 0, "op reads a buffer that was never written");
    }
    YL_REQUIRE(cin == op.cin, "op.cin does not match the source buffer");
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
    if (op.kind == YL_OP_STEM2) { hin = H; win = W; }
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
  for (auto& b : bufs) { b.off = off; off += (b.bytes + 255) / 256 * 256; }
  if (off > e->arena_bytes) {
    if (e->arena) YL_CHECK_CUDA(cudaFree(e->arena));
    e->arena = nullptr; e->arena_bytes = 0;
    YL_CHECK_CUDA(cudaMalloc(&e->arena, off));
    e->arena_bytes = off;
  }
  e->bufs = bufs; e->level_shape = lvl; e->B = B; e->H = H; e->W = W;
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
  YL_REQUIRE(n_buffers >= 1 && n_levels >= 1 && n_levels <= 8, "n_buffers >= 1, 1 <= n_levels <= 8");
  int ndev = 0;
  YL_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  YL_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this engine has no CPU fallback)");
  YL_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  YL_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  YL_REQUIRE(prop.major == 10, "yololite_b200 is built for sm_100a (B200) only");
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
  }
  yl_engine* e = new yl_engine();
  e->device = device;
  e->ops.assign(ops, ops + n_ops);
  e->n_buffers = n_buffers; e->n_levels = n_levels; e->blob_floats = blob_floats;
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
  cudaSetDevice(e->device);
  if (e->arena) cudaFree(e->arena);
  if (e->d_blob) cudaFree(e->d_blob);
  delete e;
  return 0;
}

int yl_engine_set_option(yl_engine* e, const char* key, int32_t value) {
  using namespace yl;
  YL_REQUIRE(e && key, "null argument");
  if (std::strcmp(key, "tensor_cores") == 0) { e->use_tc = value ? 1 : 0; return 0; }
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
  YL_CHECK_CUDA(cudaSetDevice(e->device));
  if (int rc = plan(e, B, H, W)) return rc;
  if (shapes) std::memcpy(shapes, e->level_shape.data(), e->level_shape.size() * sizeof(int));
  return 0;
}

static int forward_impl(yl_engine* e, const float* x, int32_t B, int32_t H, int32_t W, float* const* level_out,
                        void* stream, cudaEvent_t* ev, const unsigned char* x_u8 = nullptr) {
  using namespace yl;
  YL_REQUIRE(e && (x || x_u8) && level_out, "null argument");
  YL_REQUIRE(!x_u8 || (e->ops[0].kind == YL_OP_STEM2 && e->ops[0].src == YL_SRC_INPUT && e->ops[0].w3_off >= 0 && e->use_tc),
             "this layer program has no uint8 image entry (fused stem kernel required)");
  YL_CHECK_CUDA(cudaSetDevice(e->device));
  if (int rc = plan(e, B, H, W)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ev) YL_CHECK_CUDA(cudaEventRecord(ev[0], st));
  auto bufptr = [&](int id) -> float* { return reinterpret_cast<float*>(e->arena + e->bufs[id].off); };
  for (size_t i = 0; i < e->ops.size(); ++i) {
    const yl_op& op = e->ops[i];
    const float* in = op.src == YL_SRC_INPUT ? x : bufptr(op.src);
    float* outp;
    if (op.dst >= 0) outp = bufptr(op.dst);
    else {
      outp = level_out[-op.dst - 1];
      YL_REQUIRE(outp, "null level output pointer");
    }
    int rc = run_op(op, e->d_blob, in, op.res >= 0 ? bufptr(op.res) : nullptr, op.up >= 0 ? bufptr(op.up) : nullptr, outp, B,
                    e->op_hin[i], e->op_win[i], e->op_hout[i], e->op_wout[i], e->op_hu[i], e->op_wu[i], e->use_tc,
                    e->sm_count, st, op.src == YL_SRC_INPUT ? x_u8 : nullptr);
    if (rc) return rc;
    if (ev) YL_CHECK_CUDA(cudaEventRecord(ev[i + 1], st));
  }
  return 0;
}

int yl_forward(yl_engine* e, const float* x, int32_t B, int32_t H, int32_t W, float* const* level_out, void* stream) {
  return forward_impl(e, x, B, H, W, level_out, stream, nullptr);
}

int yl_forward_u8(yl_engine* e, const uint8_t* images_bgr, int32_t B, int32_t H, int32_t W, float* const* level_out, void* stream) {
  return forward_impl(e, nullptr, B, H, W, level_out, stream, nullptr, images_bgr);
}

int yl_forward_profile(yl_engine* e, const float* x, int32_t B, int32_t H, int32_t W, float* const* level_out,
                       void* stream, float* op_ms, int32_t n_ops) {
  using namespace yl;
  YL_REQUIRE(e && op_ms && n_ops == (int)e->ops.size(), "op_ms must hold one float per op");
  YL_CHECK_CUDA(cudaSetDevice(e->device));
  std::vector<cudaEvent_t> ev(n_ops + 1);
  for (auto& v : ev) YL_CHECK_CUDA(cudaEventCreate(&v));
  int rc = forward_impl(e, x, B, H, W, level_out, stream, ev.data());
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
    YL_CHECK_CUDA(cudaMemcpyAsync(dst, e->arena + b.off, (size_t)e->B * b.H * b.W * b.C * sizeof(float),
                                  cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  }
  return 0;
}

}  // extern "C"
