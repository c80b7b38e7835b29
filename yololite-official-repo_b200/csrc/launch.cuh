// Launch records of the tcgen05 kernels: everything a launch needs (kernel parameters, tensor maps, grid, shared memory) is
// computed ONCE per (op, shape, pointers) by *_prepare and replayed by *_launch, so yl_forward does no planning, no
// cuTensorMapEncodeTiled and no allocation per call.  Also: programmatic dependent launch (PDL) helpers.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>
#include <utility>

#include "common.cuh"

namespace yl {

// ---- PDL (griddepcontrol): a kernel launched with the programmatic-stream-serialization attribute may start while the previous
// kernel in the stream is still draining; everything before pdl_wait() (barrier init, TMEM allocation, weight copies: constant
// data only) overlaps the previous kernel's tail.  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- packed fp32 pairs (sm_100: FFMA2 / FADD2, two IEEE fp32 operations per issue slot; results identical to the scalar ones).
// The 64-bit values are register pairs: packing / unpacking compiles to no instruction when the pair is already adjacent
// (e.g. the halves of a 16-byte shared-memory load).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// two fp32 -> packed bf16 pair (element 0 in the low half) for each of the three splits: x = x1 + x2 + x3 up to 2^-24 |x|.
// RELU: round-toward-zero splits keep every residual's sign, so the .relu of the conversion IS the ReLU (a negative input yields
// 0 | 0 | 0) and x = x1 + x2 + x3 exactly.  The residuals x - x1 are one FFMA2 (x1 * -1 + x, exact) per pair.
template <bool RELU>
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  const f32x2 neg1 = pack2(-1.f, -1.f);
  f32x2 ab = pack2(a, b);
  if (RELU) asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(b), "f"(a));
  ab = fma2(pack2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xFFFF0000u)), neg1, ab);
  unpack2(ab, a, b);
  if (RELU) asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(b), "f"(a));
  ab = fma2(pack2(__uint_as_float(p2 << 16), __uint_as_float(p2 & 0xFFFF0000u)), neg1, ab);
  unpack2(ab, a, b);
  if (RELU) asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(b), "f"(a));
}

// One lane of a CONVERGED warp.  Guarding the single-thread roles (tcgen05.mma / TMA issue) with elect.sync instead of `lane == 0`
// tells ptxas that exactly one thread is active: the UTCHMMA / UTMALDG operands then live in uniform registers and the instructions
// issue back to back.  With `lane == 0` every one of them sat in a PLOP3 / ELECT / BRA.U.ANY "waterfall" loop (~7 dependent
// instructions per MMA), and the issuing thread of the stem kernel was busy 93 % of the time -- it was the bottleneck.
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t st, int pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies per DEVICE: run `set` once for every device a launch is prepared on.
template <typename F>
inline int once_per_device(bool* done /*[64]*/, std::mutex& m, F set) {
  int dev = 0;
  YL_CHECK_CUDA(cudaGetDevice(&dev));
  YL_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  std::lock_guard<std::mutex> lk(m);
  if (!done[dev]) {
    if (int rc = set()) return rc;
    done[dev] = true;
  }
  return 0;
}

struct TcParams {
  ConvParams c;
  const float* wimg;   // bf16-triple image [nslab][3 splits][Npad][32 k], 64 B rows pre-swizzled (SWIZZLE_64B), see packer.tc_image
  int mode;            // 0 pointwise, 1 dense KxK as im2col (Cin%4==0), 2 depthwise KSxKS -> pointwise
  int K, nslab, Npad, Nc, nchunks, stages, tmem_cols;
  int tiles_x, tiles_y; // spatial modes: tiles per image
  int halo_slots;       // MODE 2: halo ring depth
  int tile_w, tile_h;   // MODE 2: spatial tile of output pixels (tile_w % 4 == 0, tile_w * tile_h <= 128)
  int dw_stride;        // MODE 2: stride of the depthwise stage (1 or 2); the output tile is tile_w x tile_h, the halo covers
                        // (tile - 1) * stride + KS input pixels per dimension
  int halo_tx;          // MODE 1 + TMA: exact bytes of one halo box (the slot stride halo_bytes is rounded up to 128)
  int halo_w, halo_pix, halo_bytes;   // MODE 2: (tile_w + KS - 1) x (tile_h + KS - 1) input pixels, 128 B per pixel and K-slab
  int prod_warps;       // producer warps (8; 4 when MODE 0 runs with TMA-loaded A tiles and two epilogue groups)
  int epi2;             // MODE 0 + TMA: warps 4-7 form a second epilogue group; group g drains TMEM accumulator g (every other tile)
  int stg_stride;       // bytes of epilogue staging per warp (TC_STG_BYTES; the smem-starved MODE 3 packs them at 4608)
  int tma_out;          // epilogue writes each warp's 32 x 32 block through a swizzled staging tile + TMA tensor store
  int tma_a;            // MODE 0: the A tile (128 rows x 32 k, SWIZZLE_128B) is loaded by TMA; producers only derive lo
  int tap_tma;          // MODE 1, stride 1, per-tap padded weight image: every K-slab of A is one TMA box of the NHWC input (32 channels
                        // x tile_w x tile_h pixels, shifted by the tap, zero-filled outside = the padding); spt = slabs per tap
  int spt;
  int wstream;          // MODE 2: the weight image does not fit next to the halo ring: each K-slab of W (hi, lo) is
                        // streamed from L2 into the A stage's own W slot with cp.async.bulk
  int dense_epi;       // 1: epilogue stages whole [32][N] warp slabs in smem and writes them as one aligned span
  long long M;
  int num_tiles;
  unsigned long long* dbg;   // mapped host memory: post-mortem word written before a protocol-error trap (engine.cu: debug_words)
};

struct TcLaunch {
  TcParams p;
  CUtensorMap tmap, omap;
  dim3 grid;
  size_t smem;
  int ks;              // depthwise kernel size of MODE 2
};
unsigned long long* debug_words();      // 8 words of mapped pinned host memory (process-wide), zero when no kernel trapped
int tc_prepare(const ConvParams& c, const float* wimg, int mode, int sm_count, TcLaunch* L);
int tc_launch(const TcLaunch& L, cudaStream_t st, int pdl);

struct Stem2Params {
  const float* __restrict__ in;        // [B,3,H,W] fp32 NCHW
  const float* __restrict__ wimg;      // bf16 image: [9 taps][3 splits][N2][32] SW64 | [3 splits][32][32] SW64
  const float* __restrict__ bias2;     // [Cout]
  float* __restrict__ out;             // [B,Ho,Wo,Cout] NHWC
  int B, H, W, Hs, Ws, Ho, Wo, Cout, N2, act;
  int tiles_x, tiles_y, num_tiles;
  int a1_stages;                       // stem operand stages: 2, or 1 when the 32-channel conv2 weights leave no room for two
  const unsigned char* __restrict__ in_u8;   // image mode: [B,H,W,3] uint8 BGR (the reference's cv2 image, tools/infer.py:436-453) read
                                       // directly: (u/255 - mean)/std is affine in the integer u, so it is folded into the stem
                                       // weights; u8 values are exact in ONE bf16 -> GEMM1 takes one instruction per k-step
  const float* __restrict__ pw;        // optional fused pointwise conv after conv2 (timm blocks.0.1): [Cout][Cout] weights (k-major,
                                       // BN folded) followed by Cout biases; needs Cout == N2 == 16.  nullptr: none
  int pw_act;
};

struct Stem2Launch {
  Stem2Params p;
  CUtensorMap tmap;
  int grid;
  size_t smem;
};
int stem2_prepare(const ConvParams& c, const float* wimg, int sm_count, Stem2Launch* L);
int stem2_launch(const Stem2Launch& L, cudaStream_t st, int pdl);
int stem2_row_pixel(int q);            // GEMM1 row q of the fused stem -> halo pixel hy | hx << 8 (0xFFFF: padding row, -1: out of range)

}  // namespace yl
