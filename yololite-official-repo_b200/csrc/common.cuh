// Shared helpers for the yololite_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/yololite_b200.h"

namespace yl {

void set_error(const std::string& msg);
extern std::atomic<long long> g_tc_launches, g_simt_launches, g_post_launches;

#define YL_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      yl::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +      \
                    __FILE__ + ":" + std::to_string(__LINE__));                               \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)

#define YL_REQUIRE(cond, msg)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      yl::set_error(std::string("invalid argument: ") + (msg) + " [" #cond "]");              \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == YL_ACT_RELU) return fmaxf(v, 0.f);
  if (act == YL_ACT_SILU) return v / (1.f + __expf(-v));
  return v;
}

// accurate SiLU used where the epilogue feeds the 1e-3 logit gate: x * sigmoid(x) with libm expf.
__device__ __forceinline__ float silu_accurate(float v) { return v / (1.f + expf(-v)); }

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == YL_ACT_RELU) return fmaxf(v, 0.f);
  if (act == YL_ACT_SILU) return silu_accurate(v);
  return v;
}

// PyTorch "nearest" source index for F.interpolate(size=...) (model_v2.py:337-338):
// src = min(floor(dst * (in/out)), in-1), scale computed in fp32.
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  if (in_size * 2 == out_size) return dst >> 1;
  float scale = (float)in_size / (float)out_size;
  int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

struct ConvParams {
  const float* __restrict__ in;     // NHWC (or NCHW for the stem)
  const unsigned char* __restrict__ in_u8;   // YL_OP_STEM2 image mode: [B,H,W,3] uint8 BGR instead of `in`
  const float* __restrict__ w;      // [K][N], k = (ky*KS + kx)*Cin + ci
  const float* __restrict__ bias;   // [N] or nullptr
  const float* __restrict__ res;    // [M][N] or nullptr
  const float* __restrict__ up;     // [B][Hu][Wu][N] or nullptr
  const float* __restrict__ w2;     // DWPW: depthwise weights [k2*k2][Cin]
  const float* __restrict__ b2;     // DWPW: depthwise bias [Cin] or nullptr
  float* __restrict__ out;
  int B, Hin, Win, Cin, Hout, Wout, Cout;
  int KS, stride, pad;
  int Hu, Wu;
  int act;
  int act2;                          // DWPW: activation between the depthwise and the pointwise stage
  int anchors;                       // >0: head layout [B,A,H,W,D], D = Cout/anchors
  int wt_layout;                     // yl_op.wt_layout of the tcgen05 weight image (1: every tap padded to 32 channels)
};

// TMA descriptor over an fp32 tensor (tc_gemm.cu): dims / box innermost first, strides in bytes for dims 1..rank-1;
// out-of-bounds elements read as zero.  cuTensorMapEncodeTiled is resolved through the runtime (no libcuda link).
int make_tmap_f32(CUtensorMap* tm, const float* base, int rank, const unsigned long long* dims, const unsigned long long* strides,
                  const unsigned int* box, bool swizzle128);

// launchers (conv_kernels.cu)
int launch_stem(const ConvParams& p, cudaStream_t s);
int launch_conv_gemm(const ConvParams& p, cudaStream_t s);
int launch_dw(const ConvParams& p, cudaStream_t s);
int launch_dwpw(const ConvParams& p, cudaStream_t s);

}  // namespace yl
