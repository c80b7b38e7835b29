// Letterbox + BGR->RGB + normalise + HWC->CHW in one pass from a uint8 image (sm_100a).
//
// Restates tools/infer.py:121-131 (letterbox: cv2.resize INTER_LINEAR to (nw,nh), 114-padding) and
// :442-453 (BGR2RGB, /255, (x-mean)/std, CHW).  The resize reproduces OpenCV's 8-bit INTER_LINEAR
// fixed-point arithmetic bit for bit (11-bit coefficients from fp32 fractions, horizontal pass in int,
// vertical pass ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2; x fractions are clamped at the borders,
// y rows are clipped instead) -- pinned against cv2 itself in tests/test_oracle_pre.py.
#include "common.cuh"

namespace yl {

struct PreParams {
  const unsigned char* src;
  float* dst;
  int h0, w0, pitch, S, nh, nw, left, top;
  size_t src_stride, dst_stride;   // per image (batched launch: blockIdx.z)
};

__device__ __forceinline__ void lin_coef(int d, int n_src, double scale, bool clamp, int& i0, int& i1, int& a0, int& a1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
  }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  i0 = min(max(s, 0), n_src - 1);
  i1 = min(max(s + 1, 0), n_src - 1);
}

__global__ void __launch_bounds__(256) pre_kernel(PreParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.S) return;
  p.src += (size_t)blockIdx.z * p.src_stride;
  p.dst += (size_t)blockIdx.z * p.dst_stride;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  int bgr[3] = {114, 114, 114};
  const int ry = y - p.top, rx = x - p.left;
  if (ry >= 0 && ry < p.nh && rx >= 0 && rx < p.nw) {
    if (p.nh == p.h0 && p.nw == p.w0) {
      const unsigned char* s = p.src + (size_t)ry * p.pitch + rx * 3;
      bgr[0] = s[0]; bgr[1] = s[1]; bgr[2] = s[2];
    } else {
      const double sx = 1.0 / ((double)p.nw / (double)p.w0), sy = 1.0 / ((double)p.nh / (double)p.h0);
      int x0, x1, ax0, ax1, y0, y1, ay0, ay1;
      lin_coef(rx, p.w0, sx, true, x0, x1, ax0, ax1);
      lin_coef(ry, p.h0, sy, false, y0, y1, ay0, ay1);
      const unsigned char* r0 = p.src + (size_t)y0 * p.pitch;
      const unsigned char* r1 = p.src + (size_t)y1 * p.pitch;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = r0[x0 * 3 + c] * ax0 + r0[x1 * 3 + c] * ax1;
        const int h1 = r1[x0 * 3 + c] * ax0 + r1[x1 * 3 + c] * ax1;
        int v = ((((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2);
        bgr[c] = min(max(v, 0), 255);
      }
    }
  }
  const size_t plane = (size_t)p.S * p.S;
  float* o = p.dst + (size_t)y * p.S + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {   // output channel c = RGB -> source channel 2-c
    const float v = __fdiv_rn((float)bgr[2 - c], 255.f);
    o[c * plane] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
  }
}

// Fast path when the letterbox does not resize (nh == h0, nw == w0): a thread converts 4 horizontally adjacent output pixels.
// The 256 possible byte values per channel go through a shared-memory table built with the SAME IEEE operations
// (v / 255, then (v - mean) / std), so the result is bit-identical to the general kernel and to the reference's fp32 math.
__global__ void __launch_bounds__(256) pre_kernel_copy4(PreParams p) {
  __shared__ float lut[3][256];
  {
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn((float)threadIdx.x, 255.f);
      lut[c][threadIdx.x] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
    }
  }
  __syncthreads();
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  if (x >= p.S) return;
  p.src += (size_t)blockIdx.z * p.src_stride;
  p.dst += (size_t)blockIdx.z * p.dst_stride;
  const int ry = y - p.top;
  unsigned char px[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rx = x + i - p.left;
    if (ry >= 0 && ry < p.nh && rx >= 0 && rx < p.nw) {
      const unsigned char* s = p.src + (size_t)ry * p.pitch + rx * 3;
      px[i][0] = s[0]; px[i][1] = s[1]; px[i][2] = s[2];
    } else {
      px[i][0] = px[i][1] = px[i][2] = 114;
    }
  }
  const size_t plane = (size_t)p.S * p.S;
  float* o = p.dst + (size_t)y * p.S + x;
#pragma unroll
  for (int c = 0; c < 3; ++c)      // output channel c = RGB -> source channel 2-c
    *reinterpret_cast<float4*>(o + c * plane) = make_float4(lut[c][px[0][2 - c]], lut[c][px[1][2 - c]], lut[c][px[2][2 - c]], lut[c][px[3][2 - c]]);
}

static void launch_pre(const PreParams& p, int B, cudaStream_t st) {
  if (p.nh == p.h0 && p.nw == p.w0 && (p.S & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0 && (p.dst_stride & 3) == 0) {
    dim3 grid((p.S / 4 + 255) / 256, p.S, B);
    pre_kernel_copy4<<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((p.S + 255) / 256, p.S, B);
    pre_kernel<<<grid, 256, 0, st>>>(p);
  }
}

}  // namespace yl

extern "C" int yl_preprocess(const uint8_t* src, int32_t h0, int32_t w0, int32_t pitch, float* dst, int32_t S, int32_t nh,
                             int32_t nw, int32_t left, int32_t top, void* stream) {
  using namespace yl;
  YL_REQUIRE(src && dst, "null pointer");
  YL_REQUIRE(h0 >= 1 && w0 >= 1 && pitch >= w0 * 3 && S >= 1, "bad image geometry");
  YL_REQUIRE(nh >= 1 && nw >= 1 && left >= 0 && top >= 0 && left + nw <= S && top + nh <= S, "letterbox does not fit");
  PreParams p{src, dst, h0, w0, pitch, S, nh, nw, left, top, 0, 0};
  launch_pre(p, 1, reinterpret_cast<cudaStream_t>(stream));
  YL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int yl_preprocess_batch(const uint8_t* src, int32_t B, int32_t h0, int32_t w0, float* dst, int32_t S, int32_t nh,
                                   int32_t nw, int32_t left, int32_t top, void* stream) {
  using namespace yl;
  YL_REQUIRE(src && dst && B >= 1, "null pointer / empty batch");
  YL_REQUIRE(h0 >= 1 && w0 >= 1 && S >= 1, "bad image geometry");
  YL_REQUIRE(nh >= 1 && nw >= 1 && left >= 0 && top >= 0 && left + nw <= S && top + nh <= S, "letterbox does not fit");
  PreParams p{src, dst, h0, w0, w0 * 3, S, nh, nw, left, top, (size_t)h0 * w0 * 3, (size_t)3 * S * S};
  launch_pre(p, B, reinterpret_cast<cudaStream_t>(stream));
  YL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
