// fp32 SIMT convolution kernels on NHWC activations (sm_100a).
//
//   stem_kernel       conv_stem 3x3 s2 on the NCHW network input  (timm backbone, model_v2.py:353)
//   conv_gemm_kernel  dense KxK conv as implicit GEMM; K=1 is every pointwise conv of the backbone,
//                     FPN laterals (model_v2.py:289-291,359-361), DWConvBlock 1x1 (:23-39) and the
//                     box/obj/cls output convs (:42-53,:340-350) written straight in [B,A,S,S,5+C]
//   dw_kernel         depthwise 3x3 / 5x5, stride 1 / 2
//
// All accumulate in fp32 FMA: the 1e-3 logit gate rules out single-pass TF32/BF16 operands
// (SURVEY.md section 0 fact 5).
#include "common.cuh"

namespace yl {

// ------------------------------------------------------------------------------------------------
// stem: one thread per output pixel, all (<=32 per grid.y slice) output channels in registers
// ------------------------------------------------------------------------------------------------
constexpr int STEM_CO = 32;
constexpr int STEM_THREADS = 128;

__global__ void __launch_bounds__(STEM_THREADS) stem_kernel(ConvParams p) {
  __shared__ __align__(16) float ws[27 * STEM_CO];
  __shared__ __align__(16) float bs[STEM_CO];
  const int co0 = blockIdx.y * STEM_CO;
  const int KK = p.KS * p.KS * p.Cin;  // 27
  for (int i = threadIdx.x; i < KK * STEM_CO; i += STEM_THREADS) {
    int k = i / STEM_CO, c = i - k * STEM_CO;
    ws[i] = (co0 + c < p.Cout) ? p.w[(size_t)k * p.Cout + co0 + c] : 0.f;
  }
  if (threadIdx.x < STEM_CO)
    bs[threadIdx.x] = (p.bias && co0 + threadIdx.x < p.Cout) ? p.bias[co0 + threadIdx.x] : 0.f;
  __syncthreads();

  const long long M = (long long)p.B * p.Hout * p.Wout;
  const long long m = (long long)blockIdx.x * STEM_THREADS + threadIdx.x;
  if (m >= M) return;
  const int ox = (int)(m % p.Wout);
  const int oy = (int)((m / p.Wout) % p.Hout);
  const int b = (int)(m / ((long long)p.Wout * p.Hout));

  float acc[STEM_CO];
#pragma unroll
  for (int c = 0; c < STEM_CO; ++c) acc[c] = bs[c];

  const size_t plane = (size_t)p.Hin * p.Win;
  const float* inb = p.in + (size_t)b * p.Cin * plane;
  for (int ky = 0; ky < p.KS; ++ky) {
    const int iy = oy * p.stride + ky - p.pad;
    if (iy < 0 || iy >= p.Hin) continue;
    for (int kx = 0; kx < p.KS; ++kx) {
      const int ix = ox * p.stride + kx - p.pad;
      if (ix < 0 || ix >= p.Win) continue;
      for (int ci = 0; ci < p.Cin; ++ci) {
        const float v = __ldg(inb + ci * plane + (size_t)iy * p.Win + ix);
        const float4* wr = reinterpret_cast<const float4*>(&ws[((ky * p.KS + kx) * p.Cin + ci) * STEM_CO]);
#pragma unroll
        for (int q = 0; q < STEM_CO / 4; ++q) {
          const float4 w4 = wr[q];
          acc[q * 4 + 0] = fmaf(v, w4.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(v, w4.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(v, w4.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(v, w4.w, acc[q * 4 + 3]);
        }
      }
    }
  }
  float* o = p.out + (size_t)m * p.Cout + co0;
  if ((p.Cout & 3) == 0 && co0 + STEM_CO <= p.Cout) {
#pragma unroll
    for (int q = 0; q < STEM_CO / 4; ++q) {
      float4 v;
      v.x = act_fn(acc[q * 4 + 0], p.act);
      v.y = act_fn(acc[q * 4 + 1], p.act);
      v.z = act_fn(acc[q * 4 + 2], p.act);
      v.w = act_fn(acc[q * 4 + 3], p.act);
      reinterpret_cast<float4*>(o)[q] = v;
    }
  } else {
#pragma unroll
    for (int c = 0; c < STEM_CO; ++c)
      if (co0 + c < p.Cout) o[c] = act_fn(acc[c], p.act);
  }
}

int launch_stem(const ConvParams& p, cudaStream_t s) {
  YL_REQUIRE(p.KS * p.KS * p.Cin <= 27, "stem supports KxKxCin <= 27");
  const long long M = (long long)p.B * p.Hout * p.Wout;
  dim3 grid((unsigned)((M + STEM_THREADS - 1) / STEM_THREADS), (p.Cout + STEM_CO - 1) / STEM_CO);
  stem_kernel<<<grid, STEM_THREADS, 0, s>>>(p);
  YL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM conv: C[M,N] = im2col(A)[M,K] * W[K,N],  M = B*Hout*Wout, K = KS*KS*Cin, N = Cout
// 128 x (16*TN) tile per CTA, 256 threads, 8 x TN outputs per thread, BK = 16, register prefetch.
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128, BK = 16, GEMM_THREADS = 256, TM = 8;

// MODE 0: pointwise (A rows are contiguous), 1: generic KxK im2col gather,
// MODE 2: fused depthwise -> pointwise -- A[m][c] = act2(depthwiseKxK(in)[m][c] + b2[c]) computed while loading
//         (DWConvBlock model_v2.py:23-39; dw_start/dw_mid -> pointwise pairs of the backbone's UIR blocks)
template <int TN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS) conv_gemm_kernel(ConvParams p, int ldw) {
  constexpr int BN = 16 * TN;
  constexpr int WV = (BK * BN / 4 + GEMM_THREADS - 1) / GEMM_THREADS;  // float4 W loads per thread
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long M = (long long)p.B * p.Hout * p.Wout;
  const int N = p.Cout, K = MODE == 1 ? p.KS * p.KS * p.Cin : p.Cin;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A loader state: two rows per thread, one k-quad
  const int kq = tid & 3;
  const float* a_base[2];
  int a_oy[2], a_ox[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long m = m0 + (tid >> 2) + i * 64;
    a_ok[i] = m < M;
    const long long mm = a_ok[i] ? m : 0;
    if (MODE == 0) {
      a_base[i] = p.in + (size_t)mm * p.Cin;
      a_oy[i] = a_ox[i] = 0;
    } else {
      const int ox = (int)(mm % p.Wout);
      const int oy = (int)((mm / p.Wout) % p.Hout);
      const int b = (int)(mm / ((long long)p.Wout * p.Hout));
      a_base[i] = p.in + (size_t)b * p.Hin * p.Win * p.Cin;
      a_oy[i] = oy * p.stride - p.pad;
      a_ox[i] = ox * p.stride - p.pad;
    }
  }

  float4 a_reg[2];
  float4 w_reg[WV];

  auto load_tile = [&](int kt) {
    const int k = kt * BK + kq * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i] && k < K) {
        if (MODE == 0) {
          v = __ldg(reinterpret_cast<const float4*>(a_base[i] + k));
        } else if (MODE == 2) {
          if (p.b2) v = __ldg(reinterpret_cast<const float4*>(p.b2 + k));
          for (int ky = 0; ky < p.KS; ++ky) {
            const int iy = a_oy[i] + ky;
            if (iy < 0 || iy >= p.Hin) continue;
            for (int kx = 0; kx < p.KS; ++kx) {
              const int ix = a_ox[i] + kx;
              if (ix < 0 || ix >= p.Win) continue;
              const float4 x4 = __ldg(reinterpret_cast<const float4*>(a_base[i] + ((size_t)iy * p.Win + ix) * p.Cin + k));
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w2 + (ky * p.KS + kx) * p.Cin + k));
              v.x = fmaf(x4.x, w4.x, v.x);
              v.y = fmaf(x4.y, w4.y, v.y);
              v.z = fmaf(x4.z, w4.z, v.z);
              v.w = fmaf(x4.w, w4.w, v.w);
            }
          }
          v.x = act_fn(v.x, p.act2); v.y = act_fn(v.y, p.act2); v.z = act_fn(v.z, p.act2); v.w = act_fn(v.w, p.act2);
        } else {
          const int tap = k / p.Cin, ci = k - tap * p.Cin;
          const int ky = tap / p.KS, kx = tap - ky * p.KS;
          const int iy = a_oy[i] + ky, ix = a_ox[i] + kx;
          if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win)
            v = __ldg(reinterpret_cast<const float4*>(a_base[i] + ((size_t)iy * p.Win + ix) * p.Cin + ci));
        }
      }
      a_reg[i] = v;
    }
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int idx = tid + i * GEMM_THREADS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BK * BN / 4) {
        const int kk = idx / (BN / 4), nn = (idx - kk * (BN / 4)) * 4;
        const int kg = kt * BK + kk, ng = n0 + nn;
        if (kg < K && ng < ldw) v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)kg * ldw + ng));
      }
      w_reg[i] = v;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = (tid >> 2) + i * 64;
      As[kq * 4 + 0][r] = a_reg[i].x;
      As[kq * 4 + 1][r] = a_reg[i].y;
      As[kq * 4 + 2][r] = a_reg[i].z;
      As[kq * 4 + 3][r] = a_reg[i].w;
    }
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int idx = tid + i * GEMM_THREADS;
      if (idx < BK * BN / 4) {
        const int kk = idx / (BN / 4), nn = (idx - kk * (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Ws[kk][nn]) = w_reg[i];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int numK = (K + BK - 1) / BK;
  load_tile(0);
  store_tile();
  __syncthreads();
  for (int kt = 0; kt < numK; ++kt) {
    if (kt + 1 < numK) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float w[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = Ws[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
    if (kt + 1 < numK) {
      store_tile();
      __syncthreads();
    }
  }

  // ---- epilogue
  float bias[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + tx * TN + j;
    bias[j] = (p.bias && n < N) ? __ldg(p.bias + n) : 0.f;
  }
  const int D = p.anchors > 0 ? N / p.anchors : N;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    const float* up_row = nullptr;
    int b = 0, oy = 0, ox = 0;
    if (p.up || p.anchors > 1) {
      ox = (int)(m % p.Wout);
      oy = (int)((m / p.Wout) % p.Hout);
      b = (int)(m / ((long long)p.Wout * p.Hout));
      if (p.up) {
        const int sy = nearest_src(oy, p.Hu, p.Hout), sx = nearest_src(ox, p.Wu, p.Wout);
        up_row = p.up + (((size_t)b * p.Hu + sy) * p.Wu + sx) * N;
      }
    }
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      float t = acc[i][j] + bias[j];
      if (n < N) {
        if (p.res) t += __ldg(p.res + (size_t)m * N + n);
        if (up_row) t += __ldg(up_row + n);
      }
      v[j] = act_fn(t, p.act);
    }
    if (p.anchors > 1) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < N) {
          const int a = n / D, d = n - a * D;
          p.out[((((size_t)b * p.anchors + a) * p.Hout + oy) * p.Wout + ox) * D + d] = v[j];
        }
      }
    } else if ((TN % 4) == 0 && (N & 3) == 0) {
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const int n = n0 + tx * TN + j;
        if (n < N)
          *reinterpret_cast<float4*>(p.out + (size_t)m * N + n) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    } else if ((TN % 2) == 0 && (N & 1) == 0) {
#pragma unroll
      for (int j = 0; j < TN; j += 2) {
        const int n = n0 + tx * TN + j;
        if (n < N) *reinterpret_cast<float2*>(p.out + (size_t)m * N + n) = make_float2(v[j], v[j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < N) p.out[(size_t)m * N + n] = v[j];
      }
    }
  }
}

template <int TN>
static int launch_conv_gemm_tn(const ConvParams& p, int ldw, cudaStream_t s) {
  const long long M = (long long)p.B * p.Hout * p.Wout;
  dim3 grid((unsigned)((M + BM - 1) / BM), (p.Cout + 16 * TN - 1) / (16 * TN));
  const bool is_pw = p.KS == 1 && p.stride == 1 && p.pad == 0;
  if (p.w2)
    conv_gemm_kernel<TN, 2><<<grid, GEMM_THREADS, 0, s>>>(p, ldw);
  else if (is_pw)
    conv_gemm_kernel<TN, 0><<<grid, GEMM_THREADS, 0, s>>>(p, ldw);
  else
    conv_gemm_kernel<TN, 1><<<grid, GEMM_THREADS, 0, s>>>(p, ldw);
  YL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_dwpw(const ConvParams& p, cudaStream_t s) {
  YL_REQUIRE(p.w2 && (p.KS == 3 || p.KS == 5) && (p.stride == 1 || p.stride == 2) && p.pad == p.KS / 2, "fused depthwise (3x3 / 5x5, s1 / s2) + pointwise");
  return launch_conv_gemm(p, s);
}

int launch_conv_gemm(const ConvParams& p, cudaStream_t s) {
  YL_REQUIRE((p.Cin & 3) == 0, "conv_gemm needs Cin % 4 == 0");
  const int N = p.Cout, ldw = (N + 3) & ~3;
  // pick the column tile that wastes the fewest padded columns (ties -> wider tile)
  const int cands[6] = {8, 6, 4, 3, 2, 1};
  int best = 8;
  long long best_cost = -1;
  for (int c : cands) {
    const int bn = 16 * c;
    const long long cost = (long long)((N + bn - 1) / bn) * bn;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = c; }
  }
  switch (best) {
    case 8: return launch_conv_gemm_tn<8>(p, ldw, s);
    case 6: return launch_conv_gemm_tn<6>(p, ldw, s);
    case 4: return launch_conv_gemm_tn<4>(p, ldw, s);
    case 3: return launch_conv_gemm_tn<3>(p, ldw, s);
    case 2: return launch_conv_gemm_tn<2>(p, ldw, s);
    default: return launch_conv_gemm_tn<1>(p, ldw, s);
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise KxK: one thread = 4 channels of one output pixel
// ------------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256) dw_kernel(ConvParams p) {
  const int C4 = p.Cin >> 2;
  const long long total = (long long)p.B * p.Hout * p.Wout * C4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const long long pix = idx / C4;
  const int ox = (int)(pix % p.Wout);
  const int oy = (int)((pix / p.Wout) % p.Hout);
  const int b = (int)(pix / ((long long)p.Wout * p.Hout));
  const float4* in = reinterpret_cast<const float4*>(p.in) + (size_t)b * p.Hin * p.Win * C4 + c4;
  const float4* w = reinterpret_cast<const float4*>(p.w) + c4;
  float4 acc = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
#pragma unroll
  for (int ky = 0; ky < KS; ++ky) {
    const int iy = iy0 + ky;
    if (iy < 0 || iy >= p.Hin) continue;
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      const int ix = ix0 + kx;
      if (ix < 0 || ix >= p.Win) continue;
      const float4 v = __ldg(in + ((size_t)iy * p.Win + ix) * C4);
      const float4 k = __ldg(w + (ky * KS + kx) * C4);
      acc.x = fmaf(v.x, k.x, acc.x);
      acc.y = fmaf(v.y, k.y, acc.y);
      acc.z = fmaf(v.z, k.z, acc.z);
      acc.w = fmaf(v.w, k.w, acc.w);
    }
  }
  acc.x = act_fn(acc.x, p.act);
  acc.y = act_fn(acc.y, p.act);
  acc.z = act_fn(acc.z, p.act);
  acc.w = act_fn(acc.w, p.act);
  reinterpret_cast<float4*>(p.out)[idx] = acc;
}

int launch_dw(const ConvParams& p, cudaStream_t s) {
  YL_REQUIRE((p.Cin & 3) == 0 && p.Cin == p.Cout, "depthwise needs Cin == Cout, Cin % 4 == 0");
  const long long total = (long long)p.B * p.Hout * p.Wout * (p.Cin >> 2);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (p.KS == 3)
    dw_kernel<3><<<grid, 256, 0, s>>>(p);
  else if (p.KS == 5)
    dw_kernel<5><<<grid, 256, 0, s>>>(p);
  else
    YL_REQUIRE(false, "depthwise kernel size must be 3 or 5");
  YL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace yl
