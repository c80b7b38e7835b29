// Experiment: systematic (toward-zero) error of tcgen05 fp32 accumulation for three error-compensated operand schemes.
//   D[128][64] = A[128][K] x W[K][64], one CTA, K streamed in 32-wide slabs through shared memory.
//   scheme 0: 3xTF32   main += Ahi*Whi (kind::tf32, K=8/instr); corr += Alo*Whi + Ahi*Wlo
//   scheme 1: bf16x3   main += A1*W1 (kind::f16 bf16, K=16/instr); corr += A1*W2 + A2*W1 + A2*W2 + A1*W3 + A3*W1
//   scheme 2: fp16x2   main += Ah*Wh (kind::f16 fp16, K=16/instr); corr += Al*Wh + Ah*Wl
// P = number of main accumulators (contiguous K ranges), summed in fp32 (round-to-nearest) by the epilogue.
// Reports the mean signed error relative to the exact (fp64) result in units of 2^-24, toward zero = negative.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct P {
  const float* A; const float* W; float* out;
  int K, scheme, chains;
};

constexpr int M = 128, N = 64;

__device__ __forceinline__ void put(unsigned char* base, int row, int kk, int rowB, int bytes, float v, int scheme) {
  const uint32_t rowoff = row * rowB;
  const int chunk = (kk * bytes) >> 4, within = (kk * bytes) & 15;
  const int mask = rowB == 64 ? 3 : 7;
  unsigned char* dst = base + rowoff + ((chunk ^ ((rowoff >> 7) & mask)) << 4) + within;
  if (scheme == 0) *reinterpret_cast<float*>(dst) = v;
  else if (scheme == 1) *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(v);
  else *reinterpret_cast<__half*>(dst) = __float2half(v);
}

__device__ __forceinline__ void mma(int scheme, uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (scheme == 0)
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(P p) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  const int scheme = p.scheme;
  const int rowB = scheme == 0 ? 128 : 64, bytes = scheme == 0 ? 4 : 2;
  const int nsplit = scheme == 1 ? 3 : 2;
  unsigned char* Ap[3];
  unsigned char* Bp[3];
  for (int s = 0; s < 3; ++s) { Ap[s] = smem + s * 16384; Bp[s] = smem + 49152 + s * 8192; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  const int nslab = p.K / 32;
  uint32_t phase = 0;
  uint32_t used = 0;                                  // bit c: accumulator c has been written (thread 0 only)
  const uint32_t fmt = scheme == 0 ? 2u : scheme == 1 ? 1u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint64_t layout = scheme == 0 ? 2ull : 4ull;
  for (int slab = 0; slab < nslab; ++slab) {
    // ---- operand planes of this slab
    for (int i = tid; i < M * 32; i += 128) {
      const int r = i >> 5, kk = i & 31;
      const float a = p.A[(size_t)r * p.K + slab * 32 + kk];
      if (scheme == 0) {
        put(Ap[0], r, kk, rowB, bytes, a, 0);
        put(Ap[1], r, kk, rowB, bytes, a - __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), 0);
      } else if (scheme == 1) {
        const float a1 = __bfloat162float(__float2bfloat16(a)), r1 = a - a1;
        const float a2 = __bfloat162float(__float2bfloat16(r1));
        put(Ap[0], r, kk, rowB, bytes, a1, 1); put(Ap[1], r, kk, rowB, bytes, a2, 1); put(Ap[2], r, kk, rowB, bytes, r1 - a2, 1);
      } else {
        const float h = __half2float(__float2half(a));
        put(Ap[0], r, kk, rowB, bytes, h, 2); put(Ap[1], r, kk, rowB, bytes, a - h, 2);
      }
    }
    for (int i = tid; i < N * 32; i += 128) {
      const int n = i >> 5, kk = i & 31;
      const float w = p.W[(size_t)(slab * 32 + kk) * N + n];
      if (scheme == 0) {
        const float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xFFFFE000u), lo0 = w - hi;
        put(Bp[0], n, kk, rowB, bytes, hi, 0);
        put(Bp[1], n, kk, rowB, bytes, __uint_as_float((__float_as_uint(lo0) + 0x1000u) & 0xFFFFE000u), 0);
      } else if (scheme == 1) {
        const float w1 = __bfloat162float(__float2bfloat16(w)), r1 = w - w1;
        const float w2 = __bfloat162float(__float2bfloat16(r1));
        put(Bp[0], n, kk, rowB, bytes, w1, 1); put(Bp[1], n, kk, rowB, bytes, w2, 1); put(Bp[2], n, kk, rowB, bytes, r1 - w2, 1);
      } else {
        const float h = __half2float(__float2half(w));
        put(Bp[0], n, kk, rowB, bytes, h, 2); put(Bp[1], n, kk, rowB, bytes, w - h, 2);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int chain = (slab * p.chains) / nslab;
      const uint32_t dmain = tm + chain * N, dcorr = tm + 4 * N;
      const int ksteps = scheme == 0 ? 4 : 2;
      auto desc = [&](unsigned char* base, uint32_t ko) {
        return (uint64_t)(((smem_u32(base) + ko) & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * rowB) >> 4) << 32) | (1ull << 46) | (layout << 61);
      };
      auto go = [&](uint32_t d, int acc_id, unsigned char* a, unsigned char* b, uint32_t ko) {
        mma(scheme, d, desc(a, ko), desc(b, ko), idesc, (used >> acc_id) & 1u);
        used |= 1u << acc_id;
      };
      for (int j = 0; j < ksteps; ++j) {
        const uint32_t ko = j * 32;
        if (nsplit == 2) {
          go(dcorr, 4, Ap[1], Bp[0], ko);
          go(dcorr, 4, Ap[0], Bp[1], ko);
          go(dmain, chain, Ap[0], Bp[0], ko);
        } else {
          go(dcorr, 4, Ap[0], Bp[1], ko); go(dcorr, 4, Ap[1], Bp[0], ko); go(dcorr, 4, Ap[1], Bp[1], ko);
          go(dcorr, 4, Ap[0], Bp[2], ko); go(dcorr, 4, Ap[2], Bp[0], ko);
          go(dmain, chain, Ap[0], Bp[0], ko);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
      uint32_t ok = 0;
      for (int spin = 0; !ok && spin < (1 << 24); ++spin)
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
      if (!ok) asm volatile("trap;");
      phase ^= 1u;
    }
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t ta = tm + ((uint32_t)((tid >> 5) * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    float sum[16];
    for (int j = 0; j < 16; ++j) sum[j] = 0.f;
    for (int acc = 4; acc >= 0; --acc) {                    // corr first (small), then the main chains
      if (acc < 4 && acc >= p.chains) continue;
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta + acc * N + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(r[j]);
    }
    for (int j = 0; j < 16; ++j) p.out[tid * N + c0 + j] = sum[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  const char* names[3] = {"3xTF32", "bf16x3", "fp16x2"};
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  float *dA, *dW, *dO;
  cudaMalloc(&dA, M * 960 * 4); cudaMalloc(&dW, 960 * N * 4); cudaMalloc(&dO, M * N * 4);
  for (int data = 0; data < 2; ++data)
    for (int K : {96, 256, 960}) {
      std::mt19937 g(7 + K);
      std::normal_distribution<float> nd(0.f, 1.f);
      std::vector<float> A(M * K), W(K * N), O(M * N);
      for (auto& v : A) { float x = nd(g); v = data == 0 ? fabsf(x) : fmaxf(x, 0.f); }          // positive / ReLU-like (half zeros)
      for (auto& v : W) { float x = nd(g) / sqrtf((float)K); v = data == 0 ? fabsf(x) : x; }
      cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
      std::vector<double> ex(M * N);
      double rms = 0;
      for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int kk = 0; kk < K; ++kk) s += (double)A[r * K + kk] * (double)W[kk * N + n];
          ex[r * N + n] = s; rms += s * s;
        }
      rms = sqrt(rms / (M * N));
      for (int scheme = 0; scheme < 3; ++scheme)
        for (int chains : {1, 2, 4}) {
          P p{dA, dW, dO, K, scheme, chains};
          k<<<1, 128, 100 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
          double mean = 0, sq = 0, mx = 0;
          for (int i = 0; i < M * N; ++i) {
            const double sgn = ex[i] >= 0 ? 1.0 : -1.0;
            const double rel = ((double)O[i] - ex[i]) * sgn / (data == 0 ? fabs(ex[i]) : rms);
            mean += rel; sq += rel * rel; mx = fmax(mx, fabs(rel));
          }
          mean /= M * N;
          printf("data=%s K=%4d %s chains=%d: mean signed err %+8.2f ulp  rms %7.2f ulp  max %7.2f ulp\n", data == 0 ? "positive" : "relu*rand",
                 K, names[scheme], chains, mean * 16777216.0, sqrt(sq / (M * N)) * 16777216.0, mx * 16777216.0);
        }
    }
  return 0;
}
