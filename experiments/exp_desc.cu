// Experiment: does a tcgen05 shared-memory descriptor whose start address / SBO are NOT multiples of the swizzle
// atom (1024 B for SWIZZLE_128B, 512 B for SWIZZLE_64B) read data that was written with ADDRESS-based swizzling?
// A "plane" of pixels (pitch PW pixels, one pixel = one K-row of 64 B bf16 or 128 B tf32) is the A operand of an
// M=128 MMA whose row r = oy*8 + ox maps to plane pixel (oy + dy, ox + dx).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct P { int mode;  /*0: bf16 SW64 (64 B rows), 1: tf32 SW128 (128 B rows)*/ int PW, dx, dy; float* out; };

__global__ void __launch_bounds__(128, 1) k(P p) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int rowB = p.mode == 0 ? 64 : 128;
  const int swz_mask = p.mode == 0 ? 3 : 7;
  unsigned char* A = smem;                   // plane: 40 rows x PW pixels
  unsigned char* B = smem + 96 * 1024;       // 16 rows x rowB, canonical, aligned
  const int tid = threadIdx.x;
  // ---- fill the plane with address-based swizzle: pixel (y,x) holds K values f(y,x,k)
  const int nk = p.mode == 0 ? 32 : 32;      // K = 32 elements either way (bf16: 64 B, tf32: 128 B)
  for (int pix = tid; pix < 40 * p.PW; pix += 128) {
    const int y = pix / p.PW, x = pix % p.PW;
    const uint32_t rowaddr = smem_u32(A) + pix * rowB;
    for (int kk = 0; kk < nk; ++kk) {
      const float v = (float)((y * 7 + x * 3 + kk * 5) % 17 - 8);
      const int bytes = p.mode == 0 ? 2 : 4;
      const int chunk = (kk * bytes) >> 4, within = (kk * bytes) & 15;
      const uint32_t addr = rowaddr + (((chunk ^ ((rowaddr >> 7) & swz_mask))) << 4) + within;
      unsigned char* dst = A + (addr - smem_u32(A));
      if (p.mode == 0) *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(v);
      else *reinterpret_cast<float*>(dst) = v;
    }
  }
  for (int i = tid; i < 16 * nk; i += 128) {
    const int n = i / nk, kk = i % nk;
    const float v = (float)((n * 5 + kk * 11) % 13 - 6);
    const int bytes = p.mode == 0 ? 2 : 4;
    const int chunk = (kk * bytes) >> 4, within = (kk * bytes) & 15;
    const uint32_t rowoff = n * rowB;
    unsigned char* dst = B + rowoff + ((chunk ^ ((rowoff >> 7) & swz_mask)) << 4) + within;
    if (p.mode == 0) *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(v);
    else *reinterpret_cast<float*>(dst) = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  if (tid == 0) {
    const uint64_t layout = p.mode == 0 ? 4ull : 2ull;
    const uint32_t fmt = p.mode == 0 ? 1u : 2u;   // BF16 / TF32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = smem_u32(A) + (p.dy * p.PW + p.dx) * rowB;
    const uint32_t b0 = smem_u32(B);
    const uint32_t sboA = p.PW * rowB, sboB = 8 * rowB;
    const int ksteps = p.mode == 0 ? 2 : 4;
    for (int j = 0; j < ksteps; ++j) {
      const uint32_t ko = j * 32;
      const uint64_t da = (uint64_t)(((a0 + ko) & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sboA >> 4) << 32) | (1ull << 46) | (layout << 61);
      const uint64_t db = (uint64_t)(((b0 + ko) & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sboB >> 4) << 32) | (1ull << 46) | (layout << 61);
      if (p.mode == 0)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                     ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(j > 0 ? 1u : 0u) : "memory");
      else
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}"
                     ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(j > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    for (int spin = 0; !ok && spin < (1 << 24); ++spin)
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    if (!ok) asm volatile("trap;");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  const uint32_t ta = tm + ((uint32_t)((tid >> 5) * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) p.out[tid * 16 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int bad_total = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int PW = 8; PW <= 9; ++PW)
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          P p{mode, PW, dx, dy, d};
          cudaMemset(d, 0xff, 128 * 16 * 4);
          k<<<1, 128, 200 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d PW %d dy %d dx %d: CUDA error %s\n", mode, PW, dy, dx, cudaGetErrorString(e)); return 1; }
          std::vector<float> h(128 * 16);
          cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 16; ++n) {
              const int y = (r >> 3) + dy, x = (r & 7) + dx;
              double acc = 0;
              for (int kk = 0; kk < 32; ++kk) acc += (double)((y * 7 + x * 3 + kk * 5) % 17 - 8) * (double)((n * 5 + kk * 11) % 13 - 6);
              if (fabs(acc - h[r * 16 + n]) > 1e-3) ++bad;
            }
          printf("mode %s PW %d dy %d dx %d: %s (%d mismatches)\n", mode ? "tf32/SW128" : "bf16/SW64", PW, dy, dx, bad ? "FAIL" : "ok", bad);
          bad_total += bad;
        }
  printf(bad_total ? "SOME FAILED\n" : "ALL OK\n");
  return 0;
}
