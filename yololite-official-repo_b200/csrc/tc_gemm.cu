// tcgen05 implicit-GEMM convolution with error-compensated bf16-triple operands (sm_100a).
//
//   C[M, N] = A[M, K] * W[K, N]      M = B*Hout*Wout pixels, K = taps*Cin, N = Cout
//
// Why error compensation: the parity bar is 1e-3 abs on logits after a 69-conv stack; single-pass TF32/BF16/FP16 operands
// miss it by two orders of magnitude, fp32 SIMT FMA is ~5x below the HBM roofline for the
// 96-channel pointwise layers.  Why bf16 triples and not 3xTF32: the tensor core does not round its fp32 accumulator, it
// truncates (the addends are aligned to the largest exponent and cut), which biases every instruction toward zero by ~1.5 ulp
// when the products carry 22 significant bits (TF32 x TF32) -- measured -3.7 / -7.8 / -26 ulp per layer at K = 96 / 256 / 960
// (experiments/exp_accum.cu), and the bias adds up LINEARLY over the layers (1.5e-3 on edge_m's logits).  With x = x1 + x2 + x3
// (bf16 each) the large-magnitude chain A1*W1 has 16-bit products that mostly fit the accumulator exactly and takes half as
// many instructions (K = 16): -0.5 / -1.4 / -6 ulp.  D = A1*W1 (main accumulator) + A1*W2 + A2*W1 + A2*W2 + A1*W3 + A3*W1
// (correction accumulator, 2^-8 smaller; dropped terms 2^-24), five instructions per 16-wide k-step -- the first one,
// A1 x [W1|W2], fills both accumulators at once -- at the same tensor-pipe time as 3xTF32 and 25 % less shared memory.
//
// Structure (one persistent CTA per SM, 14 warps, warp-specialised):
//   warps 0-7  producers: bring the fp32 A tile in (TMA / cp.async, or compute it: depthwise KxK for the fused DWConvBlock,
//              im2col gather for dense 3x3), split it into three bf16 planes in the K-major UMMA layouts
//              (stage = [a1 | a2] as one 128 B SWIZZLE_128B row per pixel and K-slab + a3 as a 64 B SWIZZLE_64B row),
//              fence.proxy.async, arrive on the stage's "full" mbarrier;
//   warp  12   one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (bf16, M=128, K=16 per instruction) into one of
//              two TMEM accumulator pairs, tcgen05.commit frees the smem stage / publishes the accumulators;
//   warps 8-11 epilogue: tcgen05.ld, main + corr, + bias (folded BN) + residual + nearest-upsampled coarser level +
//              ReLU/SiLU, TMA tensor stores / 16 B stores (head layout [B,A,S,S,5+C] written directly).
// The weight image (three bf16 splits, pre-swizzled on the host) stays resident in shared memory for the CTA's lifetime or
// is streamed per K-slab.  All waits are bounded: a protocol bug traps instead of hanging the GPU.
// Resource notes: 448 threads x 128 registers and up to ~220 KB of shared memory -- one CTA per SM, no co-residency with the next
// kernel (PDL only overlaps its prologue on SMs that have already drained).  The single-thread roles are guarded with elect.sync
// (launch.cuh: elect_one), not `lane == 0`.  The shared-memory data pipe (LSU + tensor-core operand reads, 64-85 % busy) is the
// binding resource of the large-map shapes; the epilogue is on their critical path (DESIGN.md section 4 lists what was measured).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include <mutex>

#include "common.cuh"
#include "launch.cuh"

namespace yl {

constexpr int TC_PROD_WARPS = 8;                 // producer warps
constexpr int TC_EPI_WARP0 = 8;                  // epilogue warps 8..11 (warp % 4 == TMEM lane quarter)
constexpr int TC_MMA_WARP = 12;
constexpr int TC_TMA_WARP = 13;                  // MODE 2: one thread issues the halo tile loads (cp.async.bulk.tensor) and W slabs
constexpr int TC_THREADS = 32 * 14;
constexpr int TC_MAX_HALO_SLOTS = 4;
constexpr int TC_BM = 128;
constexpr int TC_P12_BYTES = TC_BM * 128;       // planes a1 | a2 of one K-slab: 128 B per row (32 bf16 + 32 bf16), SWIZZLE_128B
constexpr int TC_P3_BYTES = TC_BM * 64;         // plane a3: 64 B per row, SWIZZLE_64B
constexpr int TC_STAGE_BYTES = TC_P12_BYTES + TC_P3_BYTES;      // 24 KB per A stage
constexpr int TC_SMEM_BUDGET = 226 * 1024;      // of the 227 KB a CTA may use
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_TILE_H = 8, TC_TILE_W = 16;          // spatial tile of the dense KxK TMA path = 128 output pixels
constexpr int TC_EPI_PITCH = 36;                      // floats per staged row: 16 B aligned, conflict-free for 128-bit access
constexpr int TC_STG_BYTES = 5120;                    // per epilogue warp: the padded transpose tile (32 x 36 floats) or the 4 KB
                                                      // SWIZZLE_128B tile a TMA store reads; 1024 B aligned
constexpr int TC_AUX_BYTES = 256 + 512 + 4 * TC_STG_BYTES;      // barriers + tmem slot, bias vector, epilogue staging (one group)


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile unsigned long long* dbg = nullptr, uint32_t tag = 0) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(2000u)      // suspend-time hint (ns): sleep in hardware instead of spinning
        : "memory");
    if (spin > (1u << 22)) {                        // protocol error: fail loudly, never hang the device
      if (dbg) {      // post-mortem word in mapped host memory: tag | parity | warp | block | barrier address
        dbg[tag & 31u] = ((unsigned long long)tag << 56) | ((unsigned long long)(parity & 1u) << 55) | ((unsigned long long)(threadIdx.x >> 5) << 48) |
                 ((unsigned long long)(blockIdx.x & 0xFFFFu) << 32) | (unsigned long long)bar;
        __threadfence_system();
        for (int w = 0; w < 100; ++w) __nanosleep(1000000);      // let the other stuck waiters record their words too
      }
      asm volatile("trap;");
    }
  }
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor: start[0,14) LBO[16,30) SBO[32,46)
  // version[46,48)=1 layout[61,64)=2)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
  // K-major, SWIZZLE_64B (64 B rows), 8-row groups 512 B apart
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split3(float a, float b, uint32_t& p1, uint32_t& p2, uint32_t& p3) { split3_pair<false>(a, b, p1, p2, p3); }

// Write the three bf16 splits of 4 consecutive k (16 B fp32 source chunk `c` = k 4c..4c+3 of the K-slab) of A row `row` into a stage:
//   a1 -> bytes [0,64) of the row's 128 B SWIZZLE_128B line, a2 -> bytes [64,128), a3 -> the row's 64 B SWIZZLE_64B line of the P3 plane.
__device__ __forceinline__ void store_split3(unsigned char* stage, int row, int c, float4 a) {
  uint2 s1, s2, s3;
  split3(a.x, a.y, s1.x, s2.x, s3.x);
  split3(a.z, a.w, s1.y, s2.y, s3.y);
  const uint32_t r7 = (uint32_t)row & 7u, half = ((uint32_t)c & 1u) << 3, c2 = (uint32_t)c >> 1;
  unsigned char* line = stage + (uint32_t)(row >> 3) * 1024u + r7 * 128u;
  *reinterpret_cast<uint2*>(line + ((c2 ^ r7) << 4) + half) = s1;
  *reinterpret_cast<uint2*>(line + (((4u + c2) ^ r7) << 4) + half) = s2;
  *reinterpret_cast<uint2*>(stage + TC_P12_BYTES + (uint32_t)row * 64u + ((c2 ^ (((uint32_t)row >> 1) & 3u)) << 4) + half) = s3;
}

// SiLU / ReLU through one out-of-line helper keeps the (I-cache sensitive) epilogue small.
__device__ __noinline__ float4 act4(float4 o, int act) {
  if (act == YL_ACT_RELU) {
    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
  } else if (act == YL_ACT_SILU) {
    o.x = silu_accurate(o.x); o.y = silu_accurate(o.y); o.z = silu_accurate(o.z); o.w = silu_accurate(o.w);
  }
  return o;
}

#ifdef YL_TIMELINE
// diagnostic build only: CTA 0 stamps %globaltimer (ns) of a few events into the mapped debug words 16 + k (first occurrence only)
#define YL_STAMP(k) do { if (blockIdx.x == 0 && blockIdx.y == 0 && p.dbg && p.dbg[16 + (k)] == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[16 + (k)] = t_; } } while (0)
#define YL_STAMP_LAST(k) do { if (blockIdx.x == 0 && blockIdx.y == 0 && p.dbg) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[16 + (k)] = t_; } } while (0)
#else
#define YL_STAMP(k) do { } while (0)
#define YL_STAMP_LAST(k) do { } while (0)
#endif

template <int MODE, int KS>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap,
                                                                const __grid_constant__ CUtensorMap omap) {
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);   // SWIZZLE_128B atoms
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) YL_STAMP(0);
  const ConvParams& c = p.c;
  const int M = (int)p.M;

  // ---- shared memory carve-up (all slabs 1024 B aligned)
  // weights of one K-slab: three bf16 splits x Nc rows x 64 B (SWIZZLE_64B, K-major): [W1 | W2 | W3], so that [W1|W2] is one
  // B operand of N = 2*Nc.  Resident: nslab slabs; streamed (p.wstream): one slab per A stage.
  const uint32_t w_split_bytes = (uint32_t)p.Nc * 64u, w_slab_bytes = 3u * w_split_bytes;
  const int w_slabs = p.wstream ? p.stages : p.nslab;
  unsigned char* w_base = smem;
  unsigned char* a_ring = w_base + (size_t)w_slabs * w_slab_bytes;               // stages x (a1|a2 16K, a3 8K)
  unsigned char* stage_base = a_ring + (size_t)p.stages * TC_STAGE_BYTES;        // epilogue staging, 1024 B aligned
  unsigned char* halo = stage_base + (size_t)(p.epi2 ? 8 : 4) * p.stg_stride;     // MODE 2: halo_slots x halo_bytes
  float* w2s = reinterpret_cast<float*>(halo + ((MODE == 2 || (MODE == 1 && p.tma_a)) ? (size_t)p.halo_slots * p.halo_bytes : 0));
  // MODE 2: depthwise taps [KS*KS][nslab*32] followed by the depthwise bias row
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(w2s) + (MODE == 2 ? (size_t)(KS * KS + 1) * p.nslab * 32 * 4 : 0));
  uint64_t* full_bar = bars;                       // [stages]   producers -> MMA        (count: producer warps)
  uint64_t* empty_bar = bars + TC_MAX_STAGES;      // [stages]   MMA commit -> producers (count 1)
  uint64_t* tfull_bar = bars + 2 * TC_MAX_STAGES;  // [2]        MMA commit -> epilogue  (count 1)
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]        epilogue -> MMA         (count 4)
  uint64_t* wfull_bar = tempty_bar + 4;            // [stages]   MODE 2 streamed weights: cp.async.bulk complete_tx -> MMA
  uint64_t* hfull_bar = wfull_bar + TC_MAX_STAGES;   // [halo_slots] MODE 2: TMA complete_tx -> producers
  uint64_t* hempty_bar = hfull_bar + TC_MAX_HALO_SLOTS;   // [halo_slots] MODE 2: producers (8 warps) -> TMA thread
  uint64_t* wres_bar = hempty_bar + TC_MAX_HALO_SLOTS;   // resident weight image: cp.async.bulk complete_tx -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 2);   // keeps the staging area behind it 16 B aligned
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);         // bias of this CTA's N chunk (128 floats)
  float* dense_stage = bias_s + 128;                               // dense epilogue only: 4 warps x 32 x Nc floats

  const int chunk_n0 = blockIdx.y * p.Nc;          // first output channel of this CTA's N-chunk
  // streamed weights: rows of this chunk that exist in the image (a partial last chunk leaves stale rows in its slot: they only
  // feed output columns >= Cout, which are never stored)
  const uint32_t w_rows_bytes = (uint32_t)min(p.Nc, p.Npad - chunk_n0) * 64u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      // arrivals per A stage: all producer warps, or one group of 4 in MODE 2 (the two groups take alternate K-slabs)
      mbar_init(smem_u32(&full_bar[s]), MODE == 2 ? 4 : p.prod_warps); mbar_init(smem_u32(&empty_bar[s]), 1); mbar_init(smem_u32(&wfull_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&tfull_bar[a]), 1); mbar_init(smem_u32(&tempty_bar[a]), 4); }
    for (int s = 0; s < TC_MAX_HALO_SLOTS; ++s) { mbar_init(smem_u32(&hfull_bar[s]), 1); mbar_init(smem_u32(&hempty_bar[s]), MODE == 2 ? 4 : TC_PROD_WARPS); }
    mbar_init(smem_u32(wres_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // resident weight image (per slab the three splits' rows [chunk_n0, chunk_n0 + Nc)): bulk copies straight from the
    // pre-swizzled image in L2, asynchronous -- only the MMA thread waits for them, before its first instruction
    if (!p.wstream) {
      const int rows = min(p.Nc, p.Npad - chunk_n0);              // rows of this chunk that exist in the image
      const uint32_t bar = smem_u32(wres_bar);
      if (p.nchunks == 1) {                                         // the whole image is one contiguous block
        const uint32_t bytes = (uint32_t)p.nslab * w_slab_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(w_base)), "l"(p.wimg), "r"(bytes), "r"(bar) : "memory");
      } else {
        const uint32_t bytes = (uint32_t)rows * 64u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3u * (uint32_t)p.nslab * bytes) : "memory");
        for (int ps = 0; ps < 3 * p.nslab; ++ps)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(w_base) + (uint32_t)ps * w_split_bytes), "l"(p.wimg + ((size_t)ps * p.Npad + chunk_n0) * 16), "r"(bytes), "r"(bar)
                       : "memory");
      }
    }
  }
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows of a partial last chunk that lie past the image are zero (generic stores; disjoint from the bulk copies)
  if (!p.wstream && chunk_n0 + p.Nc > p.Npad) {
    const int rows = p.Npad - chunk_n0, zr = p.Nc - rows;          // zr zero rows at the end of every split block
    for (int i = threadIdx.x; i < 3 * p.nslab * zr * 4; i += blockDim.x) {
      const int ps = i / (zr * 4), r = i - ps * (zr * 4);
      reinterpret_cast<float4*>(w_base + (size_t)ps * w_split_bytes + (size_t)rows * 64)[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) bias_s[i] = (c.bias && i < p.Nc && chunk_n0 + i < c.Cout) ? __ldg(c.bias + chunk_n0 + i) : 0.f;
  if (MODE == 2) {
    // depthwise taps + bias row, 16 B at a time (Cin % 4 == 0; w2 / b2 are 16 B aligned blob arrays)
    const int cp4 = p.nslab * 8, k4n = c.Cin >> 2;
    for (int i = threadIdx.x; i < (KS * KS + 1) * cp4; i += TC_THREADS) {
      const int tap = i / cp4, k4 = i - tap * cp4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k4 < k4n) {
        if (tap < KS * KS) v = __ldg(reinterpret_cast<const float4*>(c.w2 + (size_t)tap * c.Cin) + k4);
        else if (c.b2) v = __ldg(reinterpret_cast<const float4*>(c.b2) + k4);
      }
      reinterpret_cast<float4*>(w2s)[i] = v;
    }
  }
  if (warp == TC_TMA_WARP && elect_one()) {          // descriptor fetches off the critical path of the first loads / stores
    if (MODE == 2 || p.tma_a || p.tap_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    if (p.tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&omap) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int tiles = p.num_tiles;
  // PDL trigger AFTER this CTA holds its tensor memory: the next kernel's CTAs may now become resident and run their own
  // prologue while this kernel works (they wait for our completion before reading activations).  Triggering before the TMEM
  // allocation could let a dependent CTA grab columns first and then block this CTA in tcgen05.alloc for ever.
  if (threadIdx.x == 0) YL_STAMP(1);            // prologue done
  pdl_launch_dependents();
  // everything above touched only constant data (weights, biases) and this CTA's shared memory / TMEM: with PDL it overlapped
  // the previous kernel's tail.  From here on activations written by earlier kernels are read.
  pdl_wait();
  if (threadIdx.x == 0) YL_STAMP(2);            // dependency resolved

  if (warp < p.prod_warps) {
    // =============================== producers ===============================
    // 256 threads; per K-slab each thread owns 4 x 16 B of the A tile: rows (t>>3)+32*i, chunk t&7.
    const int t = threadIdx.x;                      // 0..255
    const int ch = t & 7, r0 = t >> 3;
    // Row owned by this thread in piece `ip` of a K-slab for the TMA-fed paths: each 8-lane group of a warp covers one whole row
    // (8 x 16 B); the two rows of a HALF-warp differ in bits 0 and 2 (pairs 0/5, 2/7, 1/4, 3/6 of an 8-row group), so their a1 | a2
    // halves of the SWIZZLE_128B lines and their 64 B lines of the third plane fall into different banks (the plain t >> 3 order put
    // rows r, r + 1 of a half-warp on the same 16 banks: every 8 B split store took two wavefronts).
    const int l3 = lane >> 3, rsel = (l3 & 2) + 5 * (l3 & 1);
    auto piece_row = [&](int ip) { return 8 * (warp + p.prod_warps * (ip >> 1)) + (rsel ^ (ip & 1)); };
    if ((MODE == 0 && p.tma_a) || (MODE == 1 && p.tap_tma)) {
      // (MODE 1 with per-tap TMA boxes: the same, the 128 rows are the tile's pixels)
      // The TMA warp lands each K-slab of the fp32 A tile (128 rows x 128 B, SWIZZLE_128B) in the stage's a1|a2 area; every producer
      // thread reads its own 16 B pieces (4 with 8 producer warps, 8 with 4), and -- a warp covers whole rows, so a __syncwarp
      // separates all reads of a row from the writes -- overwrites the rows IN PLACE with the bf16 splits a1 | a2 (a3 goes to
      // the stage's second plane).
      const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = my_tiles * p.nslab;
      const int npass = TC_BM / (16 * p.prod_warps);             // passes of 4 pieces per thread: 1 (8 producer warps) or 2 (4)
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < total; ++j) {
        mbar_wait(smem_u32(&wfull_bar[stage]), phase, p.dbg, 1u);
        if (threadIdx.x == 0) YL_STAMP(3);      // first A slab landed
        unsigned char* st = a_ring + (size_t)stage * TC_STAGE_BYTES;
        for (int ps = 0; ps < npass; ++ps) {
          float4 a[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = piece_row(4 * ps + i);
            a[i] = *reinterpret_cast<const float4*>(st + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((ch ^ (row & 7)) << 4));
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) store_split3(st, piece_row(4 * ps + i), ch, a[i]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[stage]));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    } else if (MODE == 1 && p.tma_a) {
      // Dense KxK conv with a small Cin (timm blocks.1.0: 3x3 s2, 16 -> 48): the whole-Cin halo of the INPUT for one spatial
      // tile of 8 x 16 output pixels arrives as ONE TMA box; the im2col K-slabs are gathered shared -> shared (a 16 B piece
      // never straddles a tap because Cin % 4 == 0) and split into the three bf16 planes.
      const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int HS = p.halo_slots, HW = p.halo_w;
      const int pix_bytes = c.Cin * 4;
      int hp0[4];                                     // halo pixel of tap (0,0) for this thread's 4 output rows
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = piece_row(i);
        hp0[i] = (row / p.tile_w) * c.stride * HW + (row % p.tile_w) * c.stride;
      }
      int stage = 0, slot = 0;
      uint32_t phase = 0, hphase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(smem_u32(&hfull_bar[slot]), hphase, p.dbg, 2u);            // this tile's halo has landed
        const unsigned char* hb = halo + (size_t)slot * p.halo_bytes;
        for (int s = 0; s < p.nslab; ++s) {
          const int k0 = s * 32 + ch * 4;
          const int tap = k0 / c.Cin, ci = k0 - tap * c.Cin;
          const int ky = tap / c.KS, kx = tap - ky * c.KS;
          const int toff = (ky * HW + kx) * pix_bytes + ci * 4;
          float4 a[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            a[i] = k0 < p.K ? *reinterpret_cast<const float4*>(hb + (size_t)hp0[i] * pix_bytes + toff) : make_float4(0.f, 0.f, 0.f, 0.f);
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.dbg, 3u);
          unsigned char* st = a_ring + (size_t)stage * TC_STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) store_split3(st, piece_row(i), ch, a[i]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&full_bar[stage]));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hempty_bar[slot]));   // this warp no longer reads the halo
        if (++slot == HS) { slot = 0; hphase ^= 1; }
      }
    } else if (MODE != 2) {
      // cp.async pipeline: the 16 B fp32 pieces are copied global -> shared (zero-filled outside the image / past K) into the
      // stage's a1|a2 area (same swizzled 128 B-row layout a TMA box would give), up to `depth` K-slabs ahead of the one
      // being converted, so tens of KB per SM are in flight without holding registers.  Each thread then reads its own pieces
      // back and -- after a __syncwarp, a warp covers whole rows -- overwrites the rows in place with the bf16 splits.
      const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = my_tiles * p.nslab;
      const int depth = min(p.stages - 1, 3);
      uint32_t soff[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = r0 + 32 * i;
        soff[i] = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
      }
      const uint32_t ring = smem_u32(a_ring);
      int issued = 0, i_tile = blockIdx.x, i_s = 0, i_stage = 0;
      uint32_t i_phase = 0;
      const float* rbase[4] = {c.in, c.in, c.in, c.in};
      int roy[4] = {0, 0, 0, 0}, rox[4] = {0, 0, 0, 0};
      bool rok[4] = {false, false, false, false};
      int c_stage = 0;
      for (int j = 0; j < total; ++j) {
        while (issued < total && issued <= j + depth) {
          if (i_s == 0) {                            // new tile: per-row geometry
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = i_tile * TC_BM + r0 + 32 * i;
              rok[i] = m < M;
              const int mm = rok[i] ? m : 0;
              if (MODE == 0) {
                rbase[i] = c.in + (size_t)mm * c.Cin;
              } else {
                const int hw = c.Wout * c.Hout;
                const int b = mm / hw, rem = mm - b * hw;
                const int oy = rem / c.Wout, ox = rem - oy * c.Wout;
                rbase[i] = c.in + (size_t)b * c.Hin * c.Win * c.Cin;
                roy[i] = oy * c.stride - c.pad;
                rox[i] = ox * c.stride - c.pad;
              }
            }
          }
          const int k = i_s * 32 + ch * 4;
          int tap_y = 0, tap_x = 0, ci = k;
          if (MODE == 1) {
            const int tap = k / c.Cin;
            ci = k - tap * c.Cin;
            tap_y = tap / c.KS;
            tap_x = tap - tap_y * c.KS;
          }
          mbar_wait(smem_u32(&empty_bar[i_stage]), i_phase ^ 1, p.dbg, 4u);
          const uint32_t dst = ring + (uint32_t)i_stage * (uint32_t)TC_STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float* src = c.in;
            uint32_t nbytes = 0;
            if (rok[i] && k < p.K) {
              if (MODE == 0) {
                src = rbase[i] + k; nbytes = 16;
              } else {
                const int iy = roy[i] + tap_y, ix = rox[i] + tap_x;
                if (iy >= 0 && iy < c.Hin && ix >= 0 && ix < c.Win) {
                  src = rbase[i] + ((size_t)iy * c.Win + ix) * c.Cin + ci; nbytes = 16;
                }
              }
            }
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + soff[i]), "l"(src), "r"(nbytes) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          ++issued;
          if (++i_s == p.nslab) { i_s = 0; i_tile += gridDim.x; }
          if (++i_stage == p.stages) { i_stage = 0; i_phase ^= 1; }
        }
        switch (issued - j - 1) {                    // groups allowed to stay in flight
          case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
          case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
          case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
          default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        }
        unsigned char* st = a_ring + (size_t)c_stage * TC_STAGE_BYTES;
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(st + soff[i]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) store_split3(st, r0 + 32 * i, ch, a[i]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[c_stage]));
        if (++c_stage == p.stages) c_stage = 0;
      }
    } else {
      // fused depthwise -> pointwise on spatial tiles of tile_h x tile_w (<= 128) output pixels.  Per K-slab the
      // (tile_h + KS - 1) x (tile_w + KS - 1) x 32-channel halo of the INPUT arrives in a ring slot as ONE TMA tile
      // (cp.async.bulk.tensor.4d over the NHWC tensor, issued by the TMA warp; out-of-image pixels and channels past K
      // are zero-filled by the hardware = the conv padding).  Every producer thread computes the depthwise KS x KS
      // (+ bias, act2) for 4 horizontally adjacent pixels x 4 channels from shared memory, splits it into bf16 triples and
      // writes the A stage.  With p.wstream the K-slab of the pointwise weights travels with the A stage (cp.async.bulk from L2).
      // The 8 producer warps form two groups of 4 that take alternate K-slabs (both A stages are then in progress at
      // once); inside a group every thread owns a patch of 2 rows x 4 pixels x 4 channels, so each halo row it loads
      // feeds two output rows -- the stage is bound by shared-memory reads, and this cuts them by ~40 %.
      const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = my_tiles * p.nslab;
      const int HS = p.halo_slots;
      const int TW = p.tile_w, HW = p.halo_w;
      const int grp = t >> 7, g = (t & 127) >> 3;               // group, patch index inside the group (16 patches x 8 chunks)
      const int GPR = TW >> 2;                                   // patches per tile row
      const int rp = g / GPR, cg = g - rp * GPR;
      const bool valid = 2 * rp < p.tile_h;                      // tile_h is even; patches past the tile do nothing
      const int ty = 2 * rp, tx0 = 4 * cg;
      const int cp = p.nslab * 32;
      const uint32_t hoff = ((uint32_t)(ty * HW + tx0) * 8u + (uint32_t)ch) * 16u;
      const int arow0 = ty * TW + tx0;
      // stride-2 mapping: patch g = one row of 4 output pixels
      const bool valid2 = rp < p.tile_h;
      const uint32_t hoff2 = ((uint32_t)(2 * rp * HW + 2 * tx0) * 8u + (uint32_t)ch) * 16u;
      const int arow2 = rp * TW + tx0;
      // j % HS, j / HS and j % nslab advance by 2 per iteration: kept as running counters (HS is 2 or 4; a division by a runtime
      // value is ~60 dependent instructions at the head of every item)
      int c_slot = grp & (HS - 1), c_s = grp % p.nslab;
      uint32_t hround = (uint32_t)(grp / HS);
      for (int j = grp; j < total; j += 2) {
        const int stage = j & 1;                                           // MODE 2 runs with 2 A stages
        const uint32_t hphase = hround & 1u, phase = (uint32_t)(j >> 1) & 1u;
        mbar_wait(smem_u32(&hfull_bar[c_slot]), hphase, p.dbg, 5u);          // this item's halo tile has landed
        float4 a[8];                                              // [row 0: 4 pixels][row 1: 4 pixels]
        if (p.dw_stride == 2) {
          // stride 2: the tile is at most 64 output pixels (rows 64..127 of the MMA tile are unused); a thread owns ONE row of
          // 4 output pixels = KS rows x (KS + 6) columns of the halo
          if (valid2) {
            const unsigned char* hb = halo + (size_t)c_slot * p.halo_bytes + hoff2;
            const float* wk = w2s + c_s * 32 + ch * 4;
            // the multiply-adds run as packed pairs (FFMA2: channels (0,1) and (2,3) of every 16-byte piece)
            f32x2 acc[4][2];
            {
              const float4 b4 = *reinterpret_cast<const float4*>(wk + KS * KS * cp);
#pragma unroll
              for (int i = 0; i < 4; ++i) { acc[i][0] = pack2(b4.x, b4.y); acc[i][1] = pack2(b4.z, b4.w); }
            }
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
              f32x2 h[KS + 6][2];
#pragma unroll
              for (int x = 0; x < KS + 6; ++x) {
                const float4 h4 = *reinterpret_cast<const float4*>(hb + (size_t)(ky * HW + x) * 128);
                h[x][0] = pack2(h4.x, h4.y); h[x][1] = pack2(h4.z, h4.w);
              }
#pragma unroll
              for (int kx = 0; kx < KS; ++kx) {
                const float4 w4 = *reinterpret_cast<const float4*>(wk + (ky * KS + kx) * cp);
                const f32x2 w01 = pack2(w4.x, w4.y), w23 = pack2(w4.z, w4.w);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  acc[i][0] = fma2(h[2 * i + kx][0], w01, acc[i][0]);
                  acc[i][1] = fma2(h[2 * i + kx][1], w23, acc[i][1]);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { unpack2(acc[i][0], a[i].x, a[i].y); unpack2(acc[i][1], a[i].z, a[i].w); }
          }
        } else if (valid) {
          const unsigned char* hb = halo + (size_t)c_slot * p.halo_bytes + hoff;
          const float* wk = w2s + c_s * 32 + ch * 4;
          // the multiply-adds run as packed pairs (FFMA2: channels (0,1) and (2,3) of every 16-byte piece)
          f32x2 acc[8][2];
          {
            const float4 b4 = *reinterpret_cast<const float4*>(wk + KS * KS * cp);
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i][0] = pack2(b4.x, b4.y); acc[i][1] = pack2(b4.z, b4.w); }
          }
          f32x2 wprev[KS > 0 ? KS : 1][2];
#pragma unroll
          for (int hy = 0; hy <= KS; ++hy) {
            f32x2 h[KS + 3][2];
#pragma unroll
            for (int x = 0; x < KS + 3; ++x) {
              const float4 h4 = *reinterpret_cast<const float4*>(hb + (size_t)(hy * HW + x) * 128);
              h[x][0] = pack2(h4.x, h4.y); h[x][1] = pack2(h4.z, h4.w);
            }
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
              if (hy >= 1) {                                      // output row 1 sees this halo row as tap row hy - 1
                const f32x2 w01 = wprev[kx][0], w23 = wprev[kx][1];      // tap row hy - 1: kept from the previous halo row
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  acc[4 + i][0] = fma2(h[i + kx][0], w01, acc[4 + i][0]);
                  acc[4 + i][1] = fma2(h[i + kx][1], w23, acc[4 + i][1]);
                }
              }
              if (hy < KS) {                                      // output row 0: tap row hy
                const float4 w4 = *reinterpret_cast<const float4*>(wk + (hy * KS + kx) * cp);
                const f32x2 w01 = pack2(w4.x, w4.y), w23 = pack2(w4.z, w4.w);
                wprev[kx][0] = w01; wprev[kx][1] = w23;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  acc[i][0] = fma2(h[i + kx][0], w01, acc[i][0]);
                  acc[i][1] = fma2(h[i + kx][1], w23, acc[i][1]);
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { unpack2(acc[i][0], a[i].x, a[i].y); unpack2(acc[i][1], a[i].z, a[i].w); }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hempty_bar[c_slot]));   // this warp no longer reads the slot
        if (c.act2 == YL_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i].x = fmaxf(a[i].x, 0.f); a[i].y = fmaxf(a[i].y, 0.f); a[i].z = fmaxf(a[i].z, 0.f); a[i].w = fmaxf(a[i].w, 0.f); }
        } else if (c.act2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = act4(a[i], c.act2);
        }
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.dbg, 6u);
        unsigned char* st = a_ring + (size_t)stage * TC_STAGE_BYTES;
        const int nst = p.dw_stride == 2 ? (valid2 ? 4 : 0) : (valid ? 8 : 0);
        {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i >= nst) break;
            const int row = p.dw_stride == 2 ? arow2 + i : arow0 + (i >> 2) * TW + (i & 3);
            store_split3(st, row, ch, a[i]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[stage]));
        c_slot += 2; if (c_slot >= HS) { c_slot -= HS; ++hround; }
        c_s += 2; while (c_s >= p.nslab) c_s -= p.nslab;
      }
    }
  } else if (MODE == 0 && warp == TC_TMA_WARP) {
    // =============================== TMA issuer (MODE 0) ===============================
    if (p.tma_a && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int s = 0; s < p.nslab; ++s) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.dbg, 7u);          // the MMAs that read this stage are done
          const uint32_t bar = smem_u32(&wfull_bar[stage]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                       "r"((uint32_t)TC_P12_BYTES + (p.wstream ? 3u * w_rows_bytes : 0u)) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(smem_u32(a_ring) + (uint32_t)stage * (uint32_t)TC_STAGE_BYTES), "l"(&tmap), "r"(s * 32), "r"(tile * TC_BM), "r"(bar)
                       : "memory");
          if (p.wstream) {          // K too long for a resident weight image: this K-slab of W (three splits) rides with the A tile
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(smem_u32(w_base) + (uint32_t)stage * w_slab_bytes + (uint32_t)q3 * w_split_bytes),
                             "l"(p.wimg + (((size_t)s * 3 + q3) * p.Npad + chunk_n0) * 16), "r"(w_rows_bytes), "r"(bar) : "memory");
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (MODE == 1 && warp == TC_TMA_WARP) {
    // =============================== TMA issuer (MODE 1: whole-Cin halo per tile) ===============================
    const bool leader = elect_one();        // ONE election for the whole role: a second elect.sync further down the else-if chain
                                            // would wait forever for the lane that took an earlier branch
    if (p.tma_a && leader) {
      const int per_img = p.tiles_x * p.tiles_y;
      int slot = 0;
      uint32_t hphase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int b = tile / per_img, rem = tile - b * per_img;
        const int y0 = (rem / p.tiles_x) * p.tile_h * c.stride - c.pad, x0 = (rem % p.tiles_x) * p.tile_w * c.stride - c.pad;
        mbar_wait(smem_u32(&hempty_bar[slot]), hphase ^ 1, p.dbg, 8u);
        const uint32_t bar = smem_u32(&hfull_bar[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)p.halo_tx) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(halo) + (uint32_t)slot * (uint32_t)p.halo_bytes), "l"(&tmap), "r"(0), "r"(x0), "r"(y0), "r"(b), "r"(bar)
            : "memory");
        if (++slot == p.halo_slots) { slot = 0; hphase ^= 1; }
      }
    } else if (p.tap_tma && leader) {
      // dense k x k, stride 1: K-slab s = tap * spt + cs is the box of 32 channels [32 cs, 32 cs + 32) x tile_w x tile_h pixels of the
      // NHWC input shifted by the tap (ky - pad, kx - pad); pixels outside the image and channels past Cin arrive as zeros.  It lands
      // as 128 rows of 128 B (SWIZZLE_128B) = the layout the producers convert in place.  Streamed weights ride on the same barrier.
      const int per_img = p.tiles_x * p.tiles_y;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int b = tile / per_img, rem = tile - b * per_img;
        const int y0 = (rem / p.tiles_x) * p.tile_h - c.pad, x0 = (rem % p.tiles_x) * p.tile_w - c.pad;
        int tap = 0, cs = 0;
        for (int s = 0; s < p.nslab; ++s) {
          const int ky = tap / c.KS, kx = tap - ky * c.KS;
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.dbg, 18u);          // the MMAs that read this stage are done
          const uint32_t bar = smem_u32(&wfull_bar[stage]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                       "r"((uint32_t)TC_P12_BYTES + (p.wstream ? 3u * w_rows_bytes : 0u)) : "memory");
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(smem_u32(a_ring) + (uint32_t)stage * (uint32_t)TC_STAGE_BYTES), "l"(&tmap), "r"(cs * 32), "r"(x0 + kx), "r"(y0 + ky), "r"(b), "r"(bar)
              : "memory");
          if (p.wstream) {
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(smem_u32(w_base) + (uint32_t)stage * w_slab_bytes + (uint32_t)q3 * w_split_bytes),
                             "l"(p.wimg + (((size_t)s * 3 + q3) * p.Npad + chunk_n0) * 16), "r"(w_rows_bytes), "r"(bar) : "memory");
          }
          if (++cs == p.spt) { cs = 0; ++tap; }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (p.wstream && leader) {
      // generic gather (cp.async producers) with a K too long for resident weights (the dense 3x3 convs of the YOLOLiteMS FPN,
      // K = 9 * 196 ... 9 * 328): this thread streams the K-slab of W (three splits) that belongs to each A stage from L2
      const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it)
        for (int s = 0; s < p.nslab; ++s) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.dbg, 9u);          // the MMAs that read this stage's W slot are done
          const uint32_t bar = smem_u32(&wfull_bar[stage]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3u * w_rows_bytes) : "memory");
#pragma unroll
          for (int q3 = 0; q3 < 3; ++q3)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(w_base) + (uint32_t)stage * w_slab_bytes + (uint32_t)q3 * w_split_bytes),
                           "l"(p.wimg + (((size_t)s * 3 + q3) * p.Npad + chunk_n0) * 16), "r"(w_rows_bytes), "r"(bar) : "memory");
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    }
    __syncwarp();
  } else if (MODE == 2 && warp == TC_TMA_WARP) {
    // =============================== TMA issuer (MODE 2) ===============================
    if (elect_one()) {
      const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = my_tiles * p.nslab;
      const int per_img = p.tiles_x * p.tiles_y;
      constexpr int PAD = KS / 2;
      int i_tile = blockIdx.x, i_s = 0, slot = 0, stage = 0;
      int b = 0, y0 = 0, x0 = 0;
      uint32_t hphase = 0, phase = 0;
      for (int j = 0; j < total; ++j) {
        if (i_s == 0) {
          b = i_tile / per_img;
          const int rem = i_tile - b * per_img;
          y0 = (rem / p.tiles_x) * p.tile_h * p.dw_stride - PAD;
          x0 = (rem % p.tiles_x) * p.tile_w * p.dw_stride - PAD;
        }
        mbar_wait(smem_u32(&hempty_bar[slot]), hphase ^ 1, p.dbg, 10u);          // all producer warps are done with item j - HS
        {
          const uint32_t bar = smem_u32(&hfull_bar[slot]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)p.halo_bytes) : "memory");
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(smem_u32(halo) + (uint32_t)slot * (uint32_t)p.halo_bytes), "l"(&tmap), "r"(i_s * 32), "r"(x0), "r"(y0), "r"(b), "r"(bar)
              : "memory");
        }
        if (p.wstream) {
          // the MMAs that read this stage's A and W slots (item j - stages) are done: refill the W slot from L2
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.dbg, 11u);
          const uint32_t bar = smem_u32(&wfull_bar[stage]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(w_slab_bytes) : "memory");
#pragma unroll
          for (int q3 = 0; q3 < 3; ++q3) {
            const float* src = p.wimg + (((size_t)i_s * 3 + q3) * p.Npad + chunk_n0) * 16;
            const uint32_t dst = smem_u32(w_base) + (uint32_t)stage * w_slab_bytes + (uint32_t)q3 * w_split_bytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(w_split_bytes), "r"(bar) : "memory");
          }
        }
        if (++i_s == p.nslab) { i_s = 0; i_tile += gridDim.x; }
        if (++slot == p.halo_slots) { slot = 0; hphase ^= 1; }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == TC_MMA_WARP) {
    // =============================== MMA issuer ===============================
    // Two accumulators per tile: `main` takes A1*W1 (16-bit products: most of them add to the accumulator without loss), `corr`
    // the five 2^-8 ... 2^-16 smaller cross terms.  The tensor core truncates its fp32 accumulator, so the chain that carries the
    // magnitude must be short and made of short products; the epilogue adds the two in fp32 (round to nearest).
    if (elect_one()) {
      const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Nc >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t idesc_2n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * p.Nc) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t buf_stride = (uint32_t)(2 * p.Nc);
      if (!p.wstream) mbar_wait(smem_u32(wres_bar), 0, p.dbg, 12u);          // the resident weight image has landed
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.dbg, 13u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_main = tmem_base + (uint32_t)acc * buf_stride;
        const uint32_t d_corr = d_main + (uint32_t)p.Nc;
        uint32_t first = 0;
        for (int s = 0; s < p.nslab; ++s) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, p.dbg, 14u);
          YL_STAMP(4);                          // MMA issuer: first operand stage ready
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a12 = smem_u32(a_ring + (size_t)stage * TC_STAGE_BYTES);      // rows: [a1 (64 B) | a2 (64 B)], SWIZZLE_128B
          const uint32_t a3 = a12 + (uint32_t)TC_P12_BYTES;                            // rows: a3 (64 B), SWIZZLE_64B
          uint32_t b1 = smem_u32(w_base + (size_t)s * w_slab_bytes);
          if (p.wstream) {
            // this stage's W slab: MODES 1 / 2 wait for it here; in MODE 0 it shares the A tile's barrier, which the
            // producers have already waited on
            if (MODE == 2 || (MODE == 1 && !p.tap_tma)) mbar_wait(smem_u32(&wfull_bar[stage]), phase, p.dbg, 15u);
            b1 = smem_u32(w_base) + (uint32_t)stage * w_slab_bytes;
          }
          const uint32_t b2 = b1 + w_split_bytes, b3 = b2 + w_split_bytes;
          const int ksteps = min(2, (p.K - s * 32 + 15) >> 4);
          for (int j = 0; j < ksteps; ++j) {
            const uint32_t ko = (uint32_t)j * 32u;          // 16 bf16 = 32 B along K inside the swizzled row
            mma_bf16(d_main, make_desc(a12 + ko), make_desc64(b1 + ko), idesc_2n, first);          // A1 x [W1|W2] -> [main|corr]
            mma_bf16(d_corr, make_desc(a12 + 64u + ko), make_desc64(b1 + ko), idesc_n, 1);         // A2 x W1
            mma_bf16(d_corr, make_desc(a12 + 64u + ko), make_desc64(b2 + ko), idesc_n, 1);         // A2 x W2
            mma_bf16(d_corr, make_desc(a12 + ko), make_desc64(b3 + ko), idesc_n, 1);               // A1 x W3
            mma_bf16(d_corr, make_desc64(a3 + ko), make_desc64(b1 + ko), idesc_n, 1);              // A3 x W1
            first = 1;
          }
          mma_commit(smem_u32(&empty_bar[stage]));           // frees this A stage once the MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        mma_commit(smem_u32(&tfull_bar[acc]));               // accumulators complete
        YL_STAMP(5);                            // first tile's MMAs issued
        YL_STAMP_LAST(8);                       // last tile's MMAs issued
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < TC_MMA_WARP) {
    // =============================== epilogue ===============================
    // TMEM -> registers (lane = row) -> padded smem transpose -> coalesced 16 B global accesses (lane = 4 columns,
    // 8 lanes per 128 B row segment): bias, residual and the upsampled coarser level are read with the same mapping.
    // With p.epi2 (MODE 0 + TMA) there are two groups of four warps: group 0 = warps 8-11 drains TMEM accumulator 0 (this
    // CTA's even tiles), group 1 = warps 4-7 drains accumulator 1 (odd tiles), so two tiles' epilogues overlap.
    const int q = warp & 3;                                   // TMEM lane quarter == warp index % 4
    const int grp = warp >= TC_EPI_WARP0 ? 0 : 1;
    const int N = c.Cout;
    const int D = c.anchors > 0 ? N / c.anchors : N;
    unsigned char* sb = stage_base + (size_t)(grp * 4 + q) * p.stg_stride;
    float* stg = reinterpret_cast<float*>(sb);
    const bool vec = (N & 3) == 0 && c.anchors <= 1;
    // dense staging: N not a multiple of 4 (head outputs, 5+C channels), single chunk, plain [M][N] output
    const bool dense = p.dense_epi != 0;
    float* dstg = dense_stage + q * (32 * p.Nc);
    bool dense_pending = false;                              // a bulk store of this warp's dense staging block may still be reading it
    const int vr = lane >> 3, vc = (lane & 7) * 4;            // vector path: rows vr + 4*it, columns vc..vc+3
    const int hw = c.Wout * c.Hout;
    int acc = p.epi2 ? grp : 0;
    uint32_t acc_phase = 0;
    const int tile_step = p.epi2 ? 2 * (int)gridDim.x : (int)gridDim.x;
    for (int tile = blockIdx.x + (p.epi2 ? grp * (int)gridDim.x : 0); tile < tiles; tile += tile_step) {
      const int mw = tile * TC_BM + q * 32;                   // first row of this warp (linear modes)
      const bool spatial = MODE == 2 || (MODE == 1 && (p.tma_a || p.tap_tma));   // tile = tile_h x tile_w output pixels (else 128 consecutive rows)
      const int rows_ok = spatial ? 32 : min(32, M - mw);      // rows of this warp inside the matrix
      if (p.tma_out) {
        // TMEM -> registers (lane = row) -> +bias, act -> SWIZZLE_128B staging tile (conflict-free 16 B stores) -> ONE TMA tensor
        // store per warp and 32-column block; rows / columns outside the tensor are clipped by the hardware
        int c1 = mw, c2 = 0, c3 = 0;
        bool wvalid = mw < M;
        if (spatial) {
          const int per_img = p.tiles_x * p.tiles_y;
          const int b = tile / per_img, rem = tile - b * per_img;
          const int rpw = 32 / p.tile_w;                       // tile rows covered by one warp (tile_w divides 32)
          c1 = (rem % p.tiles_x) * p.tile_w; c2 = (rem / p.tiles_x) * p.tile_h + q * rpw; c3 = b;
          wvalid = q * rpw < p.tile_h && c2 < c.Hout;
        }
        // nearest-upsampled coarser level (FPN laterals, linear tiles): this lane's row reads 32 consecutive channels of its source
        // pixel; the loads are issued before the accumulator is read so that their latency overlaps it
        const float* up_row = nullptr;
        if (MODE == 0 && c.up) {
          const int m = min(mw + lane, M - 1);
          const int b = m / hw, rem = m - b * hw;
          const int oy = rem / c.Wout, ox = rem - oy * c.Wout;
          up_row = c.up + ((size_t)(b * c.Hu + nearest_src(oy, c.Hu, c.Hout)) * c.Wu + nearest_src(ox, c.Wu, c.Wout)) * N + chunk_n0;
        }
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.dbg, 16u);
        if (q == 0 && lane == 0) YL_STAMP(6);   // first accumulator ready
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * (uint32_t)(2 * p.Nc);
        for (int col = 0; col < p.Nc; col += 32) {
          float4 o[8];
          if (MODE == 0 && c.up) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = chunk_n0 + col + 4 * j < N ? __ldg(reinterpret_cast<const float4*>(up_row + col + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          {
            uint32_t v[32], w[32];
            if (p.Nc - col > 16) {
              tmem_ld32_nowait(taddr + (uint32_t)col, v);
              tmem_ld32_nowait(taddr + (uint32_t)(p.Nc + col), w);
              tmem_ld_wait();
            } else {
              float a[16], b[16];
              tmem_ld16(taddr + (uint32_t)col, a);
              tmem_ld16(taddr + (uint32_t)(p.Nc + col), b);
#pragma unroll
              for (int j = 0; j < 16; ++j) { v[j] = __float_as_uint(a[j]); w[j] = __float_as_uint(b[j]); v[16 + j] = 0; w[16 + j] = 0; }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bia = col + 4 * j < 128 ? *reinterpret_cast<const float4*>(bias_s + col + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
              // o += (main + corr) + bias, as packed pairs (FADD2)
              const f32x2 s01 = add2(add2(pack2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                                          pack2(__uint_as_float(w[4 * j]), __uint_as_float(w[4 * j + 1]))), pack2(bia.x, bia.y));
              const f32x2 s23 = add2(add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])),
                                          pack2(__uint_as_float(w[4 * j + 2]), __uint_as_float(w[4 * j + 3]))), pack2(bia.z, bia.w));
              if (MODE == 0 && c.up) {          // + the upsampled coarser level already in o
                unpack2(add2(pack2(o[j].x, o[j].y), s01), o[j].x, o[j].y);
                unpack2(add2(pack2(o[j].z, o[j].w), s23), o[j].z, o[j].w);
              } else {
                unpack2(s01, o[j].x, o[j].y);
                unpack2(s23, o[j].z, o[j].w);
              }
            }
          }
          if (c.act == YL_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x = fmaxf(o[j].x, 0.f); o[j].y = fmaxf(o[j].y, 0.f); o[j].z = fmaxf(o[j].z, 0.f); o[j].w = fmaxf(o[j].w, 0.f); }
          } else if (c.act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = act4(o[j], c.act);
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the staging tile
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(sb + lane * 128 + ((j ^ (lane & 7)) << 4)) = o[j];
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && wvalid) {
            if (spatial)
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                           ::"l"(&omap), "r"(chunk_n0 + col), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(sb)) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                           ::"l"(&omap), "r"(chunk_n0 + col), "r"(c1), "r"(smem_u32(sb)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        if (p.epi2) acc_phase ^= 1;
        else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      // element offset of the output row for each of the 8 rows this lane owns, -1 = outside
      int orow[8];
      if (spatial) {
        const int per_img = p.tiles_x * p.tiles_y;
        const int b = tile / per_img, rem = tile - b * per_img;
        const int TW = p.tile_w, TH = p.tile_h;
        const int y0 = (rem / p.tiles_x) * TH, x0 = (rem % p.tiles_x) * TW;
        int ry = (q * 32 + vr) / TW, rx = (q * 32 + vr) - ry * TW;      // rows vr + 4*it: step 4 pixels (TW % 4 == 0)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int y = y0 + ry, x = x0 + rx;
          orow[it] = (ry < TH && y < c.Hout && x < c.Wout) ? ((b * c.Hout + y) * c.Wout + x) * N : -1;
          rx += 4;
          if (rx >= TW) { rx -= TW; ++ry; }
        }
      } else {
#pragma unroll
        for (int it = 0; it < 8; ++it) orow[it] = (vr + 4 * it < rows_ok) ? (mw + vr + 4 * it) * N : -1;
      }
      // element offsets of the nearest-upsample source row for the 8 rows this lane owns (one division per tile)
      int up_off[8];
      if (c.up && (vec || dense)) {
        int m = min(mw + vr, M - 1);
        int b = m / hw, rem = m - b * hw;
        int oy = rem / c.Wout, ox = rem - oy * c.Wout;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          up_off[it] = ((b * c.Hu + nearest_src(oy, c.Hu, c.Hout)) * c.Wu + nearest_src(ox, c.Wu, c.Wout)) * N;
          ox += 4;
          while (ox >= c.Wout) { ox -= c.Wout; if (++oy == c.Hout) { oy = 0; if (b + 1 < c.B) ++b; } }
        }
      }
      if (dense_pending) {                                     // the previous tile's bulk store has read the staging block
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        dense_pending = false;
      }
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.dbg, 17u);
      if (q == 0 && lane == 0) YL_STAMP(6);     // first accumulator ready
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * (uint32_t)(2 * p.Nc);
      for (int col = 0; col < p.Nc; col += 32) {
        const int wcols = min(32, p.Nc - col);                // 32, or 16 for the last block
        {
          uint32_t v[32], w[32];                              // main and correction accumulators of this row
          if (wcols > 16) {
            tmem_ld32_nowait(taddr + (uint32_t)col, v);
            tmem_ld32_nowait(taddr + (uint32_t)(p.Nc + col), w);
            tmem_ld_wait();
          } else {
            float a[16], b[16];
            tmem_ld16(taddr + (uint32_t)col, a);
            tmem_ld16(taddr + (uint32_t)(p.Nc + col), b);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = __float_as_uint(a[j]); w[j] = __float_as_uint(b[j]); v[16 + j] = 0; w[16 + j] = 0; }
          }
          float4* srow = reinterpret_cast<float4*>(stg + lane * TC_EPI_PITCH);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            {
              float4 sv;
              unpack2(add2(pack2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                           pack2(__uint_as_float(w[4 * j]), __uint_as_float(w[4 * j + 1]))), sv.x, sv.y);
              unpack2(add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])),
                           pack2(__uint_as_float(w[4 * j + 2]), __uint_as_float(w[4 * j + 3]))), sv.z, sv.w);
              srow[j] = sv;
            }
        }
        __syncwarp();
        const int nb = chunk_n0 + col;                        // first global column of the block
        if (vec || dense) {
          // lane = 4 consecutive columns of rows vr, vr+4, ..., vr+28
          const int n = nb + vc;
          if (vc < wcols && n < N) {
            const float4 bia = c.bias ? __ldg(reinterpret_cast<const float4*>(c.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float* sp = stg + vr * TC_EPI_PITCH + vc;
            float4 o[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              o[it] = *reinterpret_cast<const float4*>(sp + it * 4 * TC_EPI_PITCH);
              o[it].x += bia.x; o[it].y += bia.y; o[it].z += bia.z; o[it].w += bia.w;
            }
            if (c.res) {          // all 8 loads are issued before the first use (rows past M are clamped, never stored)
              float4 t[8];
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                t[it] = __ldg(reinterpret_cast<const float4*>(c.res + max(orow[it], 0) + n));
              }
#pragma unroll
              for (int it = 0; it < 8; ++it) { o[it].x += t[it].x; o[it].y += t[it].y; o[it].z += t[it].z; o[it].w += t[it].w; }
            }
            if (c.up) {
              float4 t[8];
#pragma unroll
              for (int it = 0; it < 8; ++it) t[it] = __ldg(reinterpret_cast<const float4*>(c.up + up_off[it] + n));
#pragma unroll
              for (int it = 0; it < 8; ++it) { o[it].x += t[it].x; o[it].y += t[it].y; o[it].z += t[it].z; o[it].w += t[it].w; }
            }
            if (c.act == YL_ACT_RELU) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                o[it].x = fmaxf(o[it].x, 0.f); o[it].y = fmaxf(o[it].y, 0.f); o[it].z = fmaxf(o[it].z, 0.f); o[it].w = fmaxf(o[it].w, 0.f);
              }
            } else if (c.act) {
#pragma unroll
              for (int it = 0; it < 8; ++it) o[it] = act4(o[it], c.act);
            }
            if (!dense) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (orow[it] >= 0) *reinterpret_cast<float4*>(c.out + orow[it] + n) = o[it];
            } else {
              float* dp = dstg + vr * N + (n - chunk_n0);
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                float* d = dp + it * 4 * N;
                d[0] = o[it].x;
                if (n + 1 < N) d[1] = o[it].y;
                if (n + 2 < N) d[2] = o[it].z;
                if (n + 3 < N) d[3] = o[it].w;
              }
            }
          }
        } else {
          const int n = nb + lane;                            // generic path (anchors > 1 ...): lane = column
          if (lane < wcols && n < N) {
            const float bia = c.bias ? __ldg(c.bias + n) : 0.f;
            const int a = n / D, dd = n - a * D;
            for (int r = 0; r < rows_ok; ++r) {
              const int m = mw + r;
              float o = stg[r * TC_EPI_PITCH + lane] + bia;
              if (c.res) o += __ldg(c.res + (size_t)m * N + n);
              size_t oidx = (size_t)m * N + n;
              if (c.up || c.anchors > 1) {
                const int b = m / hw, rem = m - b * hw;
                const int oy = rem / c.Wout, ox = rem - oy * c.Wout;
                if (c.up)
                  o += __ldg(c.up + (((size_t)b * c.Hu + nearest_src(oy, c.Hu, c.Hout)) * c.Wu + nearest_src(ox, c.Wu, c.Wout)) * N + n);
                if (c.anchors > 1) oidx = ((((size_t)b * c.anchors + a) * c.Hout + oy) * c.Wout + ox) * D + dd;
              }
              c.out[oidx] = act_fn(o, c.act);
            }
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));     // TMEM is drained: the MMA warp may reuse it
      if (dense) {
        // rows of a warp are consecutive in memory ([M][N] row-major, one chunk): the staging block IS the memory image of the span
        float* dst = c.out + (size_t)mw * N;
        const int tot = rows_ok * N, tot4 = tot >> 2;
        if (tot > 0 && (tot & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)smem_u32(dstg)) & 15) == 0) {
          // ... written by ONE bulk copy (32 rows x N x 4 B is a multiple of 16 for every N); the next tile waits for its read below
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(dstg)), "r"((uint32_t)tot * 4u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          dense_pending = true;
        } else {
          for (int i = lane; i < tot4; i += 32)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(dstg)[i];
          for (int i = (tot4 << 2) + lane; i < tot; i += 32) dst[i] = dstg[i];
          __syncwarp();
        }
      }
      if (p.epi2) acc_phase ^= 1;                              // this group owns one accumulator
      else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (q == 0 && lane == 0) YL_STAMP_LAST(9);    // epilogue: last tile handed to the store engine
    if ((p.tma_out || p.dense_epi) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores are complete
    if (q == 0 && lane == 0) YL_STAMP_LAST(10);   // stores complete
  }

  // ---- teardown
  if (threadIdx.x == 0) YL_STAMP_LAST(11);     // reached the teardown barrier
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
// Decide whether (K, N) fits the resident-weight design and how N is chunked.  Shared with the packer through
// the image layout only ([2][nslab][Npad][32]), which does not depend on the chunking.
static bool tc_dense_epi(int N, int anchors, int Nc, int nchunks) { return (N & 3) != 0 && anchors <= 1 && nchunks == 1 && Nc <= 128; }

struct TcPlan {
  int Nc = 0, nchunks = 0, stages = 0, halo_slots = 0, wstream = 0, epi2 = 0;
  int tile_w = TC_TILE_W, tile_h = TC_TILE_H, halo_w = 0, halo_h = 0, halo_pix = 0, halo_bytes = 0;
  size_t smem = 0;
};

// MODE 2 spatial tile: tile_w (multiple of 4) x tile_h <= 128 pixels minimising halo pixels loaded + GEMM rows issued
static void tc_pick_tile(int ks, int stride, int Hout, int Wout, TcPlan* pl) {
  long long best = -1;
  const int maxpix = stride == 2 ? 64 : 128;                   // stride 2: one 4-pixel row per producer thread of a group
  for (int tw = 4; tw <= maxpix; tw += 4) {
    const int th = stride == 2 ? maxpix / tw : (maxpix / tw) & ~1;   // stride 1: even, a producer thread owns a 2-row patch
    if (th < 1 || (stride == 1 && th < 2)) continue;
    const int hw_ = (tw - 1) * stride + ks, hh_ = (th - 1) * stride + ks;
    const int hp = hw_ * hh_;
    if (hp * 128 > 48 * 1024 || hw_ > 256 || hh_ > 256) continue;   // keep a ring of >= 2 slots affordable
    const long long ntile = (long long)((Wout + tw - 1) / tw) * ((Hout + th - 1) / th);
    long long cost = ntile * (hp + 128);
    if (32 % tw != 0) cost += cost * 15 / 100;                  // tile widths dividing 32 can use the TMA-store epilogue
    if (best < 0 || cost < best) { best = cost; pl->tile_w = tw; pl->tile_h = th; pl->halo_w = hw_; pl->halo_h = hh_; pl->halo_pix = hp; }
  }
  pl->halo_bytes = pl->halo_pix * 128;
}

// MODE 2 with weights too large to stay resident next to the halo ring: one K-slab (three splits) of W per A stage, from L2
static bool tc_plan_stream(int nslab, int Npad, size_t dw_bytes, TcPlan* pl) {
  (void)nslab;
  // N chunks of equal size Nc <= 128 (one CTA row per chunk; each repeats the depthwise stage): fewest chunks that fit
  for (int nch = 1; nch <= 8; ++nch) {
    if ((Npad / 16) % nch) continue;
    const int Nc = Npad / nch;
    if (Nc > 128) continue;
    const size_t fixed = (size_t)2 * (TC_STAGE_BYTES + 3 * Nc * 64) + TC_AUX_BYTES + dw_bytes + 1024;
    if (fixed + 2 * (size_t)pl->halo_bytes > (size_t)TC_SMEM_BUDGET) continue;
    pl->Nc = Nc; pl->nchunks = nch; pl->wstream = 1; pl->stages = 2;
    // The ring depth must be EVEN: the two producer groups take alternate K-slabs, so with an even depth a slot always belongs
    // to the same group and a group's wait for round r of a slot follows its own round r-1.  With 3 slots a group could reach
    // round r while the OTHER group's round r-1 load was still in flight, and its parity wait on the not-yet-flipped mbarrier
    // passed spuriously (phase aliasing: ~1 deadlock per 1000 launches, found with the post-mortem words of mbar_wait).
    pl->halo_slots = fixed + 4 * (size_t)pl->halo_bytes <= (size_t)TC_SMEM_BUDGET ? 4 : 2;
    pl->smem = fixed + (size_t)pl->halo_slots * pl->halo_bytes;
    return true;
  }
  return false;
}

// MODE 0 with TMA-fed A tiles (or MODE 1, generic gather) and a K too long for a resident weight image at one N chunk: every
// stage carries its own K-slab of W (three splits) next to the A tile.  Returns false when it does not fit.
static bool tc_plan_stream0(int N, int anchors, int max_chunks, TcPlan* pl) {
  const int Npad = (N + 15) / 16 * 16;
  for (int nch = 1; nch < max_chunks; ++nch) {        // only if it needs fewer N chunks than the resident plan
    const int Nc = ((Npad / 16 + nch - 1) / nch) * 16;      // the last chunk may be partial: its missing rows are never copied
    if (Nc > 128 || (nch - 1) * Nc >= Npad) continue;
    const size_t dense_bytes = tc_dense_epi(N, anchors, Nc, nch) ? (size_t)4 * 32 * Nc * 4 : 0;
    const size_t fixed = TC_AUX_BYTES + dense_bytes + 1024;
    const size_t per_stage = (size_t)TC_STAGE_BYTES + (size_t)3 * Nc * 64;
    int stages = (int)((TC_SMEM_BUDGET - fixed) / per_stage);
    if (stages < 2) continue;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    pl->Nc = Nc; pl->nchunks = nch; pl->wstream = 1; pl->stages = stages; pl->halo_slots = 0; pl->epi2 = 0;
    pl->smem = fixed + stages * per_stage;
    return true;
  }
  return false;
}

static bool tc_plan(int K, int N, int anchors, int mode, int dw_ks, int Hout, int Wout, TcPlan* pl, bool want_epi2 = false, int dw_stride = 1) {
  if (K < 8 || N < 8) return false;
  const int nslab = (K + 31) / 32;
  const int Npad = (N + 15) / 16 * 16;
  if (mode == 2) {
    if (dw_ks != 3 && dw_ks != 5) return false;
    if (dw_stride != 1 && dw_stride != 2) return false;
    tc_pick_tile(dw_ks, dw_stride, Hout, Wout, pl);
    if (pl->halo_pix == 0) return false;
  }
  const size_t dw_bytes = mode == 2 ? (size_t)(dw_ks * dw_ks + 1) * nslab * 32 * 4 : 0;
  for (int nch = 1; nch <= 8; ++nch) {
    int Nc = ((Npad / 16 + nch - 1) / nch) * 16;
    if (Nc > 128) continue;                      // 2 buffers x (main + correction) accumulators x Nc <= 512 TMEM columns
    const size_t wbytes = (size_t)3 * nslab * Nc * 64;
    const size_t dense_bytes = tc_dense_epi(N, anchors, Nc, nch) ? (size_t)4 * 32 * Nc * 4 : 0;
    size_t fixed = wbytes + TC_AUX_BYTES + dense_bytes + 1024;
    pl->Nc = Nc; pl->nchunks = nch; pl->wstream = 0; pl->halo_slots = 0;
    if (mode == 2) {                             // halo ring (2..4 slots) + depthwise weights, 2 A stages
      fixed += dw_bytes + 2 * TC_STAGE_BYTES;
      if (fixed + 2 * pl->halo_bytes > (size_t)TC_SMEM_BUDGET) continue;
      // splitting N over CTAs repeats the depthwise stage in every chunk: when the whole N fits one CTA's TMEM, streaming
      // the weight slabs (below) is cheaper than chunking
      if (nch > 1 && Npad <= 128 && (N & 3) == 0 && tc_plan_stream(nslab, Npad, dw_bytes, pl)) return true;
      pl->halo_slots = fixed + 4 * (size_t)pl->halo_bytes <= (size_t)TC_SMEM_BUDGET ? 4 : 2;      // even: see tc_plan_stream
      pl->stages = 2; pl->smem = fixed + (size_t)pl->halo_slots * pl->halo_bytes;
      return true;
    }
    if (fixed + 2 * TC_STAGE_BYTES > (size_t)TC_SMEM_BUDGET) continue;
    // second epilogue group (MODE 0 + TMA, plain vector epilogue): four more transpose staging buffers, if 3 stages still fit
    const size_t epi2_bytes = (size_t)4 * TC_STG_BYTES;
    pl->epi2 = 0;
    if (want_epi2 && mode == 0 && dense_bytes == 0 && (N & 3) == 0 && anchors <= 1 &&
        fixed + epi2_bytes + 3 * TC_STAGE_BYTES <= (size_t)TC_SMEM_BUDGET) {
      pl->epi2 = 1;
      fixed += epi2_bytes;
    }
    int stages = (int)((TC_SMEM_BUDGET - fixed) / TC_STAGE_BYTES);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    pl->stages = stages; pl->smem = fixed + (size_t)stages * TC_STAGE_BYTES;
    return true;
  }
  if (mode == 2 && (N & 3) == 0) return tc_plan_stream(nslab, Npad, dw_bytes, pl);
  // MODE 0 needs the TMA-fed A path for streamed weights (checked at launch); MODE 1 (dense KxK with a long K, e.g. the 3x3
  // convs of the YOLOLiteMS FPN) streams them next to its gathered A stages
  if (mode == 0 || mode == 1) return tc_plan_stream0(N, anchors, 9, pl);
  return false;
}

bool tc_supported(int K, int N, int anchors, int mode, int dw_ks, int Hout, int Wout, int dw_stride) {
  TcPlan pl;
  return tc_plan(K, N, anchors, mode, dw_ks, Hout, Wout, &pl, false, dw_stride);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

int make_tmap_f32(CUtensorMap* tm, const float* base, int rank, const unsigned long long* dims, const unsigned long long* strides,
                  const unsigned int* box, bool swizzle128) {
  EncodeTiledFn enc = encode_tiled_fn();
  YL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  YL_REQUIRE(rank >= 1 && rank <= 5, "tensor map rank 1..5");
  cuuint64_t d[5], st[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; if (i + 1 < rank) st[i] = strides[i]; }
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, st, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  return 0;
}

// 4-D map over an NHWC fp32 activation: box = 32 channels x halo_w x halo_h x 1 image, dense [hy][hx][32] in shared memory;
// coordinates outside the tensor (negative / past the edge / channels past C) read as zero.
static int make_halo_tmap(CUtensorMap* tm, const float* in, int B, int H, int W, int C, int halo_w, int halo_h) {
  EncodeTiledFn enc = encode_tiled_fn();
  YL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)halo_w, (cuuint32_t)halo_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for the halo tile");
  return 0;
}

// 2-D map over a pointwise conv's A matrix [M][C] fp32: box = 32 k x 128 rows, SWIZZLE_128B = the K-major UMMA layout
static int make_a_tmap(CUtensorMap* tm, const float* in, long long M, int C) {
  EncodeTiledFn enc = encode_tiled_fn();
  YL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)C * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for the A tile");
  return 0;
}

int tc_prepare(const ConvParams& c, const float* wimg, int mode, int sm_count, TcLaunch* L) {
  TcParams& p = L->p;
  p = TcParams{};
  p.c = c;
  p.wimg = wimg;
  p.mode = mode;
  p.dbg = debug_words();
  YL_REQUIRE(mode >= 0 && mode <= 2, "tcgen05 conv modes: 0 pointwise, 1 dense KxK, 2 depthwise -> pointwise");
  const bool tap = mode == 1 && c.wt_layout == 1;          // per-tap padded K axis: dense k x k stride-1 conv fed by shifted TMA boxes
  p.spt = tap ? (c.Cin + 31) / 32 : 0;
  p.K = (mode == 0 || mode == 2) ? c.Cin : tap ? c.KS * c.KS * p.spt * 32 : c.KS * c.KS * c.Cin;
  YL_REQUIRE(c.wt_layout == 0 || (tap && c.stride == 1 && c.pad == c.KS / 2 && !c.up && c.anchors <= 1 && (c.Cout & 3) == 0 &&
                                  (reinterpret_cast<uintptr_t>(c.in) & 15) == 0),
             "per-tap padded weight image: dense k x k stride-1 conv on a 16-byte aligned input");
  p.nslab = (p.K + 31) / 32;
  p.Npad = (c.Cout + 15) / 16 * 16;
  TcPlan pl;
  static const int tma_env = [] { const char* e = getenv("YL_TC_TMA"); return e ? atoi(e) : 1; }();
  static const int epi2_env = [] { const char* e = getenv("YL_TC_EPI2"); return e ? atoi(e) : 1; }();
  const bool tma_a = mode == 0 && tma_env && (reinterpret_cast<uintptr_t>(c.in) & 15) == 0;
  YL_REQUIRE(tc_plan(p.K, c.Cout, c.anchors, mode, mode == 2 ? c.KS : 0, c.Hout, c.Wout, &pl, tma_a && epi2_env, mode == 2 ? c.stride : 1),
             "shape does not fit the tcgen05 conv kernel");
  if (tma_a && pl.nchunks > 1) {             // N was chunked only because the weights do not fit: stream them instead
    TcPlan ps;
    if (tc_plan_stream0(c.Cout, c.anchors, pl.nchunks, &ps)) pl = ps;
  }
  YL_REQUIRE(!(mode == 0 && pl.wstream) || tma_a, "streamed weights need a 16-byte aligned input (TMA)");
  p.epi2 = pl.epi2;
  p.prod_warps = pl.epi2 ? 4 : TC_PROD_WARPS;
  p.Nc = pl.Nc; p.nchunks = pl.nchunks; p.stages = pl.stages; p.halo_slots = pl.halo_slots; p.wstream = pl.wstream;
  p.tile_w = pl.tile_w; p.tile_h = pl.tile_h; p.halo_w = pl.halo_w; p.halo_pix = pl.halo_pix; p.halo_bytes = pl.halo_bytes;
  p.dw_stride = mode == 2 ? c.stride : 1;
  p.stg_stride = TC_STG_BYTES;
  p.dense_epi = tc_dense_epi(c.Cout, c.anchors, p.Nc, p.nchunks) ? 1 : 0;
  YL_REQUIRE((c.Cin & 3) == 0, "tcgen05 conv needs Cin % 4 == 0");
  p.M = (long long)c.B * c.Hout * c.Wout;
  p.num_tiles = (int)((p.M + TC_BM - 1) / TC_BM);
  // MODE 1 with a small Cin: whole-Cin halo tiles by TMA + shared-memory im2col (see the producer branch)
  size_t smem_override = 0;
  if (tap) {
    p.tap_tma = 1;
    p.tile_w = TC_TILE_W; p.tile_h = TC_TILE_H;
    p.tiles_x = (c.Wout + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (c.Hout + p.tile_h - 1) / p.tile_h;
    p.num_tiles = c.B * p.tiles_x * p.tiles_y;
    p.dense_epi = 0;
  }
  if (!tap && mode == 1 && tma_env && pl.nchunks == 1 && !c.up && c.anchors <= 1 && (c.Cout & 3) == 0 && (c.Cin & 3) == 0 && c.Cin <= 64 &&
      (c.stride == 1 || c.stride == 2) && c.pad == c.KS / 2 && (reinterpret_cast<uintptr_t>(c.in) & 15) == 0) {
    const int tw = TC_TILE_W, th = TC_TILE_H;
    const int hw_ = (tw - 1) * c.stride + c.KS, hh_ = (th - 1) * c.stride + c.KS;
    const size_t hexact = (size_t)hw_ * hh_ * c.Cin * 4;
    const size_t hbytes = (hexact + 127) / 128 * 128;
    const size_t fixed = (size_t)3 * p.nslab * pl.Nc * 64 + TC_AUX_BYTES + 1024;
    if (!pl.wstream && hw_ <= 256 && hh_ <= 256 && fixed + 2 * hbytes + 2 * TC_STAGE_BYTES <= (size_t)TC_SMEM_BUDGET) {
      p.tma_a = 1;
      pl.epi2 = 0;
      p.tile_w = tw; p.tile_h = th; p.halo_w = hw_; p.halo_pix = hw_ * hh_; p.halo_bytes = (int)hbytes; p.halo_tx = (int)hexact;
      p.halo_slots = 2;
      int stages = (int)((TC_SMEM_BUDGET - fixed - 2 * hbytes) / TC_STAGE_BYTES);
      p.stages = stages > TC_MAX_STAGES ? TC_MAX_STAGES : stages;
      if (fixed + 3 * hbytes + (size_t)p.stages * TC_STAGE_BYTES <= (size_t)TC_SMEM_BUDGET) p.halo_slots = 3;
      smem_override = fixed + (size_t)p.halo_slots * hbytes + (size_t)p.stages * TC_STAGE_BYTES;
      p.tiles_x = (c.Wout + tw - 1) / tw;
      p.tiles_y = (c.Hout + th - 1) / th;
      p.num_tiles = c.B * p.tiles_x * p.tiles_y;
      p.dense_epi = 0;
    }
  }
  if (mode == 2) {
    YL_REQUIRE(!c.up && c.anchors <= 1 && (c.Cout & 3) == 0, "fused depthwise epilogue takes no upsample/head layout");
    YL_REQUIRE((c.stride == 1 || c.stride == 2) && c.w2 && c.Hout == (c.Hin + 2 * (c.KS / 2) - c.KS) / c.stride + 1 &&
               c.Wout == (c.Win + 2 * (c.KS / 2) - c.KS) / c.stride + 1, "fused depthwise -> pointwise: depthwise stride 1 or 2");
    p.tiles_x = (c.Wout + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (c.Hout + p.tile_h - 1) / p.tile_h;
    p.num_tiles = c.B * p.tiles_x * p.tiles_y;
  }
  YL_REQUIRE(p.M * c.Cout < (1ll << 31), "output too large for 32-bit element offsets");
  YL_REQUIRE(p.M < (1ll << 31) - TC_BM, "too many output pixels for 32-bit row indices");
  YL_REQUIRE(!c.up || (long long)c.B * c.Hu * c.Wu * c.Cout < (1ll << 31), "upsample source too large for 32-bit offsets");
  int cols = 32;
  while (cols < 4 * p.Nc) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = smem_override ? smem_override : pl.smem;
  {
    static bool done[64];
    static std::mutex mtx;
    if (int rc = once_per_device(done, mtx, []() -> int {
          YL_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          YL_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          YL_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          YL_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          return 0;
        }))
      return rc;
  }
  int gx = sm_count / p.nchunks;
  if (gx < 1) gx = 1;
  if (gx > p.num_tiles) gx = p.num_tiles;
  L->grid = dim3(gx, p.nchunks);
  L->smem = smem;
  L->ks = c.KS;
  CUtensorMap& tmap = L->tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma_a) {
    if (int rc = make_a_tmap(&tmap, c.in, p.M, c.Cin)) return rc;
    p.tma_a = 1;
  }
  if (tap) {
    const unsigned long long dims[4] = {(unsigned long long)c.Cin, (unsigned long long)c.Win, (unsigned long long)c.Hin, (unsigned long long)c.B};
    const unsigned long long strides[3] = {(unsigned long long)c.Cin * 4, (unsigned long long)c.Win * c.Cin * 4, (unsigned long long)c.Hin * c.Win * c.Cin * 4};
    const unsigned int box[4] = {32, (unsigned)p.tile_w, (unsigned)p.tile_h, 1};
    if (int rc = make_tmap_f32(&tmap, c.in, 4, dims, strides, box, true)) return rc;
  }
  if (mode == 1 && p.tma_a) {
    const unsigned long long dims[4] = {(unsigned long long)c.Cin, (unsigned long long)c.Win, (unsigned long long)c.Hin, (unsigned long long)c.B};
    const unsigned long long strides[3] = {(unsigned long long)c.Cin * 4, (unsigned long long)c.Win * c.Cin * 4, (unsigned long long)c.Hin * c.Win * c.Cin * 4};
    const unsigned int box[4] = {(unsigned)c.Cin, (unsigned)p.halo_w, (unsigned)(p.halo_pix / p.halo_w), 1};
    if (int rc = make_tmap_f32(&tmap, c.in, 4, dims, strides, box, false)) return rc;
  }
  if (mode == 2) {
    YL_REQUIRE((reinterpret_cast<uintptr_t>(c.in) & 15) == 0 && pl.halo_w <= 256 && pl.halo_h <= 256, "halo tile does not fit a TMA box");
    if (int rc = make_halo_tmap(&tmap, c.in, c.B, c.Hin, c.Win, c.Cin, pl.halo_w, pl.halo_h)) return rc;
  }
  // output tensor map for the TMA-store epilogue: plain vector epilogue (N % 4 == 0, no head layout), no residual / upsample
  // source, every 32-column block inside this CTA's chunk, and in the spatial modes a tile width that divides 32
  CUtensorMap& omap = L->omap;
  memset(&omap, 0, sizeof(omap));
  static const int tmaout_env = [] { const char* e = getenv("YL_TC_TMAOUT"); return e ? atoi(e) : 1; }();
  const bool spatial = mode == 2 || (mode == 1 && (p.tma_a || p.tap_tma));
  if (tmaout_env && !p.dense_epi && (c.Cout & 3) == 0 && c.anchors <= 1 && !c.res && (!c.up || (mode == 0 && (reinterpret_cast<uintptr_t>(c.up) & 15) == 0)) && c.Cout >= 32 &&
      (p.nchunks == 1 || (p.Nc & 31) == 0) && (reinterpret_cast<uintptr_t>(c.out) & 15) == 0 && (!spatial || 32 % p.tile_w == 0)) {
    int rc;
    if (spatial) {
      const unsigned long long dims[4] = {(unsigned long long)c.Cout, (unsigned long long)c.Wout, (unsigned long long)c.Hout, (unsigned long long)c.B};
      const unsigned long long strides[3] = {(unsigned long long)c.Cout * 4, (unsigned long long)c.Wout * c.Cout * 4, (unsigned long long)c.Hout * c.Wout * c.Cout * 4};
      const unsigned int box[4] = {32, (unsigned)p.tile_w, (unsigned)(32 / p.tile_w), 1};
      rc = make_tmap_f32(&omap, c.out, 4, dims, strides, box, true);
    } else {
      const unsigned long long dims[2] = {(unsigned long long)c.Cout, (unsigned long long)p.M};
      const unsigned long long strides[1] = {(unsigned long long)c.Cout * 4};
      const unsigned int box[2] = {32, 32};
      rc = make_tmap_f32(&omap, c.out, 2, dims, strides, box, true);
    }
    if (rc) return rc;
    p.tma_out = 1;
  }
  return 0;
}

int tc_launch(const TcLaunch& L, cudaStream_t st, int pdl) {
  const int mode = L.p.mode;
  cudaError_t e;
  if (mode == 0) e = launch_ex(tc_conv_kernel<0, 0>, L.grid, TC_THREADS, L.smem, st, pdl, L.p, L.tmap, L.omap);
  else if (mode == 1) e = launch_ex(tc_conv_kernel<1, 0>, L.grid, TC_THREADS, L.smem, st, pdl, L.p, L.tmap, L.omap);
  else if (mode == 2 && L.ks == 3) e = launch_ex(tc_conv_kernel<2, 3>, L.grid, TC_THREADS, L.smem, st, pdl, L.p, L.tmap, L.omap);
  else e = launch_ex(tc_conv_kernel<2, 5>, L.grid, TC_THREADS, L.smem, st, pdl, L.p, L.tmap, L.omap);
  YL_CHECK_CUDA(e);
  return 0;
}

int launch_tc_conv(const ConvParams& c, const float* wimg, int mode, int sm_count, cudaStream_t st) {
  TcLaunch L;
  if (int rc = tc_prepare(c, wimg, mode, sm_count, &L)) return rc;
  return tc_launch(L, st, 0);
}

}  // namespace yl
