// Fused conv_stem (3x3 s2, Cin = 3, 32 ch, BN + ReLU) -> blocks.0.0 (3x3 s2 dense conv, BN + ReLU) on tcgen05, sm_100a.
// Reference: timm mobilenetv4_conv_small[_050] as called at scripts/model/model_v2.py:266-272; layer shapes from
// YoloLite_custom_training.ipynb:392-410.  The 13 MB/image stem activation never reaches HBM.
//
// Operands are error-compensated bf16 triples: every fp32 value x = x1 + x2 + x3 (x1 = bf16(x), x2 = bf16(x - x1),
// x3 = bf16(x - x1 - x2)), D = A1*W1 (main accumulator) + A1*W2 + A2*W1 + A2*W2 + A1*W3 + A3*W1 (correction accumulators),
// dropped terms are 2^-24 relative.  The three weight splits sit side by side along N ([W1 | W2 | W3]), so the six
// products take THREE instructions per k-step -- A1 x [W1|W2|W3], A2 x [W1|W2], A3 x [W1] -- into three adjacent
// accumulator blocks [main | corr | corr]: the kernel is bound by shared-memory operand reads, and this halves them.
//
// Per 16x8 tile of conv2 output pixels (one persistent CTA per SM, 17 warps, warp-specialised):
//   warps 8-15  producers, two groups of 4 warps that take alternate stem sub-tiles: the 3 x 67 x 36 fp32 input patch (NCHW)
//               arrives as one TMA box; every thread builds ONE whole row of the stem's im2col operand (K = 27 taps + a
//               constant-1 column that carries the folded BN bias, padded to 32; all patch offsets are immediates) for
//               five 128-row tiles covering the 33 x 17 halo of stem output pixels.  The rows are in PLANE order (row q
//               = pixel q of the parity planes below), so the epilogue's plane stores are bank-conflict free;
//   warp  16    one elected thread (elect.sync, see launch.cuh) issues tcgen05.mma.kind::f16 (bf16): GEMM1 = stem (M = 128,
//               N = 96|64|32, K = 32) into a ring of three TMEM accumulator triples, GEMM2 = conv2 as nine per-tap GEMMs
//               (M = 128, N = 3|2|1 x N2, K = 32), issued plane by plane as the epilogue completes the parity planes;
//   warps 0-7   epilogue: TMEM -> ReLU folded into round-toward-zero bf16 splits -> shared-memory halo stored as four parity
//               planes (row parity x column parity, per-plane full / free mbarriers); then conv2's accumulator -> bias + ReLU
//               (-> fused 16x16 pointwise conv) -> NHWC global stores.
// GEMM2 has NO im2col copy: for tap (ky, kx) the A operand of the 16x8 tile is the parity plane (ky&1, kx&1) shifted
// by (ky>>1, kx>>1) pixels, which a K-major SWIZZLE_64B descriptor addresses directly (8-row atoms = 8 consecutive
// output columns, stride-byte-offset = plane pitch).  The hardware swizzle is a function of the shared-memory ADDRESS
// (verified on B200, experiments/exp_desc.cu), so start addresses / SBOs that are not multiples of the 512 B atom work
// as long as the data was written with the same address-based XOR.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include <mutex>

#include "common.cuh"
#include "launch.cuh"

namespace yl {

constexpr int S2_EPI_WARPS = 8, S2_PROD_WARPS = 8, S2_MMA_WARP = 16, S2_THREADS = 17 * 32;
constexpr int S2_TH = 16, S2_TW = 8;                       // conv2 output tile (rows x cols) = 128 pixels
constexpr int S2_HH = 2 * S2_TH + 1, S2_HW = 2 * S2_TW + 1; // stem-output halo 33 x 17
constexpr int S2_HPIX = S2_HH * S2_HW;                     // 561
constexpr int S2_MT = (S2_HPIX + 127) / 128;               // 5 stem GEMM tiles
constexpr int S2_PR = 2 * S2_HH + 1;                       // 67 patch rows
constexpr int S2_PP = 36;                                  // patch pitch (floats): 16 B-aligned superset of the 35 columns
constexpr int S2_PLANE_BYTES = (S2_HPIX * 64 + 511) / 512 * 512;   // one bf16 split of the halo (32 ch x 2 B per pixel); a multiple of
                                                                   // the 512 B swizzle period so all three splits share one XOR pattern
constexpr int S2_A1_SPLIT = 128 * 64, S2_A1_STAGE = 3 * S2_A1_SPLIT;
constexpr int S2_PATCH_BYTES = 3 * S2_PR * S2_PP * 4;      // 28944
constexpr int S2_U8_PITCH = 128;                           // image mode: bytes per patch row (36 pixels x 3 = 108, padded to a 16 B multiple)
constexpr int S2_WST_BYTES = 3 * 32 * 64;                  // stem weights, three splits
constexpr int S2_ACC1_RING = 3;                            // stem accumulator ring: 3 x [main 32 | corr 32 | corr 32] TMEM columns
constexpr uint32_t S2_ACC1_COLS = 96, S2_ACC2_COL = S2_ACC1_RING * S2_ACC1_COLS, S2_ACC2_COLS = 96;   // conv2: 2 x 3*N2 (<= 96)
// parity planes (py, px): pixel offsets inside one split, pitch 9 (px = 0) or 8 (px = 1)
__host__ __device__ constexpr int s2_plane_off(int py, int px) { return py ? (px ? 433 : 289) : (px ? 153 : 0); }
constexpr int S2_ROWS = S2_MT * 128;                       // rows of the five stem GEMM tiles (561 real)
// GEMM1 row q (= pixel index inside the parity planes) -> halo pixel, packed hy | hx << 8 (0xFFFF: padding row)
__host__ __device__ inline unsigned s2_row_pixel(int q) {
  if (q >= S2_HPIX) return 0xFFFFu;
  const int py = q >= 289, px = py ? q >= 433 : q >= 153;
  const int l = q - s2_plane_off(py, px), pitch = px ? 8 : 9;
  const int i = l / pitch, jx = l - i * pitch;
  return (unsigned)(2 * i + py) | ((unsigned)(2 * jx + px) << 8);
}


namespace s2 {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000u)
        : "memory");
    if (spin > (1u << 22)) asm volatile("trap;");   // protocol error: fail loudly, never hang the device
  }
}
// K-major SWIZZLE_64B descriptor: 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint64_t desc64(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (4ull << 61);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split3(float a, float b, uint32_t& p1, uint32_t& p2, uint32_t& p3) { split3_pair<false>(a, b, p1, p2, p3); }
__device__ __forceinline__ void split3_relu(float a, float b, uint32_t& p1, uint32_t& p2, uint32_t& p3) { split3_pair<true>(a, b, p1, p2, p3); }
}  // namespace s2

template <int A1S>      // stem operand stages (2; 1 when the 32-channel conv2 weights leave no room for two)
__global__ void __launch_bounds__(S2_THREADS, 1) stem2_kernel(const Stem2Params p, const __grid_constant__ CUtensorMap tmap) {
  using namespace s2;
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- shared memory carve-up
  const int w2_bytes = 27 * p.N2 * 64;                              // conv2 weights: 3 splits x 9 taps x N2 rows x 64 B
  unsigned char* w2s = smem;                                        // 1024-aligned (N2 % 16 == 0 -> 27 * N2 * 64 % 1024 == 0)
  unsigned char* wst = w2s + w2_bytes;                              // 6 KB
  unsigned char* a1 = wst + S2_WST_BYTES;                           // 2 stages x 3 splits x 8 KB (1024-aligned)
  unsigned char* planes = a1 + A1S * S2_A1_STAGE;           // 3 splits x 35904 B
  float* patch = reinterpret_cast<float*>(planes + 3 * S2_PLANE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(patch) + S2_PATCH_BYTES);
  uint64_t* a1_full = bars;            // [2]  producers (4 warps of a group) -> MMA
  uint64_t* a1_empty = bars + 2;       // [2]  MMA commit -> producers
  uint64_t* acc1_full = bars + 4;      // [3]  MMA commit -> epilogue   (ring over the stem sub-tiles)
  uint64_t* acc1_free = bars + 7;      // [3]  epilogue (8 warps) -> MMA
  uint64_t* plane_full = bars + 10;    // [4]  epilogue (8 warps) -> MMA: parity plane pl of this tile's halo is complete
  uint64_t* plane_free = bars + 14;    // [4]  MMA commit -> epilogue: conv2's taps on plane pl have read it
  uint64_t* acc2_full = bars + 18;     // [2]  MMA commit -> epilogue
  uint64_t* acc2_free = bars + 20;     // [2]  epilogue (8 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  uint64_t* patch_full = bars + 23;    //      TMA complete_tx -> producers: the input patch of the next tile has landed
  float* pws = reinterpret_cast<float*>(bars + 24);       // fused pointwise: 16 x 16 weights + 16 biases
  unsigned short* lut = reinterpret_cast<unsigned short*>(pws + 16 * 16 + 16);   // [S2_ROWS] row -> halo pixel (s2_row_pixel)
  const bool has_pw = p.pw != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&a1_full[s]), S2_PROD_WARPS / 2); mbar_init(smem_u32(&a1_empty[s]), 1); }
    for (int j = 0; j < S2_ACC1_RING; ++j) { mbar_init(smem_u32(&acc1_full[j]), 1); mbar_init(smem_u32(&acc1_free[j]), S2_EPI_WARPS); }
    mbar_init(smem_u32(patch_full), 1);
    for (int pl = 0; pl < 4; ++pl) { mbar_init(smem_u32(&plane_full[pl]), S2_EPI_WARPS); mbar_init(smem_u32(&plane_free[pl]), 1); }
    // with the fused pointwise conv the two column-half warps of a lane quarter take alternate tiles (4 arrivals per buffer)
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&acc2_full[b]), 1); mbar_init(smem_u32(&acc2_free[b]), has_pw ? S2_EPI_WARPS / 2 : S2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == S2_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {   // resident weight images, copied verbatim (pre-split, pre-swizzled on the host)
    const int w24 = w2_bytes >> 4, total4 = (w2_bytes + S2_WST_BYTES) >> 4;
    const int stem_skip = p.in_u8 ? (S2_WST_BYTES >> 4) : 0;      // image mode: the second stem image (weights folded with 1/(255 std), -mean/std)
    for (int i = threadIdx.x; i < total4; i += S2_THREADS)
      reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(p.wimg) + (i < w24 ? i : i + stem_skip));
  }
  if (has_pw)
    for (int i = threadIdx.x; i < 16 * 16 + 16; i += S2_THREADS) pws[i] = __ldg(p.pw + i);
  for (int i = threadIdx.x; i < S2_ROWS; i += S2_THREADS) lut[i] = (unsigned short)s2_row_pixel(i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int tiles = p.num_tiles;
  const int per_img = p.tiles_x * p.tiles_y;
  pdl_launch_dependents();      // after the TMEM allocation (see tc_gemm.cu)
  pdl_wait();      // the prologue above read only weights; the network input / output buffers are touched from here on
  // TMEM columns: stem ring slot r: [96r, 96r + 96) = main | corr | corr (32 each); conv2 buffer b: 288 + 96b + {0, N2, 2*N2}
  constexpr uint32_t ACC2_COL = S2_ACC2_COL;

  if (warp >= S2_EPI_WARPS && warp < S2_EPI_WARPS + S2_PROD_WARPS) {
    // =============================== producers ===============================
    // Two groups of 4 warps take alternate stem sub-tiles n (group = n & 1 = the A1 stage when there are two); inside a group
    // every thread builds ONE whole im2col row: its 27 patch offsets are immediates, the row's 4 x 16 B chunks x 3 splits are
    // twelve conflict-free 16 B stores.
    const int t = threadIdx.x - 32 * S2_EPI_WARPS;          // 0..255
    const int grp = t >> 7, r = t & 127;
    // the 3 x 67 x 36 input patch (NCHW, zero outside the image = the stem's padding) is ONE TMA box over [B*3][H][W]
    auto issue_patch = [&](int tile) {
      if (p.in_u8) {
        // image mode: the 67-row BGR patch (bytes 96 tx - 16 ... + 128 of each row: 4 bytes of slack, then 36 pixels) by 16-byte
        // cp.async pieces, zero-filled outside the image; W % 16 == 0 keeps every piece aligned and entirely inside or outside
        if (tile < tiles) {
          const int b = tile / per_img, rem = tile - b * per_img;
          const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
          const int iy0 = 4 * S2_TH * ty - 3, xb0 = 3 * (4 * S2_TW * tx - 4) - 4;   // first row; first byte column (16 B aligned)
          const unsigned char* img = p.in_u8 + (size_t)b * p.H * p.W * 3;
          const uint32_t pbase = smem_u32(patch);
          for (int idx = t; idx < S2_PR * (S2_U8_PITCH / 16); idx += 32 * S2_PROD_WARPS) {
            const int row = idx >> 3, c16 = idx & 7;
            const int iy = iy0 + row, xb = xb0 + 16 * c16;
            const bool ok = iy >= 0 && iy < p.H && xb >= 0 && xb + 15 < p.W * 3;
            const unsigned char* src = ok ? img + (size_t)iy * p.W * 3 + xb : p.in_u8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(pbase + (uint32_t)(row * S2_U8_PITCH + 16 * c16)), "l"(src),
                         "r"(ok ? 16u : 0u) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        return;
      }
      if (t == 0 && tile < tiles) {
        const int b = tile / per_img, rem = tile - b * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int iy0 = 4 * S2_TH * ty - 3, ixa = 4 * S2_TW * tx - 4;
        const uint32_t bar = smem_u32(patch_full);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)S2_PATCH_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(patch)), "l"(&tmap), "r"(ixa), "r"(iy0), "r"(3 * b), "r"(bar) : "memory");
        // There is one patch buffer, so this load's latency is exposed once per tile: pull the NEXT tile's patch into L2 now, its
        // load then pays the L2 latency instead of the DRAM latency
        const int nt = tile + (int)gridDim.x;
        if (nt < tiles) {
          const int b2 = nt / per_img, rem2 = nt - b2 * per_img;
          const int ty2 = rem2 / p.tiles_x, tx2 = rem2 - ty2 * p.tiles_x;
          asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                       ::"l"(&tmap), "r"(4 * S2_TW * tx2 - 4), "r"(4 * S2_TH * ty2 - 3), "r"(3 * b2) : "memory");
        }
      }
    };
    issue_patch(blockIdx.x);
    uint32_t pphase = 0;
    uint32_t n = 0;                                          // running stem-tile counter -> group, A1 stage, phases
    const uint32_t rsw = (uint32_t)(r >> 1) & 3u, rbase = (uint32_t)r * 64u;      // this row's line and SWIZZLE_64B XOR
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      if (p.in_u8) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");       // the patch pieces of every producer have landed
      } else {
        mbar_wait(smem_u32(patch_full), pphase);             // this tile's patch has landed
        pphase ^= 1u;
      }
      const int trem = tile % per_img;
      const int sy0 = 2 * S2_TH * (trem / p.tiles_x) - 1, sx0 = 2 * S2_TW * (trem % p.tiles_x) - 1;   // stem-output coords of halo (0,0)
      for (int j = 0; j < S2_MT; ++j, ++n) {
        if ((int)(n & 1u) != grp) continue;
        const uint32_t stage = A1S == 2 ? (uint32_t)grp : 0u;
        const uint32_t px = lut[j * 128 + r];
        const bool live = px != 0xFFFFu;                      // rows past the 561 halo pixels are never read back
        const int hy = (int)(px & 0xFFu), hx = (int)(px >> 8);
        uint4 o1[4], o2[4], o3[4];
        if (p.in_u8) {
          // image mode: BGR bytes (channel ci of the RGB tensor is byte 2 - ci), the patch row starts 4 bytes before pixel ixa.
          // k = 27: constant 1 (bias + full mean/std term), 28: top-border row, 29: left-border column, 30: top-left corner
          // (they take back the padded taps' share of the mean/std term), 31: zero.  Integers 0..255 are exact in ONE bf16.
          if (live) {
            const unsigned char* pb = reinterpret_cast<const unsigned char*>(patch) + 2 * hy * S2_U8_PITCH + 6 * hx;
            float e[32];
#pragma unroll
            for (int k = 0; k < 27; ++k) {
              const int tap = k / 3, ci = k - tap * 3, ky = tap / 3, kx = tap - ky * 3;
              e[k] = (float)pb[ky * S2_U8_PITCH + 4 + (kx + 1) * 3 + (2 - ci)];
            }
            const float top = (sy0 + hy == 0) ? 1.f : 0.f, left = (sx0 + hx == 0) ? 1.f : 0.f;
            e[27] = 1.f; e[28] = top; e[29] = left; e[30] = top * left; e[31] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t w[4];
#pragma unroll
              for (int m = 0; m < 4; ++m)      // exact: pack the high halves of the two fp32 patterns
                w[m] = __byte_perm(__float_as_uint(e[8 * c + 2 * m]), __float_as_uint(e[8 * c + 2 * m + 1]), 0x7632);
              o1[c] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        } else if (live) {
          // taps kx = 0, 1, 2 are patch columns 2 hx + 1 .. + 3.  Rows come in plane order, so consecutive lanes step hx by 2 = one
          // 16 B group of the patch row: even hx takes .y .z .w of ONE conflict-free 16 B load, odd hx .w of that group and the
          // first 8 B of the next (a 4 B load per tap would be a 4-way bank conflict at this 16 B lane stride)
          const float* pb = patch + 2 * hy * S2_PP + 4 * (hx >> 1);
          float e[32];
          if (hx & 1) {
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const float* q = pb + (ci * S2_PR + ky) * S2_PP;
                const float4 a = *reinterpret_cast<const float4*>(q);
                const float2 b = *reinterpret_cast<const float2*>(q + 4);
                e[(ky * 3 + 0) * 3 + ci] = a.w; e[(ky * 3 + 1) * 3 + ci] = b.x; e[(ky * 3 + 2) * 3 + ci] = b.y;
              }
          } else {
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const float4 a = *reinterpret_cast<const float4*>(pb + (ci * S2_PR + ky) * S2_PP);
                e[(ky * 3 + 0) * 3 + ci] = a.y; e[(ky * 3 + 1) * 3 + ci] = a.z; e[(ky * 3 + 2) * 3 + ci] = a.w;
              }
          }
          e[27] = 1.f; e[28] = 0.f; e[29] = 0.f; e[30] = 0.f; e[31] = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            split3(e[8 * c + 0], e[8 * c + 1], o1[c].x, o2[c].x, o3[c].x);
            split3(e[8 * c + 2], e[8 * c + 3], o1[c].y, o2[c].y, o3[c].y);
            split3(e[8 * c + 4], e[8 * c + 5], o1[c].z, o2[c].z, o3[c].z);
            split3(e[8 * c + 6], e[8 * c + 7], o1[c].w, o2[c].w, o3[c].w);
          }
        }
        // the MMAs that read this stage (sub-tile n - A1S) are done; their commits alternate between the two barriers, so a
        // group never waits on a barrier whose phase the OTHER group's sub-tile could have advanced (no parity aliasing)
        if (n >= (uint32_t)A1S) mbar_wait(smem_u32(&a1_empty[(n - A1S) & 1u]), ((n - A1S) >> 1) & 1u);
        if (live) {
          unsigned char* st = a1 + stage * S2_A1_STAGE + rbase;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t off = (((uint32_t)c ^ rsw) << 4);
            *reinterpret_cast<uint4*>(st + off) = o1[c];
            if (!p.in_u8) {
              *reinterpret_cast<uint4*>(st + S2_A1_SPLIT + off) = o2[c];
              *reinterpret_cast<uint4*>(st + 2 * S2_A1_SPLIT + off) = o3[c];
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&a1_full[stage]));
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");         // every producer is done reading the patch
      issue_patch(tile + gridDim.x);
    }
  } else if (warp == S2_MMA_WARP) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      auto idesc = [](uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24); };
      const uint32_t i1a = idesc(96), i1b = idesc(64), i1c = idesc(32);
      const uint32_t N2 = (uint32_t)p.N2, i2a = idesc(3 * N2), i2b = idesc(2 * N2), i2c = idesc(N2);
      // Descriptor bases: every operand address below is base + a compile-time (or per-launch) multiple of 16 B, and all
      // of shared memory fits the 14-bit start-address field, so each MMA's descriptors cost one 64-bit add.
      const uint64_t dA1 = desc64(smem_u32(a1), 512), dWs = desc64(smem_u32(wst), 512), dW2 = desc64(smem_u32(w2s), 512);
      const uint64_t dP[2] = {desc64(smem_u32(planes), 9u * 64u), desc64(smem_u32(planes), 8u * 64u)};   // SBO = plane pitch
      const uint32_t w2_tap16 = 3u * N2 * 4u;                                                            // in 16 B units
      uint32_t n = 0, slot = 0, sphase = 0;
      auto gemm1 = [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j, ++n) {
          const uint32_t stage = A1S == 2 ? (n & 1u) : 0u, aph = A1S == 2 ? ((n >> 1) & 1u) : (n & 1u);
          mbar_wait(smem_u32(&acc1_free[slot]), sphase ^ 1u);
          mbar_wait(smem_u32(&a1_full[stage]), aph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d0 = tmem_base + S2_ACC1_COLS * slot;
          const uint64_t da0 = dA1 + (uint64_t)(stage * (S2_A1_STAGE >> 4));
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t db = dWs + (uint64_t)((ks * 32) >> 4);
            mma_bf16(d0, da0 + (uint64_t)((0 * S2_A1_SPLIT + ks * 32) >> 4), db, i1a, ks > 0);        // A1 x [W1|W2|W3]
            if (!p.in_u8) {                                                                           // image mode: A = A1 exactly
              mma_bf16(d0 + 32u, da0 + (uint64_t)((1 * S2_A1_SPLIT + ks * 32) >> 4), db, i1b, 1u);    // A2 x [W1|W2]
              mma_bf16(d0 + 64u, da0 + (uint64_t)((2 * S2_A1_SPLIT + ks * 32) >> 4), db, i1c, 1u);    // A3 x [W1]
            }
          }
          mma_commit(smem_u32(&a1_empty[n & 1u]));        // the barriers alternate with n even when there is one stage (see producers)
          mma_commit(smem_u32(&acc1_full[slot]));
          if (++slot == S2_ACC1_RING) { slot = 0; sphase ^= 1u; }
        }
      };
      // conv2 of tile `it`, the taps that read parity plane pl = 2 py + px: (0,0): taps (0,0) (0,2) (2,0) (2,2); (0,1): (0,1) (2,1);
      // (1,0): (1,0) (1,2); (1,1): (1,1).  The planes complete in this order (GEMM1 rows are in plane order), so conv2 starts on
      // plane 0 while the epilogue still writes planes 1-3, and releases every plane as soon as its own taps have read it: neither
      // conv2 nor the wait for it sits between two tiles any more.
      auto gemm2_plane = [&](uint32_t it, int pl) {
        const uint32_t b = it & 1u;
        if (pl == 0) mbar_wait(smem_u32(&acc2_free[b]), ((it >> 1) & 1u) ^ 1u);
        mbar_wait(smem_u32(&plane_full[pl]), it & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d0 = tmem_base + ACC2_COL + S2_ACC2_COLS * b;
        const int py = pl >> 1, px = pl & 1;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          if ((ky & 1) != py || (kx & 1) != px) continue;
          const int pw = px ? 8 : 9;
          const int aoff = (s2_plane_off(py, px) + (ky >> 1) * pw + (kx >> 1)) * 64;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t db = dW2 + (uint64_t)((uint32_t)tap * w2_tap16 + (uint32_t)(ks * 2));
            const uint32_t acc = (pl > 0 || tap > 0 || ks > 0) ? 1u : 0u;      // tap (0,0) of plane 0 is the first
            mma_bf16(d0, dP[px] + (uint64_t)((0 * S2_PLANE_BYTES + aoff + ks * 32) >> 4), db, i2a, acc);
            mma_bf16(d0 + N2, dP[px] + (uint64_t)((1 * S2_PLANE_BYTES + aoff + ks * 32) >> 4), db, i2b, 1u);
            mma_bf16(d0 + 2u * N2, dP[px] + (uint64_t)((2 * S2_PLANE_BYTES + aoff + ks * 32) >> 4), db, i2c, 1u);
          }
        }
        mma_commit(smem_u32(&plane_free[pl]));
        if (pl == 3) mma_commit(smem_u32(&acc2_full[b]));
      };
      // Issue order = the order in which the epilogue enables the pieces (it drains sub-tile j of a tile, then completes plane
      // j - 1): G1(it,3) G1(it,4) G2(it).p0 G1(it+1,0) G2(it).p1 G1(it+1,1) G2(it).p2 G1(it+1,2) G2(it).p3; the accumulator ring
      // holds 3 sub-tiles, so G1(., j) follows the drain of the sub-tile three before it.
      uint32_t it = 0;
      gemm1(0, 3);
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const bool more = tile + (int)gridDim.x < tiles;
        gemm1(3, 5);
        gemm2_plane(it, 0);
        if (more) gemm1(0, 1);
        gemm2_plane(it, 1);
        if (more) gemm1(1, 2);
        gemm2_plane(it, 2);
        if (more) gemm1(2, 3);
        gemm2_plane(it, 3);
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    const int q4 = warp & 3, half = warp >> 2;                 // TMEM lane quarter, column half
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    auto out_epi = [&](uint32_t it, int tile) {
      const uint32_t b = it & 1u;
      if (has_pw) {
        // conv2 (16 ch, bias + act) -> pointwise 16 -> 16 (bias + act) per pixel in registers: this warp reads all 16 columns of
        // its 32 pixels; the other column-half warp of the lane quarter takes the next tile
        if ((uint32_t)half != b) return;
        mbar_wait(smem_u32(&acc2_full[b]), (it >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int bi = tile / per_img, rem = tile - bi * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int r = q4 * 32 + lane;
        const int y = ty * S2_TH + (r >> 3), x = tx * S2_TW + (r & 7);
        float xin[16];
        {
          uint32_t m[16], k[16], k2[16];
          const uint32_t col = ACC2_COL + S2_ACC2_COLS * b;
          tmem_ld16(lane_addr + col, m);
          tmem_ld16(lane_addr + col + 16u, k);
          tmem_ld16(lane_addr + col + 32u, k2);
          tmem_ld_wait();
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&acc2_free[b]));
#pragma unroll
          for (int c = 0; c < 16; ++c)
            xin[c] = act_fn(__uint_as_float(m[c]) + (__uint_as_float(k[c]) + __uint_as_float(k2[c])) + __ldg(p.bias2 + c), p.act);
        }
        float yo[16];
        {
          f32x2 y2[8];                                          // packed pairs (FFMA2)
#pragma unroll
          for (int n = 0; n < 8; ++n) y2[n] = pack2(pws[256 + 2 * n], pws[256 + 2 * n + 1]);
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const f32x2 xk = pack2(xin[kk], xin[kk]);
#pragma unroll
            for (int n4 = 0; n4 < 4; ++n4) {
              const float4 w4 = *reinterpret_cast<const float4*>(pws + kk * 16 + n4 * 4);      // same address in every lane: broadcast
              y2[2 * n4] = fma2(xk, pack2(w4.x, w4.y), y2[2 * n4]);
              y2[2 * n4 + 1] = fma2(xk, pack2(w4.z, w4.w), y2[2 * n4 + 1]);
            }
          }
#pragma unroll
          for (int n = 0; n < 8; ++n) unpack2(y2[n], yo[2 * n], yo[2 * n + 1]);
        }
        if (y < p.Ho && x < p.Wo) {
          float* dst = p.out + (((size_t)bi * p.Ho + y) * p.Wo + x) * 16;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(dst + 4 * g) = make_float4(act_fn(yo[4 * g], p.pw_act), act_fn(yo[4 * g + 1], p.pw_act),
                                                                  act_fn(yo[4 * g + 2], p.pw_act), act_fn(yo[4 * g + 3], p.pw_act));
        }
        return;
      }
      mbar_wait(smem_u32(&acc2_full[b]), (it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int bi = tile / per_img, rem = tile - bi * per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int r = q4 * 32 + lane;
      const int y = ty * S2_TH + (r >> 3), x = tx * S2_TW + (r & 7);
      const int ncol = p.N2 >> 1;                               // columns per half: 8 (N2 = 16) or 16 (N2 = 32)
      for (int c0 = 0; c0 < ncol; c0 += 8) {
        uint32_t m[8], k[8], k2[8];
        const uint32_t col = ACC2_COL + S2_ACC2_COLS * b + (uint32_t)(half * ncol + c0);
        tmem_ld8(lane_addr + col, m);
        tmem_ld8(lane_addr + col + (uint32_t)p.N2, k);
        tmem_ld8(lane_addr + col + 2u * (uint32_t)p.N2, k2);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) k[g] = __float_as_uint(__uint_as_float(k[g]) + __uint_as_float(k2[g]));
        const int nn = half * ncol + c0;
        if (y < p.Ho && x < p.Wo) {
          float* dst = p.out + (((size_t)bi * p.Ho + y) * p.Wo + x) * p.Cout + nn;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (nn + 4 * g < p.Cout) {
              const float4 bia = __ldg(reinterpret_cast<const float4*>(p.bias2 + nn + 4 * g));
              float4 o;
              o.x = act_fn(__uint_as_float(m[4 * g + 0]) + __uint_as_float(k[4 * g + 0]) + bia.x, p.act);
              o.y = act_fn(__uint_as_float(m[4 * g + 1]) + __uint_as_float(k[4 * g + 1]) + bia.y, p.act);
              o.z = act_fn(__uint_as_float(m[4 * g + 2]) + __uint_as_float(k[4 * g + 2]) + bia.z, p.act);
              o.w = act_fn(__uint_as_float(m[4 * g + 3]) + __uint_as_float(k[4 * g + 3]) + bia.w, p.act);
              *reinterpret_cast<float4*>(dst + 4 * g) = o;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc2_free[b]));
    };
    uint32_t it = 0, slot = 0, sphase = 0;
    int prev_tile = -1;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int rem = tile % per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int sy0 = 2 * S2_TH * ty - 1, sx0 = 2 * S2_TW * tx - 1;
      // only tiles on the image border have halo pixels outside the stem output (conv2's zero padding)
      const bool border = sy0 < 0 || sx0 < 0 || sy0 + S2_HH > p.Hs || sx0 + S2_HW > p.Ws;
      for (int j = 0; j < S2_MT; ++j) {
        mbar_wait(smem_u32(&acc1_full[slot]), sphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[16];                                          // main + corr + corr of this thread's 16 channels
        {
          const uint32_t ta = lane_addr + S2_ACC1_COLS * slot + 16u * (uint32_t)half;
          uint32_t m[16], k[16], k2[16];
          tmem_ld16(ta, m);
          tmem_ld16(ta + 32u, k);
          tmem_ld16(ta + 64u, k2);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; c += 2)      // main + (corr + corr), as packed pairs (FADD2)
            unpack2(add2(pack2(__uint_as_float(m[c]), __uint_as_float(m[c + 1])),
                         add2(pack2(__uint_as_float(k[c]), __uint_as_float(k[c + 1])), pack2(__uint_as_float(k2[c]), __uint_as_float(k2[c + 1])))),
                    v[c], v[c + 1]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc1_free[slot]));
        if (++slot == S2_ACC1_RING) { slot = 0; sphase ^= 1u; }
        // sub-tile j is the first to write plane j (rows 128 j ... straddle planes j - 1 and j): conv2 of the previous tile has read it
        if (j < 4) mbar_wait(smem_u32(&plane_free[j]), (it & 1u) ^ 1u);
        // GEMM1 row q IS pixel q of the parity planes (s2_row_pixel): consecutive lanes write consecutive 64 B plane rows, whose
        // SWIZZLE_64B chunk positions make every 8-lane group of a 16 B store hit eight different bank groups
        const uint32_t q = (uint32_t)(j * 128 + q4 * 32 + lane);
        if (q < (uint32_t)S2_HPIX) {
          if (border) {
            const uint32_t pxl = lut[q];
            const int sy = sy0 + (int)(pxl & 0xFFu), sx = sx0 + (int)(pxl >> 8);
            if (!(sy >= 0 && sy < p.Hs && sx >= 0 && sx < p.Ws)) {
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] = 0.f;
            }
          }
          uint32_t o1[8], o2[8], o3[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) split3_relu(v[2 * c], v[2 * c + 1], o1[c], o2[c], o3[c]);   // ReLU; the bias rode in the GEMM (k = 27)
          const uint32_t sw = (q >> 1) & 3u;                   // planes is 1024 B aligned: the XOR is address bits 7-8 = (q >> 1) & 3
          unsigned char* row = planes + q * 64u;
          const uint32_t d0 = (((uint32_t)(2 * half)) ^ sw) << 4, d1 = (((uint32_t)(2 * half + 1)) ^ sw) << 4;
          *reinterpret_cast<uint4*>(row + d0) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
          *reinterpret_cast<uint4*>(row + d1) = make_uint4(o1[4], o1[5], o1[6], o1[7]);
          *reinterpret_cast<uint4*>(row + S2_PLANE_BYTES + d0) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
          *reinterpret_cast<uint4*>(row + S2_PLANE_BYTES + d1) = make_uint4(o2[4], o2[5], o2[6], o2[7]);
          *reinterpret_cast<uint4*>(row + 2 * S2_PLANE_BYTES + d0) = make_uint4(o3[0], o3[1], o3[2], o3[3]);
          *reinterpret_cast<uint4*>(row + 2 * S2_PLANE_BYTES + d1) = make_uint4(o3[4], o3[5], o3[6], o3[7]);
        }
        if (j >= 1) {                                          // plane j - 1 is complete (as far as this warp's rows go)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&plane_full[j - 1]));
        }
      }
      if (prev_tile >= 0) out_epi(it - 1, prev_tile);         // runs while the tensor core works on this tile's conv2
      prev_tile = tile;
    }
    if (prev_tile >= 0) out_epi(it - 1, prev_tile);
  }

  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == S2_MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
static size_t stem2_smem_bytes(int N2, int a1_stages) {
  return (size_t)27 * N2 * 64 + S2_WST_BYTES + (size_t)a1_stages * S2_A1_STAGE + 3 * S2_PLANE_BYTES + S2_PATCH_BYTES + 256 +
         (16 * 16 + 16) * 4 + S2_ROWS * 2 + 1024;
}
static int stem2_a1_stages(int N2) { return stem2_smem_bytes(N2, 2) <= (size_t)227 * 1024 ? 2 : 1; }

bool stem2_supported(const ConvParams& c) {
  const int N2 = (c.Cout + 15) / 16 * 16;
  // image mode: the folded-normalisation border terms cover the top / left padding only (even H, W: no bottom / right padding)
  const bool in_ok = c.in_u8 ? ((reinterpret_cast<uintptr_t>(c.in_u8) & 15) == 0 && (c.Hin & 1) == 0 && (c.Win & 15) == 0)
                             : ((reinterpret_cast<uintptr_t>(c.in) & 15) == 0 && (c.Win & 3) == 0);
  return c.KS == 3 && c.stride == 2 && c.Cin == 32 && (c.Cout & 3) == 0 && N2 <= 32 && in_ok && stem2_smem_bytes(N2, stem2_a1_stages(N2)) <= (size_t)227 * 1024 && !c.res && !c.up &&
         c.anchors <= 1 && c.bias != nullptr;
}

// c: geometry of the SECOND conv as set up by engine.cu for YL_OP_STEM2 (Hin/Win = network input size, Hout/Wout = conv2 output)
int stem2_prepare(const ConvParams& c, const float* wimg, int sm_count, Stem2Launch* L) {
  Stem2Params& p = L->p;
  p = Stem2Params{};
  p.in = c.in; p.in_u8 = c.in_u8; p.wimg = wimg; p.bias2 = c.bias; p.out = c.out;
  p.B = c.B; p.H = c.Hin; p.W = c.Win;
  p.Hs = (c.Hin + 2 - 3) / 2 + 1; p.Ws = (c.Win + 2 - 3) / 2 + 1;
  p.Ho = c.Hout; p.Wo = c.Wout; p.Cout = c.Cout; p.N2 = (c.Cout + 15) / 16 * 16; p.act = c.act;
  YL_REQUIRE(stem2_supported(c), "shape does not fit the fused stem kernel");
  p.pw = c.b2; p.pw_act = c.act2;
  YL_REQUIRE(!p.pw || (p.Cout == 16 && p.N2 == 16), "fused pointwise after conv2 needs 16 channels");
  p.tiles_x = (p.Wo + S2_TW - 1) / S2_TW;
  p.tiles_y = (p.Ho + S2_TH - 1) / S2_TH;
  const long long nt = (long long)p.B * p.tiles_x * p.tiles_y;
  YL_REQUIRE(nt < (1ll << 31) && (long long)p.B * p.Ho * p.Wo * p.Cout < (1ll << 40), "too many tiles");
  p.num_tiles = (int)nt;
  p.a1_stages = stem2_a1_stages(p.N2);
  const size_t smem = stem2_smem_bytes(p.N2, p.a1_stages);
  {
    static bool done[64];
    static std::mutex mtx;
    if (int rc = once_per_device(done, mtx, []() -> int {
          YL_CHECK_CUDA(cudaFuncSetAttribute(stem2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          YL_CHECK_CUDA(cudaFuncSetAttribute(stem2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          return 0;
        }))
      return rc;
  }
  int gx = sm_count < p.num_tiles ? sm_count : p.num_tiles;
  L->grid = gx;
  L->smem = smem;
  CUtensorMap& tmap = L->tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (!p.in_u8) {
    const unsigned long long dims[3] = {(unsigned long long)p.W, (unsigned long long)p.H, (unsigned long long)p.B * 3};
    const unsigned long long strides[2] = {(unsigned long long)p.W * 4, (unsigned long long)p.H * p.W * 4};
    const unsigned int box[3] = {(unsigned)S2_PP, (unsigned)S2_PR, 3};
    if (int rc = make_tmap_f32(&tmap, p.in, 3, dims, strides, box, false)) return rc;
  }
  return 0;
}

// host-side view of the GEMM1 row order (tests/test_cabi.py checks that it enumerates the 33 x 17 halo plane by plane)
int stem2_row_pixel(int q) { return q >= 0 && q < S2_ROWS ? (int)s2_row_pixel(q) : -1; }

int stem2_launch(const Stem2Launch& L, cudaStream_t st, int pdl) {
  cudaError_t e;
  if (L.p.a1_stages == 2) e = launch_ex(stem2_kernel<2>, dim3(L.grid), S2_THREADS, L.smem, st, pdl, L.p, L.tmap);
  else e = launch_ex(stem2_kernel<1>, dim3(L.grid), S2_THREADS, L.smem, st, pdl, L.p, L.tmap);
  YL_CHECK_CUDA(e);
  return 0;
}

int launch_stem2(const ConvParams& c, const float* wimg, int sm_count, cudaStream_t st) {
  Stem2Launch L;
  if (int rc = stem2_prepare(c, wimg, sm_count, &L)) return rc;
  return stem2_launch(L, st, 0);
}

}  // namespace yl
