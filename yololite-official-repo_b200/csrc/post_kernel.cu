// Fused per-batch postprocess (sm_100a): sigmoid -> anchor-free decode -> score threshold -> class-wise NMS
// in ONE launch.
//
// Reference semantics restated (paths relative to the reference repo):
//   scripts/helpers/utils_ms.py:25-123   decode (center "v8": (sigmoid*2-0.5+grid)*stride, size "softplus")
//   tools/infer.py:466-475               score = sigmoid(obj) * max_c sigmoid(cls_c) (C==1: obj only),
//                                        first-max class on ties, keep iff score > conf (fp32, strict)
//   tools/infer.py:476-493 + :134-152    per class (ascending): torchvision.ops.nms, keep[:max_det]
//   torchvision.ops.nms (CPU)            stable descending sort; suppress iff inter/(a+b-inter) > iou
//                                        (fp32 IoU compared against the double threshold; NaN keeps)
//
// Kernel structure: grid = B * tiles_per_image CTAs.  Phase 1 (every CTA): stream one contiguous tile of
// 256 anchors x D logits HBM -> shared memory with coalesced 16 B loads, one thread per anchor scores it and
// appends survivors (64-bit sort key + decoded box) to the image's candidate list.  Phase 2 (the CTA that
// finishes an image last, found with a per-image ticket): sort the keys (class asc, score desc, anchor asc)
// in shared memory, run greedy NMS per class segment (one warp per segment, boxes and dead-bits in
// registers for segments <= 128), compact the survivors into the fixed-capacity output.
#include <mutex>

#include "common.cuh"
#include "launch.cuh"

namespace yl {

constexpr int POST_THREADS = 256;
constexpr int POST_TILE = 256;           // anchors per CTA
constexpr int POST_SMEM_KEYS = 8192;     // candidates sortable in shared memory
constexpr int POST_MAX_LEVELS = 8;
constexpr int REG_SEG = 4;               // register path handles segments up to 32*REG_SEG boxes

struct PostParams {
  const float* lvl[POST_MAX_LEVELS];
  int A[POST_MAX_LEVELS], Sh[POST_MAX_LEVELS], Sw[POST_MAX_LEVELS];
  int n_lvl[POST_MAX_LEVELS];        // anchors per image in the level
  int lvl_off[POST_MAX_LEVELS];      // first flat anchor index of the level
  int tile_off[POST_MAX_LEVELS + 1]; // first tile of the level within an image
  int n_levels;
  int B, D, C, img_size, N;
  float conf;
  double iou;
  int max_det, cap;
  int direct;             // 1: phase 1 reads the logits straight from global memory (only the objectness logit for the anchors
                          //    that fail the cheap early-out), 0: streams whole tiles through shared memory
  int smem_keys;          // candidates per image that can be sorted in shared memory
  // scratch
  int* count;             // [B] candidates
  int* done;              // [B] finished tiles
  unsigned long long* keys;   // [B][N]
  unsigned long long* keys2;  // [B][N]
  float4* cbox;           // [B][N] decoded box by anchor index
  unsigned char* gflags;  // [B][N]
  // outputs
  float* boxes; float* scores; long long* classes; long long* anchor_idx; int* counts;   // each may be null when `packed` is set
  float* packed;          // optional [B][cap + 1][6] fp32: row 0 = (count, overflow flag, total kept, 0, 0, 0), rows 1.. =
                          // (x1, y1, x2, y2, score, class): ONE buffer a multi-GPU gather can ship as it is
};

__device__ __forceinline__ float sigmoid_exact(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }
__device__ __forceinline__ float softplus_exact(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// key = class(12) | ~scorebits(32) | anchor(20)
__device__ __forceinline__ unsigned long long make_key(int cls, float score, int anchor) {
  return ((unsigned long long)cls << 52) | ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(score)) << 20) |
         (unsigned long long)anchor;
}
__device__ __forceinline__ int key_cls(unsigned long long k) { return (int)(k >> 52); }
__device__ __forceinline__ int key_anchor(unsigned long long k) { return (int)(k & 0xFFFFFu); }
__device__ __forceinline__ float key_score(unsigned long long k) {
  return __uint_as_float(0xFFFFFFFFu - (unsigned)((k >> 20) & 0xFFFFFFFFu));
}

__device__ __forceinline__ float4 decode_box(float tx, float ty, float tw, float th, int gx, int gy, float stride,
                                             float lim) {
  const float px = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sigmoid_exact(tx), 2.f), 0.5f), (float)gx), stride);
  const float py = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sigmoid_exact(ty), 2.f), 0.5f), (float)gy), stride);
  const float hw = __fmul_rn(__fmul_rn(softplus_exact(tw), stride), 0.5f);
  const float hh = __fmul_rn(__fmul_rn(softplus_exact(th), stride), 0.5f);
  float4 b;
  b.x = fminf(fmaxf(__fsub_rn(px, hw), 0.f), lim);
  b.y = fminf(fmaxf(__fsub_rn(py, hh), 0.f), lim);
  b.z = fminf(fmaxf(__fadd_rn(px, hw), 0.f), lim);
  b.w = fminf(fmaxf(__fadd_rn(py, hh), 0.f), lim);
  return b;
}

// suppress j by i?  torchvision CPU nms: ovr = inter / (iarea + areas[j] - inter); ovr > thr (double)
// Device-scope release / acquire fence (the ticket protocol needs no more: writes -> CTA barrier -> fence -> ticket atomic on the
// producer side, ticket atomic -> fence -> CTA barrier -> reads through L2 on the consumer side); __threadfence() is fence.sc.gpu.
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ bool iou_gt(const float4& a, float area_a, const float4& b, double thr) {
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return (double)ovr > thr;   // NaN -> false -> kept
}

__global__ void __launch_bounds__(POST_THREADS) post_kernel(PostParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_last, s_M, s_scan[POST_THREADS / 32], s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_per_image = p.tile_off[p.n_levels];
  const int b = blockIdx.x / tiles_per_image;
  const int t = blockIdx.x - b * tiles_per_image;
  int l = 0;
  while (l + 1 < p.n_levels && t >= p.tile_off[l + 1]) ++l;
  const int a0 = (t - p.tile_off[l]) * POST_TILE;            // first anchor of the tile within the level
  const int n_tile = min(POST_TILE, p.n_lvl[l] - a0);
  const int D = p.D, C = p.C;

  // ---------------- phase 1: stream the tile, score, append candidates
  {
    float* tile = reinterpret_cast<float*>(smem_raw);
    const float* src = p.lvl[l] + ((size_t)b * p.n_lvl[l] + a0) * D;
    const int n_el = n_tile * D;
    if (!p.direct) {
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = n_el >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(tile);
#pragma unroll 4
        for (int i = tid; i < n4; i += POST_THREADS) d4[i] = __ldcs(s4 + i);
        for (int i = (n4 << 2) + tid; i < n_el; i += POST_THREADS) tile[i] = __ldcs(src + i);
      } else {
        for (int i = tid; i < n_el; i += POST_THREADS) tile[i] = __ldcs(src + i);
      }
      __syncthreads();
    }

    if (tid < n_tile) {
      // at a detection threshold (conf >= 0.05) almost every anchor fails the objectness early-out: reading 4 of its
      // 4*(5+C) bytes from global memory beats staging the whole tile
      const float* row = (p.direct ? src : tile) + tid * D;
      const float so = sigmoid_exact(row[4]);
      if (so > p.conf) {                       // score <= sigmoid(obj): cheap early out
        float score = so;
        int cls = 0;
        if (C > 1) {
          float mx = row[5];
          int mi = 0;
          for (int c = 1; c < C; ++c) {
            const float v = row[5 + c];
            if (v > mx) { mx = v; mi = c; }
          }
          const float sc = sigmoid_exact(mx);
          score = __fmul_rn(so, sc);
          cls = mi;
          if (score > p.conf) {
            // torch takes max over the sigmoid VALUES (first max wins): a smaller logit earlier in the row
            // can round to the same sigmoid -> it is the reference's class.
            for (int c = 0; c < mi; ++c) {
              const float v = row[5 + c];
              if (v > mx - 1.0f || mx > 15.f) {
                if (sigmoid_exact(v) == sc) { cls = c; break; }
              }
            }
          }
        }
        if (score > p.conf) {
          const int a_lvl = a0 + tid;                      // a*Sh*Sw + y*Sw + x
          const int cell = a_lvl % (p.Sh[l] * p.Sw[l]);
          const int gy = cell / p.Sw[l], gx = cell - gy * p.Sw[l];
          const float stride = (float)((double)p.img_size / (double)p.Sh[l]);
          const float4 box = decode_box(row[0], row[1], row[2], row[3], gx, gy, stride, (float)(p.img_size - 1));
          const int n = p.lvl_off[l] + a_lvl;
          const int slot = atomicAdd(&p.count[b], 1);
          p.keys[(size_t)b * p.N + slot] = make_key(cls, score, n);
          p.cbox[(size_t)b * p.N + n] = box;
        }
      }
    }
  }

  // ---------------- ticket: is this the last tile of image b?
  // the CTA barrier orders every thread's candidate writes before thread 0, whose (cumulative) fence orders them before its ticket
  __syncthreads();
  if (tid == 0) {
    fence_acq_rel_gpu();
    const int ticket = atomicAdd(&p.done[b], 1);
    s_last = (ticket == tiles_per_image - 1);
    if (s_last) {
      fence_acq_rel_gpu();
      s_M = atomicAdd(&p.count[b], 0);
    }
  }
  __syncthreads();
  if (!s_last) return;
  fence_acq_rel_gpu();

  // ---------------- phase 2: sort + per-class NMS + compaction, one CTA for the whole image
  const int M = s_M;
  unsigned long long* gkeys = p.keys + (size_t)b * p.N;
  const float4* cbox = p.cbox + (size_t)b * p.N;
  unsigned long long* keys;
  unsigned char* flags;        // 0 = alive/undecided, 1 = suppressed, 2 = kept
  float4* sbox = nullptr;      // boxes of the sorted candidates, when they fit next to keys and flags (the usual detection case)
  if (M <= p.smem_keys) {
    keys = reinterpret_cast<unsigned long long*>(smem_raw);
    flags = smem_raw + (size_t)p.smem_keys * 8;
    int P = 1;
    while (P < M) P <<= 1;
    const size_t boff = (size_t)(P < 2 ? 2 : P) * 8;         // 16-byte aligned, past the P sorted keys
    if (boff + (size_t)M * 16 <= (size_t)p.smem_keys * 8) sbox = reinterpret_cast<float4*>(smem_raw + boff);
    for (int i = tid; i < P; i += POST_THREADS) keys[i] = i < M ? __ldcg(gkeys + i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < P; i += POST_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long x = keys[i], y = keys[ixj];
            const bool asc = (i & k) == 0;
            if ((x > y) == asc) { keys[i] = y; keys[ixj] = x; }
          }
        }
        __syncthreads();
      }
    }
  } else {
    // rank sort through global memory (keys are unique: the anchor index is part of the key)
    keys = p.keys2 + (size_t)b * p.N;
    flags = p.gflags + (size_t)b * p.N;
    for (int i = tid; i < M; i += POST_THREADS) {
      const unsigned long long ki = __ldcg(gkeys + i);
      int r = 0;
      for (int j = 0; j < M; ++j) r += (__ldcg(gkeys + j) < ki);
      keys[r] = ki;
    }
    __syncthreads();
  }
  for (int i = tid; i < M; i += POST_THREADS) flags[i] = 0;
  if (sbox)                    // one parallel gather of the candidates' boxes instead of a global gather per class segment
    for (int i = tid; i < M; i += POST_THREADS) sbox[i] = __ldcg(cbox + key_anchor(keys[i]));
  __syncthreads();

  // ---- greedy NMS, one warp per class segment.  Every warp walks ALL 32-candidate chunks and takes the segments whose ordinal is
  // congruent to its index: at a detection threshold there are ~100 candidates in ~60 classes, and handing out segments by the
  // chunk their first candidate sits in left those ~60 segments to 3 of the 8 warps (21 of the kernel's 49 us, in-kernel
  // %globaltimer stamps of scripts/timeline_post.py).
  const double thr = p.iou;
  int seg_ord = 0;                                           // segments that start before the current chunk
  for (int base = 0; base < M; base += 32) {
    const int pos = base + lane;
    const bool is_start = pos < M && (pos == 0 || key_cls(keys[pos]) != key_cls(keys[pos - 1]));
    const unsigned all_starts = __ballot_sync(0xffffffffu, is_start);
    unsigned starts = all_starts;
    while (starts) {
      const int bit = __ffs(starts) - 1;
      starts &= starts - 1;
      if (((seg_ord + __popc(all_starts & ((1u << bit) - 1u))) & (POST_THREADS / 32 - 1)) != warp) continue;
      const int s = base + bit;
      const int c = key_cls(keys[s]);
      if (s + 1 >= M || key_cls(keys[s + 1]) != c) {         // a class with ONE candidate: kept, nothing to suppress
        if (lane == 0) flags[s] = 2;
        continue;
      }
      // segment end: first position with a different class (binary search, uniform across the warp)
      int lo = s + 1, hi = M;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key_cls(keys[mid]) == c) lo = mid + 1; else hi = mid;
      }
      const int e = lo, n = e - s;
      int kept = 0;
      if (n <= 32 * REG_SEG) {
        float4 bx[REG_SEG];
        float ar[REG_SEG];
        unsigned dead = 0;         // bit r: my r-th box is suppressed / beyond the end
#pragma unroll
        for (int r = 0; r < REG_SEG; ++r) {
          const int q = s + lane + 32 * r;
          if (q < e) {
            bx[r] = sbox ? sbox[q] : __ldcg(cbox + key_anchor(keys[q]));
            ar[r] = __fmul_rn(__fsub_rn(bx[r].z, bx[r].x), __fsub_rn(bx[r].w, bx[r].y));
          } else {
            bx[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            ar[r] = 0.f;
            dead |= 1u << r;
          }
        }
        unsigned keepbits = 0;
#pragma unroll
        for (int r = 0; r < REG_SEG; ++r) {
          if (32 * r >= n) break;
          for (int o = 0; o < 32 && 32 * r + o < n; ++o) {
            const unsigned d = __shfl_sync(0xffffffffu, dead, o);
            if ((d >> r) & 1u) continue;
            const bool over = p.max_det > 0 && kept >= p.max_det;
            if (over) {              // keep[:max_det]: everything after the max_det-th survivor is dropped
              if (lane == o) dead |= 1u << r;
              continue;
            }
            ++kept;
            if (lane == o) keepbits |= 1u << r;
            float4 bi;
            bi.x = __shfl_sync(0xffffffffu, bx[r].x, o);
            bi.y = __shfl_sync(0xffffffffu, bx[r].y, o);
            bi.z = __shfl_sync(0xffffffffu, bx[r].z, o);
            bi.w = __shfl_sync(0xffffffffu, bx[r].w, o);
            const float ai = __shfl_sync(0xffffffffu, ar[r], o);
#pragma unroll
            for (int r2 = 0; r2 < REG_SEG; ++r2) {
              const int rel = lane + 32 * r2;
              if (rel > 32 * r + o && !((dead >> r2) & 1u)) {
                if (iou_gt(bi, ai, bx[r2], thr)) dead |= 1u << r2;
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < REG_SEG; ++r) {
          const int q = s + lane + 32 * r;
          if (q < e) flags[q] = ((keepbits >> r) & 1u) ? 2 : 1;
        }
      } else {
        // memory path: flags live in shared/global memory, boxes re-read through L1/L2
        for (int i = s; i < e; ++i) {
          __syncwarp();
          if (flags[i]) continue;
          if (p.max_det > 0 && kept >= p.max_det) {
            for (int j = i + lane; j < e; j += 32) if (!flags[j]) flags[j] = 1;
            break;
          }
          ++kept;
          const float4 bi = sbox ? sbox[i] : __ldcg(cbox + key_anchor(keys[i]));
          const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
          if (lane == 0) flags[i] = 2;
          for (int j = i + 1 + lane; j < e; j += 32) {
            if (!flags[j]) {
              const float4 bj = sbox ? sbox[j] : __ldcg(cbox + key_anchor(keys[j]));
              if (iou_gt(bi, ai, bj, thr)) flags[j] = 1;
            }
          }
        }
        __syncwarp();
      }
    }
    seg_ord += __popc(all_starts);
  }
  __syncthreads();

  // ---- compaction in sorted order (class asc, score desc, anchor asc)
  int running = 0;
  for (int base = 0; base < M; base += POST_THREADS) {
    const int pos = base + tid;
    const bool k = pos < M && flags[pos] == 2;
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) s_scan[warp] = __popc(bal);
    __syncthreads();
    if (tid == 0) {
      int acc = 0;
      for (int w = 0; w < POST_THREADS / 32; ++w) { const int v = s_scan[w]; s_scan[w] = acc; acc += v; }
      s_total = acc;
    }
    __syncthreads();
    if (k) {
      const int o = running + s_scan[warp] + __popc(bal & ((1u << lane) - 1u));
      if (o < p.cap) {
        const unsigned long long key = keys[pos];
        const int n = key_anchor(key);
        const size_t q = (size_t)b * p.cap + o;
        const float4 bx = sbox ? sbox[pos] : __ldcg(cbox + n);
        if (p.boxes) reinterpret_cast<float4*>(p.boxes)[q] = bx;
        if (p.scores) p.scores[q] = key_score(key);
        if (p.classes) p.classes[q] = key_cls(key);
        if (p.anchor_idx) p.anchor_idx[q] = n;
        if (p.packed) {
          float* r = p.packed + ((size_t)b * (p.cap + 1) + 1 + o) * 6;      // 24-byte rows: 8-byte aligned
          reinterpret_cast<float2*>(r)[0] = make_float2(bx.x, bx.y);
          reinterpret_cast<float2*>(r)[1] = make_float2(bx.z, bx.w);
          reinterpret_cast<float2*>(r)[2] = make_float2(key_score(key), (float)key_cls(key));
        }
      }
    }
    running += s_total;
    __syncthreads();
  }
  if (tid == 0) {
    if (p.counts) p.counts[b] = running <= p.cap ? running : (p.cap | (1 << 30));
    if (p.packed) {
      float* h = p.packed + (size_t)b * (p.cap + 1) * 6;
      h[0] = (float)min(running, p.cap); h[1] = running > p.cap ? 1.f : 0.f; h[2] = (float)running; h[3] = 0.f; h[4] = 0.f; h[5] = 0.f;
    }
    // self-clean the tickets so the scratch can be reused by the next launch without a memset
    p.count[b] = 0;
    p.done[b] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// decode only (utils_ms.py:25-123): box [B,N,4], obj [B,N,1], cls [B,N,C]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_kernel(PostParams p, float* box, float* obj, float* cls) {
  const long long total = (long long)p.B * p.N * p.D;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int d = (int)(idx % p.D);
  const long long bn = idx / p.D;
  const int n = (int)(bn % p.N);
  const int b = (int)(bn / p.N);
  int l = 0;
  while (l + 1 < p.n_levels && n >= p.lvl_off[l + 1]) ++l;
  const int a_lvl = n - p.lvl_off[l];
  const float* row = p.lvl[l] + ((size_t)b * p.n_lvl[l] + a_lvl) * p.D;
  if (d >= 5) {
    cls[((size_t)b * p.N + n) * p.C + (d - 5)] = row[d];
  } else if (d == 4) {
    obj[(size_t)b * p.N + n] = row[4];
  } else {
    const int cell = a_lvl % (p.Sh[l] * p.Sw[l]);
    const int gy = cell / p.Sw[l], gx = cell - gy * p.Sw[l];
    const float stride = (float)((double)p.img_size / (double)p.Sh[l]);
    const float4 bx = decode_box(row[0], row[1], row[2], row[3], gx, gy, stride, (float)(p.img_size - 1));
    box[((size_t)b * p.N + n) * 4 + d] = d == 0 ? bx.x : d == 1 ? bx.y : d == 2 ? bx.z : bx.w;
  }
}

static int fill_levels(PostParams& p, const float* const* level_logits, const int32_t* level_dims, int n_levels,
                       int B, int D, int img_size) {
  YL_REQUIRE(n_levels >= 1 && n_levels <= POST_MAX_LEVELS, "1..8 levels");
  YL_REQUIRE(D >= 5, "D = 5 + C with C >= 0");
  YL_REQUIRE(B >= 1, "B >= 1");
  p.n_levels = n_levels;
  p.B = B; p.D = D; p.C = D - 5; p.img_size = img_size;
  long long off = 0;
  int toff = 0;
  for (int l = 0; l < n_levels; ++l) {
    YL_REQUIRE(level_logits[l] != nullptr, "null level pointer");
    p.lvl[l] = level_logits[l];
    p.A[l] = level_dims[l * 3]; p.Sh[l] = level_dims[l * 3 + 1]; p.Sw[l] = level_dims[l * 3 + 2];
    YL_REQUIRE(p.A[l] >= 1 && p.Sh[l] >= 1 && p.Sw[l] >= 1, "level dims must be positive");
    p.n_lvl[l] = p.A[l] * p.Sh[l] * p.Sw[l];
    p.lvl_off[l] = (int)off;
    p.tile_off[l] = toff;
    off += p.n_lvl[l];
    toff += (p.n_lvl[l] + POST_TILE - 1) / POST_TILE;
  }
  p.tile_off[n_levels] = toff;
  YL_REQUIRE(off < (1ll << 20), "at most 2^20 anchors per image");
  YL_REQUIRE(p.C <= 4096, "at most 4096 classes");
  p.N = (int)off;
  return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace yl

extern "C" size_t yl_postprocess_scratch_bytes(int32_t B, int64_t N) {
  if (B < 1 || N < 1) return 0;
  const size_t bn = (size_t)B * (size_t)N;
  return yl::align_up(2 * (size_t)B * sizeof(int), 256) + yl::align_up(bn * 8, 256) * 2 +
         yl::align_up(bn * 16, 256) + yl::align_up(bn, 256);
}

namespace yl {

// the device that owns `ptr` becomes current for the lifetime of the guard (postprocess / decode / preprocess take raw pointers
// and no engine, so the device comes from the memory itself); the caller's device is restored afterwards
struct PtrDeviceGuard {
  int prev = -1;
  bool changed = false;
  int enter(const void* ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return 0;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (prev != at.device) { if (cudaSetDevice(at.device) != cudaSuccess) { cudaGetLastError(); return -2; } changed = true; }
    return 0;
  }
  ~PtrDeviceGuard() { if (changed) cudaSetDevice(prev); }
};

int post_run(const float* const* level_logits, const int32_t* level_dims, int n_levels, int B, int D, int img_size, float conf, double iou,
             int max_det_per_class, int cap, float* boxes, float* scores, int64_t* classes, int64_t* anchor_idx, int32_t* counts,
             float* packed, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  PostParams p{};
  if (int rc = fill_levels(p, level_logits, level_dims, n_levels, B, D, img_size)) return rc;
  YL_REQUIRE(scratch && (packed || (boxes && scores && classes && anchor_idx && counts)), "null output/scratch pointer");
  YL_REQUIRE(cap >= 1, "cap >= 1");
  YL_REQUIRE(scratch_bytes >= yl_postprocess_scratch_bytes(B, p.N), "scratch too small");
  YL_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "scratch must be 256-byte aligned");
  YL_REQUIRE(!packed || (reinterpret_cast<uintptr_t>(packed) & 7) == 0, "packed output must be 8-byte aligned");
  p.conf = conf; p.iou = iou; p.max_det = max_det_per_class; p.cap = cap;
  p.boxes = boxes; p.scores = scores; p.classes = reinterpret_cast<long long*>(classes);
  p.anchor_idx = reinterpret_cast<long long*>(anchor_idx); p.counts = counts; p.packed = packed;
  unsigned char* s = reinterpret_cast<unsigned char*>(scratch);
  const size_t bn = (size_t)B * p.N;
  p.count = reinterpret_cast<int*>(s); p.done = p.count + B;
  s += align_up(2 * (size_t)B * sizeof(int), 256);
  p.keys = reinterpret_cast<unsigned long long*>(s); s += align_up(bn * 8, 256);
  p.keys2 = reinterpret_cast<unsigned long long*>(s); s += align_up(bn * 8, 256);
  p.cbox = reinterpret_cast<float4*>(s); s += align_up(bn * 16, 256);
  p.gflags = s;
  YL_CHECK_CUDA(cudaMemsetAsync(p.count, 0, 2 * (size_t)B * sizeof(int), st));
  p.direct = conf >= 0.05f ? 1 : 0;
  p.smem_keys = p.direct ? POST_SMEM_KEYS / 2 : POST_SMEM_KEYS;
  const size_t tile_bytes = p.direct ? 0 : (size_t)POST_TILE * D * sizeof(float);
  const size_t sort_bytes = (size_t)p.smem_keys * 9;
  const size_t smem = (tile_bytes > sort_bytes ? tile_bytes : sort_bytes) + 16;
  YL_REQUIRE(smem <= 200 * 1024, "5+C too large for the shared-memory tile (C <= 195)");
  {   // the dynamic shared memory limit is a per-device function attribute
    static size_t smem_set[64];
    static std::mutex mtx;
    int dev = 0;
    YL_CHECK_CUDA(cudaGetDevice(&dev));
    YL_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    std::lock_guard<std::mutex> lk(mtx);
    if (smem > smem_set[dev]) {
      YL_CHECK_CUDA(cudaFuncSetAttribute(post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set[dev] = smem;
    }
  }
  const unsigned grid = (unsigned)B * (unsigned)p.tile_off[n_levels];
  ++g_post_launches;
  post_kernel<<<grid, POST_THREADS, smem, st>>>(p);
  YL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace yl

extern "C" int yl_postprocess_ex(const float* const* level_logits, const int32_t* level_dims, int32_t n_levels, int32_t B, int32_t D,
                                 int32_t img_size, float conf, double iou, int32_t max_det_per_class, int32_t cap, float* boxes,
                                 float* scores, int64_t* classes, int64_t* anchor_idx, int32_t* counts, float* packed, void* scratch,
                                 size_t scratch_bytes, void* stream) {
  using namespace yl;
  YL_REQUIRE(level_logits && level_dims && scratch, "null argument");
  PtrDeviceGuard dg;
  YL_REQUIRE(dg.enter(scratch) == 0, "cannot select the device that owns the scratch buffer");
  return post_run(level_logits, level_dims, n_levels, B, D, img_size, conf, iou, max_det_per_class, cap, boxes, scores, classes,
                  anchor_idx, counts, packed, scratch, scratch_bytes, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int yl_postprocess(const float* const* level_logits, const int32_t* level_dims, int32_t n_levels,
                              int32_t B, int32_t D, int32_t img_size, float conf, double iou,
                              int32_t max_det_per_class, int32_t cap, float* boxes, float* scores,
                              int64_t* classes, int64_t* anchor_idx, int32_t* counts, void* scratch,
                              size_t scratch_bytes, void* stream) {
  using namespace yl;
  YL_REQUIRE(boxes && scores && classes && anchor_idx && counts && scratch, "null output/scratch pointer");
  return yl_postprocess_ex(level_logits, level_dims, n_levels, B, D, img_size, conf, iou, max_det_per_class, cap, boxes, scores,
                           classes, anchor_idx, counts, nullptr, scratch, scratch_bytes, stream);
}

extern "C" int yl_decode(const float* const* level_logits, const int32_t* level_dims, int32_t n_levels, int32_t B,
                         int32_t D, int32_t img_size, float* box, float* obj, float* cls, void* stream) {
  using namespace yl;
  PostParams p{};
  if (int rc = fill_levels(p, level_logits, level_dims, n_levels, B, D, img_size)) return rc;
  YL_REQUIRE(box && obj && (cls || D == 5), "null output pointer");
  PtrDeviceGuard dg;
  YL_REQUIRE(dg.enter(box) == 0, "cannot select the device that owns the output");
  const long long total = (long long)B * p.N * D;
  decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, box, obj, cls);
  YL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
