/*
 * yololite_b200 -- C ABI of the B200 (sm_100a) engine for YoloLite's detection forward pass + postprocess.
 *
 * The reference (Lillthorin/YoloLite-Official-Repo) has no FFI: its boundary is the Python duck type
 * `model(x) -> list[Tensor]` plus a handful of free functions.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference repo root).  Conventions:
 *
 *   - plain pointers and sizes only; every tensor is allocated by the caller (so on the Python side the
 *     outputs are ordinary torch.Tensors) and passed as a raw DEVICE pointer unless the name says host;
 *   - the engine owns packed weights and scratch, nothing else;
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*), no host sync inside
 *     yl_forward / yl_postprocess / yl_preprocess.  After yl_engine_plan for a shape, yl_forward / yl_engine_detect allocate
 *     nothing and are graph-capturable (the first call for a NEW shape sizes the arena with cudaMalloc: plan before capturing);
 *   - the caller's current CUDA device is never changed: entry points select the engine's (or the buffers') device and restore;
 *   - return value 0 = ok, negative = error; the message is in yl_last_error() (thread local);
 *   - there is NO CPU fallback: every call fails loudly without a CUDA device.
 */
#ifndef YOLOLITE_B200_H
#define YOLOLITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YL_ABI_VERSION 4

typedef struct yl_engine yl_engine;

/* ---- layer program ------------------------------------------------------------------------------
 * The host side (Python packer) lowers a reference checkpoint {"state_dict","meta"}
 * (tools/train.py:62-75, tools/infer.py:80-102) into a flat list of fused ops over numbered NHWC fp32
 * activation buffers plus one fp32 weight blob with BatchNorm folded into the preceding conv
 * (scripts/model/model_v2.py:15-53,250-377 and the timm backbone it calls at :266-272).           */
enum yl_op_kind {
  YL_OP_STEM = 0,  /* dense 3x3 conv on the NCHW network input (Cin=3) -> NHWC, +bias, act             */
  YL_OP_CONV = 1,  /* dense KxK conv as implicit GEMM, NHWC -> NHWC; K=1 is the pointwise conv; epilogue:
                      +bias, +residual buffer, +nearest-upsampled coarser buffer, act, head layout       */
  YL_OP_DW = 2,    /* depthwise KxK conv (K = 3 or 5), NHWC, +bias, act                                   */
  YL_OP_DWPW = 3,  /* fused depthwise k2 x k2 (3 or 5, stride stride2 = 1 or 2, +bias b2, act2) -> pointwise + bias + residual + act;
                      the depthwise result never leaves shared memory.  Covers DWConvBlock (model_v2.py:23-39: 3x3, no
                      bias, no act2) and the dw_start -> pw_exp / dw_mid -> pw_proj pairs of timm's
                      UniversalInvertedResidual (backbone called at model_v2.py:266-272)                   */
  YL_OP_STEM2 = 4  /* fused conv_stem (3x3 s2, Cin=3, NCHW input, 32 ch, +bias+ReLU) -> dense 3x3 s2 conv + bias + act
                      (timm blocks.0.0): the stem activation (13 MB/image at 640 px) never leaves shared memory.
                      cin = 3, cout = channels of the second conv, k/stride/act = the second conv's, k2 = stem
                      channels (32), w2_off = [27][32] stem weights followed by 32 stem biases, wt_off required.
                      Shapes the fused kernel cannot take (W % 4 != 0, cout > 32) run unfused: SIMT stem -> conv -> pointwise  */
};
enum yl_act { YL_ACT_NONE = 0, YL_ACT_RELU = 1, YL_ACT_SILU = 2 };

#define YL_SRC_INPUT (-1)          /* op.src: the network input x                                      */
#define YL_SRC_FEATURE(i) (-(2 + (i))) /* op.src: externally supplied backbone feature i (NHWC fp32), yl_forward_features */
#define YL_FEATURE_INDEX(src) (-(src) - 2)
#define YL_DST_LEVEL(l) (-(1 + (l))) /* op.dst: write output level l ([B,A,S,S,5+C], model_v2.py:340-350) */

typedef struct yl_op {
  int32_t kind;      /* yl_op_kind */
  int32_t src;       /* activation buffer id, or YL_SRC_INPUT */
  int32_t dst;       /* activation buffer id, or YL_DST_LEVEL(l) */
  int32_t res;       /* buffer id added to the output (same shape) before act, or -1 */
  int32_t up;        /* buffer id of a coarser map, nearest-resized to the output size and added, or -1 */
  int32_t cin, cout;
  int32_t k, stride; /* padding is k/2 */
  int32_t act;       /* yl_act */
  int32_t anchors;   /* >0 only for head output convs: cout = anchors*(5+C), stored [B,A,H,W,5+C] */
  int32_t k2;        /* YL_OP_DWPW: depthwise kernel size (3 or 5); YL_OP_STEM2: stem channels; otherwise 0 */
  int64_t w_off;     /* float offset of the GEMM/stencil weights in the blob */
  int64_t b_off;     /* float offset of the bias (cout floats), or -1 */
  int64_t w2_off;    /* YL_OP_DWPW: float offset of depthwise weights [k2*k2][cin]; otherwise -1 */
  int64_t wt_off;    /* float offset of the tcgen05 weight image [ceil(K/32)][3 splits][ceil16(cout)][32 k] in bf16 (two per float
                        slot): w = w1 + w2 + w3 pre-split and pre-swizzled (64 B rows, SWIZZLE_64B, K-major), or -1 */
  int64_t w3_off;    /* YL_OP_STEM2: float offset of the bf16-triple weight image of the fused stem kernel
                        ([9 taps][3 splits][ceil16(cout)][32] conv2 | [3 splits][32][32] stem incl. bias row k = 27, each
                        row 64 B, SWIZZLE_64B K-major, two bf16 per float slot; followed by a second [3 splits][32][32] stem image
                        for uint8 input with 1/(255 std) and the -mean/std terms folded in: rows k = 27..30 = constant, top-border,
                        left-border and top-left-corner terms), or -1 (older tf32 kernel) */
  int64_t b2_off;    /* YL_OP_DWPW: float offset of the depthwise bias (cin floats, folded BN), or -1.
                        YL_OP_STEM2: float offset of a fused pointwise conv applied after the second conv (timm blocks.0.1):
                        [cout][cout] weights (k-major, BN folded) followed by cout biases, cout = 16 only; or -1 */
  int32_t act2;      /* YL_OP_DWPW: yl_act applied to the depthwise result before the pointwise conv;
                        YL_OP_STEM2: yl_act of the fused pointwise conv */
  int32_t stride2;   /* YL_OP_DWPW: stride of the depthwise stage (0 or 1 = 1, 2); the output size follows it */
  int32_t wt_layout; /* K axis of the wt_off image: 0 = k = (ky*k + kx)*cin + ci packed densely (slabs of 32 may straddle taps);
                        1 = every tap padded to a multiple of 32 channels, k' = (ky*k + kx)*ceil32(cin) + ci (dense k x k convs with
                        stride 1 and a long cin: each K-slab is then ONE shifted TMA box of the NHWC input, no gather) */
  int32_t reserved0;
} yl_op;

/* Build an engine on `device` from a layer program and a HOST weight blob.
 * Replaces: tools/infer.py:34-102 build_model_from_meta + load_state_dict + model.to(device).eval().  */
int yl_engine_create(const yl_op* ops, int32_t n_ops, const float* blob_host, size_t blob_floats,
                     int32_t n_buffers, int32_t n_levels, int32_t device, yl_engine** out);
int yl_engine_destroy(yl_engine* e);

/* Engine options.  "tensor_cores": 1 (default) = run eligible convs on the tcgen05 bf16-triple kernel, 0 = fp32 SIMT kernels only.
 * "pdl": 1 (default) = launch the tcgen05 kernels with programmatic dependent launch (prologue overlaps the previous kernel's
 * tail).  "graph": 1 = capture each distinct call (same pointers, shape and thresholds) once into a CUDA graph and replay it
 * (default 0; the first call with new pointers runs eagerly, the second captures).  Unknown keys return an error.          */
int yl_engine_set_option(yl_engine* e, const char* key, int32_t value);

/* Shapes of the output levels for an input of B x 3 x H x W: shapes[l*4 + {0,1,2,3}] = A, S_h, S_w, 5+C.
 * Also (re)sizes the activation arena.  Replaces nothing 1:1; the reference gets shapes from the tensors
 * model.forward returns (model_v2.py:352-377).                                                          */
int yl_engine_plan(yl_engine* e, int32_t B, int32_t H, int32_t W, int32_t* shapes /* n_levels*4 */);

/* model.forward(x) (scripts/model/model_v2.py:352-377; callers tools/infer.py:456,
 * scripts/helpers/evaluate.py:273,290,423).  x: [B,3,H,W] fp32 NCHW, normalised.  level_out[l]: caller-
 * allocated [B,A,S,S,5+C] fp32 contiguous, channel order (tx,ty,tw,th,obj,cls0..).  x is not modified.  */
int yl_forward(yl_engine* e, const float* x, int32_t B, int32_t H, int32_t W, float* const* level_out,
               void* stream);

/* model(x) on an image batch that needs no letterbox resize: images_bgr is [B,H,W,3] uint8 BGR on the device (what
 * cv2.imread gives the reference, tools/infer.py:436), H = W = the network input size.  Equivalent to yl_preprocess_batch
 * (BGR->RGB, /255, (x-mean)/std, CHW; tools/infer.py:442-453) followed by yl_forward, but the normalisation is folded into the
 * stem weights (it is affine in the integer pixel value) and the fp32 image never exists in HBM.  Needs the fused stem op
 * (YL_OP_STEM2 with w3_off), even H and W % 16 == 0; returns -1 otherwise (callers fall back to yl_preprocess_batch + yl_forward). */
int yl_forward_u8(yl_engine* e, const uint8_t* images_bgr, int32_t B, int32_t H, int32_t W, float* const* level_out,
                  void* stream);

/* FPN + heads only (scripts/model/model_v2.py:124-133,201-224 / :289-294,359-377): the layer program reads the backbone's feature
 * maps instead of an image (ops with src = YL_SRC_FEATURE(i); built by the packer's `lower(..., from_features=True)`).
 * feats[i]: NHWC fp32 [B,H_i,W_i,C_i] on the device (a torch channels_last tensor is exactly this), finest level first
 * ([c2,] c3, c4, c5 -- what timm's features_only backbone returns at model_v2.py:195,353); feat_dims[i*3 + {0,1,2}] = H_i, W_i, C_i. */
int yl_engine_plan_features(yl_engine* e, int32_t B, const int32_t* feat_dims, int32_t n_feats, int32_t* shapes /* n_levels*4 */);
int yl_forward_features(yl_engine* e, const float* const* feats, const int32_t* feat_dims, int32_t n_feats, int32_t B,
                        float* const* level_out, void* stream);

/* model(x) + postprocess in ONE call (tools/infer.py:456-493 for a batch): exactly one of x (fp32 NCHW, normalised) and
 * images_bgr (uint8 HWC BGR, see yl_forward_u8) is non-null.  The logits stay in engine-owned level buffers (yl_engine_levels);
 * outputs as yl_postprocess_ex.  With the "graph" option the whole call is one CUDA graph launch.                          */
int yl_engine_detect(yl_engine* e, const float* x, const uint8_t* images_bgr, int32_t B, int32_t H, int32_t W, int32_t img_size,
                     float conf, double iou, int32_t max_det_per_class, int32_t cap, float* boxes, float* scores, int64_t* classes,
                     int64_t* anchor_idx, int32_t* counts, float* packed, void* stream);
/* Device pointers / shapes (n_levels*4: A, S_h, S_w, 5+C) of the engine-owned level buffers of the current plan.           */
int yl_engine_levels(yl_engine* e, float** level_ptrs /* n_levels */, int32_t* shapes /* n_levels*4, may be NULL */);

/* Same as yl_forward but brackets every op with CUDA events on `stream` and returns the device time of each
 * op in milliseconds (op_ms[n_ops]); synchronises the stream.  Used by bench.py for the per-kernel roofline. */
int yl_forward_profile(yl_engine* e, const float* x, int32_t B, int32_t H, int32_t W, float* const* level_out,
                       void* stream, float* op_ms /* host, n_ops */, int32_t n_ops);

/* Run ONE op of a layer program on caller-provided NHWC device tensors (unit tests / micro-benchmarks of a
 * single kernel).  blob_dev is the DEVICE copy of the weight blob the op's offsets refer to.                */
int yl_run_op(const yl_op* op, const float* blob_dev, const float* in, const float* res, const float* up, float* out,
              int32_t B, int32_t Hin, int32_t Win, int32_t Hu, int32_t Wu, int32_t use_tensor_cores, void* stream);

/* Debug/parity tap: copy activation buffer `buf` ([B,H,W,C] NHWC fp32) of the last yl_forward into `dst`
 * (device).  dims receives H, W, C.  dst may be NULL to query dims only.                                */
int yl_engine_read_buffer(yl_engine* e, int32_t buf, float* dst, int32_t* dims /*3*/, void* stream);

/* ---- postprocess --------------------------------------------------------------------------------
 * One fused kernel per batch: sigmoid, anchor-free decode, score threshold, class-wise NMS.
 * Replaces: scripts/helpers/utils_ms.py:25-123 decode_preds_anchorfree (center_mode "v8", wh_mode
 * "softplus"), tools/infer.py:466-493 (score = sigmoid(obj)*max sigmoid(cls); C==1 -> obj only; strict
 * score > conf; per-class torchvision.ops.nms, keep[:max_det] per class via tools/infer.py:134-152), and
 * with max_det_per_class = 0 (unlimited), conf 0.001, iou 0.65 the evaluation variant
 * scripts/helpers/helpers.py:86-153.
 *
 * level_logits[l]: [B, A_l, Sh_l, Sw_l, D] fp32 (D = 5 + C).  level_dims[l*3 + {0,1,2}] = A_l, Sh_l, Sw_l.
 * Outputs (capacity `cap` detections per image, caller-allocated, device):
 *   boxes [B,cap,4] f32 xyxy px in the img_size square, scores [B,cap] f32, classes [B,cap] i64,
 *   anchor_idx [B,cap] i64 (flat index n = level offset + a*S*S + y*S + x, the reference's N axis),
 *   counts [B] i32 (= min(K, cap); bit 30 set in counts[b] if K > cap).
 * Order within an image: class ascending, then score descending, ties by anchor index -- exactly the
 * concatenation order of tools/infer.py:477-488.
 * `scratch` is device memory of yl_postprocess_scratch_bytes(B, N) bytes, reusable across calls.       */
size_t yl_postprocess_scratch_bytes(int32_t B, int64_t n_anchors_total);
int yl_postprocess(const float* const* level_logits, const int32_t* level_dims, int32_t n_levels,
                   int32_t B, int32_t D, int32_t img_size, float conf, double iou,
                   int32_t max_det_per_class, int32_t cap,
                   float* boxes, float* scores, int64_t* classes, int64_t* anchor_idx, int32_t* counts,
                   void* scratch, size_t scratch_bytes, void* stream);

/* Same with optional outputs: any of boxes / scores / classes / anchor_idx / counts may be NULL when `packed` is given.
 * packed: [B][cap + 1][6] fp32, row 0 of an image = (count = min(K, cap), overflow flag, K, 0, 0, 0), rows 1..count =
 * (x1, y1, x2, y2, score, class) -- the fixed-capacity payload of the multi-GPU gather written by the
 * kernel itself, so that the gather is ONE collective with no packing pass.                                                */
int yl_postprocess_ex(const float* const* level_logits, const int32_t* level_dims, int32_t n_levels,
                      int32_t B, int32_t D, int32_t img_size, float conf, double iou,
                      int32_t max_det_per_class, int32_t cap,
                      float* boxes, float* scores, int64_t* classes, int64_t* anchor_idx, int32_t* counts, float* packed,
                      void* scratch, size_t scratch_bytes, void* stream);

/* Decode only (utils_ms.py:25-123): box [B,N,4], obj [B,N,1], cls [B,N,C] -- obj/cls are raw logits.   */
int yl_decode(const float* const* level_logits, const int32_t* level_dims, int32_t n_levels, int32_t B,
              int32_t D, int32_t img_size, float* box, float* obj, float* cls, void* stream);

/* ---- preprocess / back-map (tools/infer.py:121-131,432-453,507-516) --------------------------------
 * src: uint8 HWC BGR image on the device (h0 x w0 x 3, row pitch `pitch` bytes).  Writes one image of
 * the batch: dst [3,S,S] fp32 = letterbox(114 pad, bilinear as cv2.INTER_LINEAR) -> RGB -> /255 ->
 * (x-mean)/std -> CHW.  nh,nw,left,top as computed by the host (int(round()) sizing).                  */
int yl_preprocess(const uint8_t* src, int32_t h0, int32_t w0, int32_t pitch, float* dst, int32_t S,
                  int32_t nh, int32_t nw, int32_t left, int32_t top, void* stream);

/* Same for a contiguous batch of B equally sized images: src [B,h0,w0,3] uint8 (device), dst [B,3,S,S] fp32.  */
int yl_preprocess_batch(const uint8_t* src, int32_t B, int32_t h0, int32_t w0, float* dst, int32_t S, int32_t nh, int32_t nw,
                        int32_t left, int32_t top, void* stream);

/* Process-wide launch counters: key "tc_launches" (tcgen05 conv kernel), "simt_launches" (fp32 SIMT conv kernels),
 * "post_launches", "graph_launches" / "graph_captures" (CUDA graph replays / captures), "prepares" (launch records built).  Unknown key -> -1.  Lets tests assert WHICH kernel a call used.                       */
long long yl_stat(const char* key);

const char* yl_last_error(void);
int yl_abi_version(void);
int yl_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* YOLOLITE_B200_H */
