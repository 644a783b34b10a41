/* iou_b200.h -- C ABI of libiou_b200.so (sm_100a).
 *
 * Drop-in boundary for the IoU-aware RetinaNet inference hot path of
 * ShengkaiWu/IoU-aware-single-stage-object-detector (an mmdetection v0.6.0 fork).
 * Each entry point names the reference interface it replaces (file:line relative
 * to the reference tree).  Conventions:
 *   - every pointer is a DEVICE pointer supplied by the caller unless the
 *     parameter is documented as "host";
 *   - the library never allocates device memory: scratch comes in through
 *     `workspace` (size from the matching *_workspace_bytes call);
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the host;
 *   - return value 0 = ok, otherwise a negative code; text via iou_last_error();
 *   - inputs are borrowed and never written.
 */
#ifndef IOU_B200_H_
#define IOU_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IOU_MAX_LEVELS 8
#define IOU_MAX_ANCHORS 16
#define IOU_MAX_CANDIDATES 6144   /* per image (sum over levels of min(n_l, nms_pre)) */
#define IOU_MAX_NMS_BOXES 6144    /* iou_nms: boxes per call of the shared-memory path; iou_soft_nms: hard limit */

#define IOU_OK 0
#define IOU_ERR_INVALID (-1)
#define IOU_ERR_CUDA (-2)
#define IOU_ERR_UNSUPPORTED (-3)
#define IOU_ERR_WORKSPACE (-4)

/* Text of the last error raised on the calling thread ("" if none). */
const char* iou_last_error(void);
/* ABI version (bumped on any signature change). */
int iou_abi_version(void);
/* sizeof of the ABI structs as compiled (0: iou_postproc_cfg, 1: iou_conv_desc, 2: iou_conv_segment),
 * so that a foreign-language binding can verify its own struct layout. */
size_t iou_sizeof(int what);

/* ------------------------------------------------------------------ get_bboxes
 * Static description of IoUawareRetinaHead.get_bboxes / get_bboxes_single
 * (mmdet/models/anchor_heads/iou_aware_retina_head.py:390-564) +
 * multiclass_nms (mmdet/core/post_processing/bbox_nms.py:6-67).               */
typedef struct iou_postproc_cfg {
  int32_t num_levels;                      /* FPN levels, P3..P7 = 5                         */
  int32_t num_anchors;                     /* A = len(ratios)*len(scales) (anchor_head.py:79) */
  int32_t num_classes;                     /* C = cls_out_channels (sigmoid: num_classes-1)   */
  int32_t nms_pre;                         /* test_cfg.nms_pre (<=0: keep all)                */
  int32_t max_per_img;                     /* test_cfg.max_per_img                            */
  int32_t feat_h[IOU_MAX_LEVELS];
  int32_t feat_w[IOU_MAX_LEVELS];
  int32_t stride[IOU_MAX_LEVELS];          /* anchor_strides                                  */
  float base_anchors[IOU_MAX_LEVELS][IOU_MAX_ANCHORS][4]; /* AnchorGenerator.base_anchors    */
  float target_means[4];
  float target_stds[4];
  float alpha;                             /* score = cls^alpha * iou^(1-alpha); 0.5 at :510;
                                              1.0 = plain RetinaHead (iou maps may then be NULL) */
  float score_thr;                         /* test_cfg.score_thr (strict >)                   */
  float iou_thr;                           /* test_cfg.nms.iou_thr (strict >, nms_kernel.cu:60)*/
  float wh_ratio_clip;                     /* delta2bbox wh_ratio_clip, 16/1000               */
  int32_t decode_mode;                     /* IOU_DECODE_DELTA: anchors + delta2bbox (transforms.py:44-78);
                                              IOU_DECODE_DISTANCE: FCOS points (x*s + s/2, y*s + s/2) +
                                              distance2bbox (transforms.py:169-190, iou_aware_fcos_head.py:392-401),
                                              num_anchors must be 1 and reg holds (l, t, r, b) distances   */
} iou_postproc_cfg;
enum { IOU_DECODE_DELTA = 0, IOU_DECODE_DISTANCE = 1 };

/* Number of candidate rows per image that enter NMS: sum_l min(H_l*W_l*A, nms_pre). */
int iou_postproc_num_candidates(const iou_postproc_cfg* cfg);
size_t iou_postproc_workspace_bytes(const iou_postproc_cfg* cfg, int n_img);

/* Stage 1 -- replaces iou_aware_retina_head.py:499-554 for a whole batch.
 * cls/reg/iou: host arrays of num_levels device pointers; level l is laid out
 * NHWC, i.e. [n_img][H_l][W_l][A*C], [..][A*4], [..][A] fp32 (== the reference's
 * permute(1,2,0).reshape(-1,C) rows).  img_info: [n_img][8] fp32 =
 * (img_h, img_w, sf_x1, sf_y1, sf_x2, sf_y2, 0, 0); boxes are divided by sf if
 * rescale != 0.  Outputs: boxes [n_img][M][4]; scores_cm [n_img][C][M]
 * (class-major); cand_idx [n_img][M] int32 level-local anchor index.          */
int iou_decode_candidates(const iou_postproc_cfg* cfg, int n_img,
                          const float* const* cls, const float* const* reg,
                          const float* const* iou, const float* img_info, int rescale,
                          float* boxes, float* scores_cm, int32_t* cand_idx,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Stage 2 -- replaces multiclass_nms (bbox_nms.py:6-67) incl. the per-class
 * nms_cuda.nms calls (ops/nms/src/nms_kernel.cu:70-131) for a whole batch.
 * dets [n_img][max_per_img][5], labels [n_img][max_per_img] int64,
 * counts [n_img] int32 (rows beyond counts[i] are zero).                      */
int iou_batched_nms(const iou_postproc_cfg* cfg, int n_img, const float* boxes,
                    const float* scores_cm, float* dets, int64_t* labels, int32_t* counts,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Stage 1 + 2 == IoUawareRetinaHead.get_bboxes (iou_aware_retina_head.py:390-461). */
int iou_get_bboxes(const iou_postproc_cfg* cfg, int n_img,
                   const float* const* cls, const float* const* reg, const float* const* iou,
                   const float* img_info, int rescale,
                   float* dets, int64_t* labels, int32_t* counts,
                   void* workspace, size_t workspace_bytes, void* stream);

/* iou_get_bboxes with the per-anchor maximum class logit already reduced by the producer of the class maps
 * (iou_conv_desc.group_max_cols = num_classes on retina_cls): cls_max2 is a host array of num_levels device pointers,
 * level l = fp32 [n_img][H_l*W_l*A][2] whose two entries' max is the anchor's max class logit; a NULL entry (or a NULL
 * array) makes that level read its class map as iou_get_bboxes does.  Same results as iou_get_bboxes, bit for bit:
 * the top-k pre-selection (iou_aware_retina_head.py:510-519) only needs the max, the candidates' class rows are
 * still gathered from `cls`. */
int iou_get_bboxes_premax(const iou_postproc_cfg* cfg, int n_img,
                          const float* const* cls, const float* const* reg, const float* const* iou,
                          const float* const* cls_max2, const float* img_info, int rescale,
                          float* dets, int64_t* labels, int32_t* counts,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ plain NMS
 * Drop-in for mmdet.ops.nms.nms_cuda.nms (ops/nms/src/nms_cuda.cpp:8-13,
 * nms_kernel.cu:70-131): dets [n][5] fp32 (x1,y1,x2,y2,score); suppress at
 * IoU > thr; keep_idx receives the ORIGINAL indices of kept boxes in ascending
 * order (int64, capacity n), keep_count the number kept.  Any n, like the reference:
 * up to IOU_MAX_NMS_BOXES boxes run in one block's shared memory (workspace may be
 * NULL); above that, n x ceil(n/64) mask words live in `workspace`
 * (iou_nms_workspace_bytes(n) bytes: ~n*n/8), with the sort and the greedy scan that
 * the reference runs on the host (nms_kernel.cu:99-123) on the device. */
size_t iou_nms_workspace_bytes(int n);
int iou_nms(const float* dets, int n, float iou_thr, int64_t* keep_idx, int32_t* keep_count,
            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ soft NMS ("next" row)
 * Drop-in for mmdet.ops.nms.soft_nms_cpu.soft_nms_cpu (ops/nms/src/soft_nms_cpu.pyx:22-127; Python wrapper
 * ops/nms/nms_wrapper.py:52-78).  dets [n][5] fp32 DEVICE rows (x1,y1,x2,y2,score); method 1 = linear,
 * 2 = gaussian (the wrapper's method_codes), anything else the .pyx accepts = hard 0/1 weights (3 here).
 * out_dets [n][5] receives the surviving rows with their decayed scores in SELECTION order, out_inds [n]
 * (int64) their original row numbers, out_count the number of survivors.  n <= IOU_MAX_NMS_BOXES.
 * Results are bit-identical to the reference's CPU loop, including its tie order (see csrc/postproc.cu). */
int iou_soft_nms(const float* dets, int n, float iou_thr, int method, float sigma, float min_score,
                 float* out_dets, int64_t* out_inds, int32_t* out_count, void* stream);

/* multiclass_nms with nms_cfg = dict(type='soft_nms', iou_thr, method, sigma, min_score) for a whole batch
 * (core/post_processing/bbox_nms.py:29-67 calling nms_wrapper.soft_nms per class).  Same buffers as
 * iou_batched_nms; cfg->iou_thr is the soft-NMS iou_thr.  Rows of one class keep their selection order. */
int iou_batched_soft_nms(const iou_postproc_cfg* cfg, int n_img, const float* boxes, const float* scores_cm,
                         int method, float sigma, float min_score, float* dets, int64_t* labels,
                         int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ focal loss
 * Drop-in for sigmoid_focal_loss_cuda.forward / .backward
 * (ops/sigmoid_focal_loss/src/sigmoid_focal_loss.cpp:17-43,
 *  sigmoid_focal_loss_cuda.cu:24-105).  logits [n][c], targets [n] int64
 * (0 = background, class d <-> d+1).  The reference dispatches fp16 / fp32 / fp64
 * (AT_DISPATCH_FLOATING_TYPES_AND_HALF, .cu:128,167): `dtype` selects the element
 * type of logits / losses / d_losses / d_logits; the un-suffixed entry points are fp32. */
#define IOU_DTYPE_F32 0
#define IOU_DTYPE_F16 1
#define IOU_DTYPE_F64 2
int iou_sigmoid_focal_loss_forward(const float* logits, const int64_t* targets, int n, int c,
                                   float gamma, float alpha, float* losses, void* stream);
int iou_sigmoid_focal_loss_backward(const float* logits, const int64_t* targets,
                                    const float* d_losses, int n, int c, float gamma, float alpha,
                                    float* d_logits, void* stream);
int iou_sigmoid_focal_loss_forward_dtype(const void* logits, int dtype, const int64_t* targets, int n, int c,
                                         float gamma, float alpha, void* losses, void* stream);
int iou_sigmoid_focal_loss_backward_dtype(const void* logits, int dtype, const int64_t* targets,
                                          const void* d_losses, int n, int c, float gamma, float alpha,
                                          void* d_logits, void* stream);

/* ------------------------------------------------------------------ conv engine
 * Replaces the F.conv2d (+BatchNorm eval, +bias, +ReLU, +residual add) calls of
 * ResNet.forward (backbones/resnet.py:224-267,507-518), FPN.forward
 * (necks/fpn.py:97-136) and IoUawareRetinaHead.forward_single
 * (anchor_heads/iou_aware_retina_head.py:171-219) with one tcgen05 "tap GEMM".
 *
 * Activation layout ("padded rows"): a feature map of n images, H x W, Cch
 * channels is a bf16 matrix [rows][2*Cch]; row = (img*(H+2) + y+1)*(W+2) + x+1,
 * columns [0,Cch) hold hi = bf16(v), columns [Cch,2Cch) hold lo = bf16(v - hi);
 * the one-pixel border rows are zero.  Several maps (FPN levels) may be
 * concatenated as "segments", each starting at a multiple of 128 rows.
 * Weight layout: bf16 [taps*cout_pad][2*Cin] (hi | lo), tap-major, K contiguous.
 * passes == 2 (IOU_FMT_F16F8, declared with the layout kernels): the same row sizes hold
 * [Wh: Cin x fp16][per 8 input channels: Wl8 x 8 | W8 x 8], Wl8 = e4m3((w - Wh) * 2^11 * s_n),
 * W8 = e4m3(w * s_n), s_n a power of two per output channel; acc = Wh*hi + scale[n] * (Wl8*x8 + W8*l8). */
#define IOU_CONV_MAX_SEG 8
#define IOU_CONV_MAX_TAPS 9
#define IOU_CONV_MAX_SRC 4

typedef struct iou_conv_segment {
  int32_t row_start;     /* first row of the segment (multiple of 128)           */
  int32_t n_img, h, w;   /* unpadded output geometry of the segment               */
} iou_conv_segment;

enum { IOU_OUT_PADDED_BF16X2 = 0,  /* padded-rows bf16 hi|lo, same geometry           */
       IOU_OUT_DENSE_F32 = 1 };    /* fp32 [n][h][w][cout] per segment (head outputs)  */
enum { IOU_RES_NONE = 0, IOU_RES_SAME = 1, IOU_RES_UPSAMPLE2 = 2 };

typedef struct iou_conv_desc {
  int32_t cin, cout, cout_pad, block_n;     /* cin % 64 == 0; cout_pad % block_n == 0    */
  int32_t num_taps;
  int32_t tap_src[IOU_CONV_MAX_TAPS];       /* which source matrix a tap reads           */
  int32_t tap_dy[IOU_CONV_MAX_TAPS];        /* row offset = dy*(w+2) + dx per segment    */
  int32_t tap_dx[IOU_CONV_MAX_TAPS];
  int32_t num_src;
  const void* src[IOU_CONV_MAX_SRC];        /* bf16 [src_rows][2*cin]                    */
  int64_t src_rows;
  const void* weight;                       /* bf16 [num_taps*cout_pad][2*cin]           */
  const float* scale;                       /* [cout_pad] per-channel multiplier (BN fold) or NULL; passes == 2: REQUIRED,
                                               the multiplier of the e4m3 correction accumulator (2^-11 / s_n) */
  const float* shift;                       /* [cout_pad] bias / BN shift or NULL        */
  int32_t relu;
  int32_t res_mode;
  const void* residual;                     /* bf16 padded rows [*][2*cout]              */
  iou_conv_segment res_seg[IOU_CONV_MAX_SEG]; /* geometry of residual (UPSAMPLE2: coarse map) */
  int32_t out_mode;
  void* out;                                /* PADDED: bf16 [rows][2*cout]               */
  void* out_dense[IOU_CONV_MAX_SEG];        /* DENSE: fp32 base pointer per segment      */
  int32_t dense_split;                      /* DENSE: columns >= split go to out_dense2  */
  void* out_dense2[IOU_CONV_MAX_SEG];
  int32_t num_seg;
  iou_conv_segment seg[IOU_CONV_MAX_SEG];
  int32_t passes;                           /* 3 = hi*hi+hi*lo+lo*hi (fp32-grade), 4 = +lo*lo, 1 = bf16;
                                               2 = IOU_FMT_F16F8 maps and weights: one fp16 MMA pass + one e4m3
                                               pass of doubled K (two bf16-pass equivalents), see below          */
  int64_t out_rows;                         /* PADDED: rows allocated behind `out` (TMA store clips there) */
  int64_t res_rows;                         /* rows allocated behind `residual`                             */
  int32_t diag_k;                           /* grouped conv (cin == cout, block_n == 64): output tile j contracts
                                               only input channels [64j, 64j+64); weight is [taps*cout][2*64]   */
  int32_t two_cta;                          /* 1: run as CTA pairs (tcgen05 cta_group::2, cluster of 2)      */
  /* Fused stride-2 phase split (ABI version 4; PADDED output, one segment, block_n % 64 == 0): the epilogue also
   * writes every output row (img, yp, xp) to phase_out[(yp&1)*2 + (xp&1)] (where non-NULL) in the layout of
   * iou_phase_split, i.e. at (u, v) = ((yp>>1)+1, (xp>>1)+1) of a zero-initialised [n][(h+1)/2+2][(w+1)/2+2][2*cout]
   * map.  phase_only != 0: `out` is not written at all (may be NULL) -- for maps only a stride-2 conv reads.     */
  void* phase_out[4];
  int32_t phase_only;
  /* Sources with fewer channels than `cin` (ABI version 5): src[i] is [src_rows][2*src_cin[i]] (0 = cin) and the taps
   * reading it contract only its src_cin[i] channels (the first src_cin[i] K columns of their weight rows, in both
   * planes) -- K-concatenated GEMMs in one accumulator, e.g. conv3(t2) + downsample(x) of a bottleneck's first block. */
  int32_t src_cin[IOU_CONV_MAX_SRC];
  /* Split-K through N-concatenated partial sums (ABI version 7; 0 or 1 = off): the sources hold k_split * cin channels;
   * output channels [j * cout/k_split, (j+1) * cout/k_split) are the PARTIAL sums of a (cout/k_split)-channel convolution
   * over input channels [j*cin, (j+1)*cin) (weight: [taps*cout_pad][2*cin], row block j = that channel slice).  A conv
   * with few output rows and a long K (FPN P6: 22 row tiles, K = 18 432) gets k_split times more work items of
   * 1/k_split the length; iou_sum_channel_groups adds the partial maps (+ bias).  cout_pad/block_n % k_split == 0. */
  int32_t k_split;
  /* Epilogue width (ABI version 8): 0 = library default (12 epilogue warps for the padded-rows convs that add a
   * same-geometry residual, 8 otherwise; IOU_WIDE overrides), 1 = 12 warps if eligible (passes == 2, padded-rows
   * output), -1 = 8 warps. */
  int32_t wide;
  /* Column-group maxima (ABI version 8; DENSE output only): with group_max_cols = G > 0 (a multiple of 16, >= 32,
   * dividing block_n and cout) the epilogue also writes, for every output pixel and every group g of G consecutive
   * output channels, TWO partial maxima whose max is max_c out[pixel][g*G + c]: group_max_out[seg] is fp32
   * [n_img*h*w][cout/G][2] (the two epilogue warps of a row quadrant take alternate 16-column chunks; each stores the
   * max over its own).  For retina_cls (G = num_classes) this is the per-anchor max class logit that
   * iou_get_bboxes_premax consumes instead of re-reading the class map (iou_aware_retina_head.py:510-519). */
  void* group_max_out[IOU_CONV_MAX_SEG];
  int32_t group_max_cols;
} iou_conv_desc;

typedef struct iou_conv_plan iou_conv_plan;
/* Validates the descriptor, encodes the TMA tensor maps, sizes the launch. */
int iou_conv_plan_create(const iou_conv_desc* desc, iou_conv_plan** plan_out);
int iou_conv_run(const iou_conv_plan* plan, void* stream);
void iou_conv_plan_destroy(iou_conv_plan* plan);
/* Chains two plans into ONE launch (ABI version 8): `second` must be a plain 1x1 conv whose only source is the padded-rows
 * output of `first` (also a plain 1x1 conv, optionally with a same-geometry residual), both as CTA pairs with passes == 2,
 * same segments, one N tile in `second` -- relu(bn3(conv3(t2)) + x) followed by the next bottleneck's conv1
 * (backbones/resnet.py:224-226,256-265).  Each CTA pair then computes both convs for its rows: when `first` has ONE N tile
 * its epilogue writes its output slabs straight into the second conv's A ring in shared memory (next to the TMA store of the
 * output map, which still happens), otherwise the second conv reads them back while they are still in L2.  Both plans must be
 * of the 8-warp kind (iou_conv_desc.wide = -1).  On success *plan_out == first (which now owns `second`; run
 * and destroy it like any plan); on IOU_ERR_INVALID both plans are untouched and can be run one after the other. */
int iou_conv_chain_plan_create(iou_conv_plan* first, iou_conv_plan* second, iou_conv_plan** plan_out);
/* 2*MAC flops the plan performs on real (non-padding) outputs, for rooflines. */
double iou_conv_plan_flops(const iou_conv_plan* plan);
/* Epilogue warps of the kernel variant the plan launches: 8, or 12 for the wide variant (iou_conv_desc.wide). */
int iou_conv_plan_epilogue_warps(const iou_conv_plan* plan);

/* ------------------------------------------------------------------ layout kernels
 * (elementwise, HBM-bound helpers around the conv engine)                      */
/* NCHW fp32 -> padded rows bf16 hi|lo. */
int iou_pack_nchw(const float* src, int n, int c, int h, int w, void* dst, int64_t dst_row_start,
                  void* stream);
/* padded rows bf16 hi|lo -> NCHW fp32 (hi + lo). */
int iou_unpack_nchw(const void* src, int64_t src_row_start, int n, int c, int h, int w, float* dst,
                    void* stream);
/* Stem im2col (resnet.py:454-462, conv 7x7 s2 p3 on 3 channels): NCHW fp32 image
 * -> padded rows [n][(ho+2)][(wo+2)][2*kpad] with k = (r*7+s)*3 + c, zero padded to kpad. */
int iou_im2col_stem(const float* img, int n, int h, int w, int kpad, void* dst, void* stream);
/* Stem without an im2col buffer (resnet.py:454-462): NCHW fp32 image -> padded rows
 * [n][(ho+2)][(wo+2)][2*64] where the 64 "channels" of output pixel x' are the four horizontally
 * neighbouring pixels x'-2..x'+1 of the 2x2 space-to-depth image (16 channels each, 12 real):
 * kk = j*16 + (py*2+px)*3 + ch.  The 7x7/s2 conv is then 4 taps (dy=-2..1) of K=64 on iou_conv_run. */
int iou_stem_pack(const float* img, int n, int h, int w, void* dst, void* stream);
/* The same with the four VERTICAL neighbours y'-2..y'+1 at column x' in the 64 channels: the conv is then the 4 taps
 * dx = -2..1 at dy = 0, which share ONE A window in iou_conv_run (half the shared-memory fill of the stem conv). */
int iou_stem_pack_v(const float* img, int n, int h, int w, void* dst, void* stream);
/* 3x3 stride-2 pad-1 max pool (resnet.py:466) on padded rows (inputs >= 0). */
int iou_maxpool3x3s2(const void* src, int n, int c, int h, int w, void* dst, void* stream);
/* Splits a padded-rows map into the 4 stride-2 phase maps laid out in the OUTPUT
 * geometry ((h+1)/2 x (w+1)/2): phase[py][px][u][v] = in_padded[2(u-1)+py][2(v-1)+px]. */
int iou_phase_split(const void* src, int n, int c, int h, int w, void* const* dst4,
                    int phase_mask, void* stream);   /* bits 0..3: phases to write; bit 4: apply ReLU while copying */
/* The same layout kernels for either element format of a padded-rows map (ABI version 3).  Both formats spend
 * 4 bytes per element as [hi plane: c x 16 bit][lo plane: c x 16 bit], one 16-byte vector per 8 channels and plane:
 *   IOU_FMT_BF16X2: hi = bf16(v), lo = bf16(v - hi)                          (conv passes = 3; the *_fmt-less calls)
 *   IOU_FMT_F16F8 : hi = fp16(v), lo vector = [x8 x 8 | l8 x 8], x8 = e4m3(v), l8 = e4m3((v - hi) * 2^11)
 *                                                                            (conv passes = 2; needs c % 8 == 0)
 * iou_phase_split copies bytes and serves both (its fused ReLU needs the format: iou_phase_split_fmt). */
enum { IOU_FMT_BF16X2 = 0, IOU_FMT_F16F8 = 1 };
int iou_pack_nchw_fmt(const float* src, int n, int c, int h, int w, void* dst, int64_t dst_row_start, int fmt,
                      void* stream);
int iou_unpack_nchw_fmt(const void* src, int64_t src_row_start, int n, int c, int h, int w, float* dst, int fmt,
                        void* stream);
int iou_stem_pack_v_fmt(const float* img, int n, int h, int w, void* dst, int fmt, void* stream);
int iou_maxpool3x3s2_fmt(const void* src, int n, int c, int h, int w, void* dst, int fmt, void* stream);
int iou_phase_split_fmt(const void* src, int n, int c, int h, int w, void* const* dst4, int phase_mask, int fmt,
                        void* stream);       /* fmt only matters for the fused ReLU (bit 4 of phase_mask) */

/* ------------------------------------------------------------------ GroupNorm towers (IoUawareFCOSHead, "next" row rank 4)
 * In-place GroupNorm (+ReLU) of a padded-rows map holding num_seg segments (FPN levels): the norm layer of
 * ConvModule with norm_cfg type 'GN' (mmdet/models/utils/conv_module.py:140-163, norm.py:7,44-50;
 * torch.nn.GroupNorm: per (image, group) mean / biased variance over H x W x C/groups, eps inside the sqrt).
 * gamma/beta: device [c].  workspace: iou_group_norm_workspace_bytes(sum of n_img over segments, groups). */
size_t iou_group_norm_workspace_bytes(int total_images, int groups);
int iou_group_norm_relu(void* map, int c, int num_seg, const iou_conv_segment* seg, int groups,
                        const float* gamma, const float* beta, float eps, int relu, void* workspace,
                        size_t workspace_bytes, void* stream);
int iou_group_norm_relu_fmt(void* map, int c, int num_seg, const iou_conv_segment* seg, int groups,
                            const float* gamma, const float* beta, float eps, int relu, void* workspace,
                            size_t workspace_bytes, int fmt, void* stream);   /* either element format (IOU_FMT_*) */
/* Range statistics of a padded-rows map [rows][c] in either element format: out4[0] = max |v| (fp32 bits in the low
 * word), out4[1] = elements at or beyond the fp16 limit 65504 (the IOU_FMT_F16F8 encode SATURATES there -- the
 * reference's fp32 would not; inf / NaN count too), out4[2] = elements with |v| > 448 (their e4m3 parts are saturated:
 * fp16 precision only), out4[3] = non-zero elements.  The detector runs it over every activation map after the first
 * batch of a plan (api/detectors.py) -- the guard for activation ranges outside what the fp16 + e4m3 scheme holds. */
int iou_range_stats(const void* map, int64_t rows, int c, int fmt, uint64_t* out4, void* stream);
/* out[row][c] = bias[c] + sum_j part[row][j*c_out + c] on interior rows of one [n][h+2][w+2] padded-rows segment (border
 * rows are written as zeros): the reduction behind a k_split convolution.  part: [rows][2*groups*c_out], out:
 * [rows][2*c_out], both in element format fmt; bias may be NULL; c_out % 8 == 0. */
int iou_sum_channel_groups(const void* part, int n, int h, int w, int c_out, int groups, const float* bias, void* out,
                           int fmt, void* stream);
/* x[i] = exp(x[i] * scale) on n dense fp32 values: bbox_pred = scale(fcos_reg(feat)).exp()
 * (mmdet/models/anchor_heads/iou_aware_fcos_head.py:108). */
int iou_scale_exp(float* x, size_t n, float scale, void* stream);

/* ------------------------------------------------------------------ pre-processing ("next" row)
 * ImageTransform without the resize (mmdet/datasets/transforms.py:31-50; mmcv 0.2.8 imnormalize,
 * impad_to_multiple): src uint8 [n][h][w][3] (BGR, DEVICE pointer) -> dst fp32 [n][3][pad_h][pad_w] =
 * (pixel - mean) / std per output channel (channel order reversed if to_rgb), optional horizontal flip,
 * zero padding at the bottom / right.  mean3 / std3 are HOST pointers to 3 floats (output-channel order). */
int iou_preprocess_u8(const unsigned char* src, int n, int h, int w, int pad_h, int pad_w,
                      const float* mean3, const float* std3, int to_rgb, int flip, float* dst,
                      void* stream);

/* The same with the resize of ImageTransform.__call__ in front (transforms.py:33-40: mmcv.imrescale / imresize =
 * cv2.resize(..., INTER_LINEAR) on the uint8 frame): src [n][src_h][src_w][3] is resized to dst_h x dst_w with
 * OpenCV's 8-bit fixed-point bilinear arithmetic (bit-identical to cv2.resize), then normalised, flipped, padded
 * and transposed exactly as iou_preprocess_u8 does.  The caller computes dst_h/dst_w (mmcv imrescale rule). */
int iou_preprocess_resize_u8(const unsigned char* src, int n, int src_h, int src_w, int dst_h, int dst_w,
                             int pad_h, int pad_w, const float* mean3, const float* std3, int to_rgb, int flip,
                             float* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* IOU_B200_H_ */
