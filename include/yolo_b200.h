/*
 * yolo_b200.h — C-ABI of libyolo_b200.so: the per-box detection hot path of
 * KhaledSharif/yolo-from-scratch as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI of its own (it is pure Python on PyTorch); each entry point below
 * replaces one Python-level function of reference `train.py` (cited as file:line) and is what a
 * ctypes binding in that file would call.  INTEGRATION.md shows the binding.
 *
 * Conventions (all entry points):
 *   - every pointer is a CALLER-OWNED DEVICE pointer unless the name ends in `_host`;
 *     the library never allocates, frees or retains memory (the tests-only yb_selftest_sigmoid aside);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream); no entry point synchronises the device or the stream.  One entry point,
 *     yb_batched_nms with YB_NMS_GRAPH, additionally uses a library-owned side stream between two
 *     events (fork after the work already enqueued on `stream`, join before its last kernel): its
 *     score sort overlaps with the rest; every result is still ordered after `stream`, and under
 *     stream capture the fork/join becomes two parallel branches of the captured graph;
 *   - return value 0 = success, otherwise a cudaError_t or a negative YB_E* code;
 *     yb_last_error() returns a thread-local, human readable message for the last failure;
 *   - head tensors use the model-output layout of train.py:608-609: contiguous fp32
 *     (B, H, W, A, 5+nc), row = [tx, ty, tw, th, obj, cls...] raw logits;
 *   - `anchors` are fp32 (A,2) = [w,h] in pixels (train.py:372-374, 386-388);
 *   - tensors must be 16-byte aligned (every torch allocation is).
 */
#ifndef YOLO_B200_H
#define YOLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB_MAX_SCALES 4
#define YB_MAX_ANCHORS 8

/* Layout of head tensors (pred / grad): the model-output contract of train.py:608-609,
 * (B, H, W, A, 5+nc), or — f-2, SURVEY 8f — the head conv's own output (B, A*(5+nc), H, W), i.e. the
 * tensor BEFORE the reference's view/permute/contiguous: the path then reads the conv output directly
 * and that full read+write pass (and its backward) disappears.  Dense targets always keep the
 * reference layout; row indices, candidate order and all results are identical in both layouts. */
#define YB_LAYOUT_BHWAC 0
#define YB_LAYOUT_NCHW 1

#define YB_EINVAL (-1)   /* bad argument (shape, null pointer, alignment) */
#define YB_EWORKSPACE (-2) /* workspace too small */

/* Library/ABI version (major*100+minor) and last error text (thread local). */
int yb_version(void);
const char* yb_last_error(void);

/* ------------------------------------------------------------------------------------------
 * a-1  decode_predictions(raw_preds, anchors, img_size=640)            train.py:712-779
 *   out[...,0] = ((2*sigmoid(tx) - 0.5) + gx) / W        (:758)
 *   out[...,1] = ((2*sigmoid(ty) - 0.5) + gy) / H        (:759)
 *   out[...,2] = (aw/img) * (2*sigmoid(tw))^2            (:773)
 *   out[...,3] = (ah/img) * (2*sigmoid(th))^2            (:774)
 *   out[...,4:] = pred[...,4:]                            (:737, clone)
 * yb_decode_bwd is its vector-Jacobian product: grad_in = J^T grad_out (autograd of the above).
 * ---------------------------------------------------------------------------------------- */
int yb_decode_fwd(const float* pred, const float* anchors, float* out,
                  int B, int H, int W, int A, int nc, float img_size, void* stream);
int yb_decode_bwd(const float* pred, const float* anchors, const float* grad_out, float* grad_in,
                  int B, int H, int W, int A, int nc, float img_size, void* stream);

/* ------------------------------------------------------------------------------------------
 * a-2  ciou_loss(pred_boxes, target_boxes, eps=1e-7)                    train.py:634-710
 *   boxes are (N,4) xywh.  loss_out[0] = mean_i(1 - CIoU_i)  (NaN when N == 0, like torch).
 *   grad_pred / grad_tgt (nullable) receive d loss / d box (alpha is a constant, :701-702).
 *   `partials` is a caller-provided scratch of yb_ciou_scratch_bytes(N) bytes.
 * ---------------------------------------------------------------------------------------- */
size_t yb_ciou_scratch_bytes(long long N);
int yb_ciou_fwd_bwd(const float* pred_boxes, const float* tgt_boxes, long long N, float eps,
                    float* loss_out, float* grad_pred, float* grad_tgt,
                    void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a-3 / a-4  yolo_loss (train.py:781-838) and yolo_loss_multiscale (train.py:840-886),
 *            forward and backward fused.
 *
 * Stage 1  yb_loss_partials: one pass over every scale.
 *   For each scale s it accumulates partials[s] = { sum_pos(1-CIoU), n_pos, sum_all BCE(obj),
 *   sum_pos,cls BCE(cls) } (4 doubles per scale) over the LOCAL batch, and when grad[s] != NULL
 *   writes the dense gradient: the objectness column fully normalised by
 *   coef_obj[s] / (B_global*H*W*A), box/class columns of positive rows UN-normalised, zeros
 *   elsewhere.  Positive = target objectness > 0.5 (:809).  Decode uses `img_size` (the
 *   reference always passes its default 640 here, :796).
 * Between the stages a data-parallel caller all-reduces(sum) the S*4 doubles.
 * Stage 2  yb_loss_finalize: takes the (globally reduced) partials, scales the box/class
 *   gradient entries of the local positive rows by coef_box[s]/P_s and coef_cls[s]/(P_s*nc),
 *   and writes out4 = { total, sum_s bbox_s, sum_s obj_s, sum_s cls_s } plus, if per_scale != NULL,
 *   per_scale[s] = { bbox_s, obj_s, cls_s } (means, exactly the three values yolo_loss returns).
 *   total = sum_s (w_box*bbox_s + w_obj[s]*obj_s + w_cls*cls_s)   (:836, :879).
 *
 * coef_* are d(final scalar)/d(term_s): for `total.backward()` they are w_box, w_obj[s], w_cls.
 * `ws` is a caller-provided workspace of yb_loss_workspace_bytes(...) bytes shared by both stages.
 * ---------------------------------------------------------------------------------------- */
typedef struct yb_loss_desc {
    int S;                          /* number of scales, 1..YB_MAX_SCALES */
    int B;                          /* local batch */
    long long B_global;             /* global batch (== B on one GPU): obj mean normaliser */
    int A, nc;
    int H[YB_MAX_SCALES], W[YB_MAX_SCALES];
    float img_size;                 /* decode img_size, 640 in the reference's loss (:796) */
    float eps;                      /* CIoU eps, 1e-7 (:634) */
    float w_box, w_cls;             /* 0.05, 0.5 (:836) */
    float w_obj[YB_MAX_SCALES];     /* 1.0 for yolo_loss; [4.0,1.0,0.4] multiscale (:865) */
    float coef_box[YB_MAX_SCALES];  /* upstream coefficients for the gradient */
    float coef_obj[YB_MAX_SCALES];
    float coef_cls[YB_MAX_SCALES];
    const float* pred[YB_MAX_SCALES];
    const float* tgt[YB_MAX_SCALES];
    const float* anchors[YB_MAX_SCALES];
    float* grad[YB_MAX_SCALES];     /* nullable: forward only (torch.no_grad callers) */
    int layout;                     /* YB_LAYOUT_* of pred and grad */
} yb_loss_desc;

size_t yb_loss_workspace_bytes(const yb_loss_desc* d);
int yb_loss_partials(const yb_loss_desc* d, double* partials /* S*4 */,
                     void* ws, size_t ws_bytes, void* stream);
int yb_loss_finalize(const yb_loss_desc* d, const double* partials /* S*4, reduced */,
                     float* out4, float* per_scale /* S*3, nullable */,
                     void* ws, size_t ws_bytes, void* stream);

/* f-4: the same stage 1 fed by label lists instead of dense targets — the label loop of
 * YOLODataset.__getitem__ (train.py:147-205, as yb_build_targets) runs on the device and hands the
 * loss a sparse form of its result: per-scale lists of the assigned rows, their 32-byte target rows
 * and a 1-bit-per-row positive map.  Neither the dense targets (3 tensors the size of the heads) nor
 * their host->device copy (train.py:900-902) exist on this path; d->tgt[] is ignored.
 *   labels (B,max_gt,5) fp64, n_gt (B) int32, letterbox (B,5) fp64, anchors_all (S,A,2) fp32,
 *   assign_img_size = the dataset's img_size (:159-167); status as in yb_build_targets.
 * Results equal yb_loss_partials on the dense targets yb_build_targets would have produced.
 * Stage 2 is yb_loss_finalize with the same workspace. */
size_t yb_loss_sparse_workspace_bytes(const yb_loss_desc* d, int max_gt);
int yb_loss_partials_sparse(const yb_loss_desc* d, const double* labels, const int* n_gt,
                            const double* letterbox, const float* anchors_all, int max_gt,
                            int assign_img_size, int* status, double* partials /* S*4 */,
                            void* ws, size_t ws_bytes, void* stream);

/* x[i] *= *factor for i < n, skipped entirely (no memory traffic) when *factor == 1.0f.
 * Used by the autograd wrapper to apply the upstream gradient of `total` to the gradient the
 * fused kernel already produced (loss.backward() passes exactly 1.0).  factor is a device pointer. */
int yb_scale_inplace(float* x, long long n, const float* factor, void* stream);

/* ------------------------------------------------------------------------------------------
 * a-5  target assignment                                         train.py:108-131, 147-205
 *   yb_anchor_iou: YOLODataset.compute_anchor_iou for n boxes x A anchors (fp32, +1e-16).
 *   yb_build_targets: the label loop of YOLODataset.__getitem__ for a batch.
 *     labels   (B, max_gt, 5) fp64 = raw label lines [class, xc, yc, w, h]
 *     n_gt     (B) int32
 *     letterbox(B, 5) fp64 = [orig_w, orig_h, scale, pad_top, pad_left]  (:136-137)
 *     anchors  (S, A, 2) fp32
 *     targets[s] (B, G_s, G_s, A, 5+nc) fp32, fully written (zero-filled then assigned)
 *   Bit-exact with the reference: fp64 letterbox/cell math (:159-166,:184-189), fp32 shape IoU,
 *   strict '>' across scales then first argmax (:177-180), first GT wins a slot (:193),
 *   nc==1 writes class slot 5 irrespective of label id (:201-202).  Python's negative-index
 *   wrap-around for cells in [-G, 0) is reproduced; rows the reference would reject
 *   (IndexError) set bit 0 of *status (device int32, nullable).
 * ---------------------------------------------------------------------------------------- */
int yb_anchor_iou(const float* box_wh, const float* anchors, float* iou_out,
                  int n, int A, void* stream);
int yb_build_targets(const double* labels, const int* n_gt, const double* letterbox,
                     const float* anchors, float* const targets_host[YB_MAX_SCALES],
                     int B, int max_gt, int S, const int* G_host, int A, int nc, int img_size,
                     int* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * a-6  candidate filter + box build of predict()                  train.py:1152-1222, 1227-1229
 *   For every image b and every row in P3->P4->P5, row-major (gy,gx,a) order:
 *     keep rows with sigmoid(obj) > conf  (:1166-1167, objectness only)
 *     class prob/id = max sigmoid(cls) (first max; nc==1: slot 5, id 0)  (:1184-1189)
 *     xyxy in pixels, minus (pad_left,pad_top), divided by scale  (:1192-1213)
 *     score = sigmoid(obj) * class prob  (:1216)
 *   One pass over the heads (decoupled look-back over groups of 1,024 rows; the workspace is zeroed by a
 *   memset node of the same call).
 *   Output is per image, order preserving, with a fixed stride of `cap` candidates per image:
 *     boxes (B,cap,4) fp32, scores (B,cap) fp32, classes (B,cap) int64, counts (B) int32.
 *   letterbox (B,3) fp32 = [scale, pad_top, pad_left], NULL = identity.
 *   cap must be >= total rows per image (sum_s H_s*W_s*A) unless the caller knows better;
 *   candidates beyond cap are dropped and counts[b] saturates at cap.
 * ---------------------------------------------------------------------------------------- */
typedef struct yb_heads_desc {
    int S, B, A, nc;
    int H[YB_MAX_SCALES], W[YB_MAX_SCALES];
    float img_size;
    const float* pred[YB_MAX_SCALES];
    const float* anchors[YB_MAX_SCALES];
    int layout;                     /* YB_LAYOUT_* of pred (filter: both; decode/eval: BHWAC only) */
} yb_heads_desc;

size_t yb_filter_workspace_bytes(const yb_heads_desc* d);
int yb_filter_compact(const yb_heads_desc* d, float conf_thres, const float* letterbox,
                      float* boxes, float* scores, int64_t* classes, int* counts, int cap,
                      void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a-7  torchvision.ops.batched_nms(boxes, scores, idxs, iou_threshold)       train.py:1232-1233
 *      (torchvision 0.26.0, torchvision/ops/boxes.py:51-120 and its CUDA nms kernel)
 *   Batched over B independent images with a fixed stride of `cap` candidates per image and
 *   counts[b] valid entries (counts == NULL: every image has `cap`).
 *   classes == NULL gives plain torchvision.ops.nms.
 *   Arithmetic is torchvision's CUDA kernel bit for bit: fp32,
 *     inter = max(0, min(ax2,bx2)-max(ax1,bx1)) * max(0, min(ay2,by2)-max(ay1,by1))
 *     iou   = inter / ( fma(bx2-bx1, by2-by1, (ax2-ax1)*(ay2-ay1)) - inter ),  a = higher score
 *     suppress when iou > (float)iou_threshold; stable descending score order (lower index wins
 *     ties).
 *   trick_max_numel reproduces batched_nms' size dispatch (boxes.py:80): an image with
 *   4*counts[b] <= trick_max_numel uses the coordinate-offset trick (boxes + cls*(max+1), fp32,
 *   boxes.py:98-101), larger ones the per-class loop (boxes.py:113-120).  Pass 100000 for CUDA
 *   callers, 4000 for CPU callers, -1 to force per-class, LLONG_MAX to force the trick.
 *   Output: keep (B,cap) int64 = kept candidate indices (0..counts[b]) in descending score order,
 *   n_keep (B) int32.
 *
 *   algo selects how the same result is computed:
 *     YB_NMS_GRAPH    (default) sparse suppression graph: spatial/area-ordered tile culling, edge
 *                     list, parallel fixed-point resolve.  Workspace holds the edge list; an image
 *                     with more edges than fit (heavily clustered boxes) or class ids >= 512 is
 *                     resolved inside the same launch by a blocked greedy pass (M x kept pair
 *                     tests, exact arithmetic), so n_keep[b] is never negative for this algorithm.
 *     YB_NMS_BITMASK  dense blocked IoU bitmask + serial scan (torchvision's structure, batched);
 *                     class ids < 65536; never overflows with yb_nms_workspace_bytes().
 * ---------------------------------------------------------------------------------------- */
#define YB_NMS_GRAPH 0
#define YB_NMS_BITMASK 1
size_t yb_nms_workspace_bytes(int B, int cap);      /* enough for either algorithm, worst case */
size_t yb_nms_min_workspace_bytes(int B, int cap);  /* fixed part; the rest holds mask rows / edges */
size_t yb_nms_graph_workspace_bytes(int B, int cap, int edges_per_box); /* graph algo, sized edge list */
int yb_batched_nms(const float* boxes, const float* scores, const int64_t* classes,
                   const int* counts, int B, int cap, double iou_threshold,
                   long long trick_max_numel, int algo, int64_t* keep, int* n_keep,
                   void* ws, size_t ws_bytes, void* stream);

/* Statistics of the last YB_NMS_GRAPH call that used workspace `ws` (bench/tests; synchronises the
 * stream), summed over the B images: IoU pair tests executed by the half-precision filter, edges found, and
 * pairs that passed the filter and were decided exactly in fp32.  Outputs are HOST pointers (nullable). */
int yb_nms_graph_stats(const void* ws, size_t ws_bytes, int B, int cap, unsigned long long* evals_host,
                       unsigned long long* edges_host, unsigned long long* cands_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * detection list of predict()                                              train.py:1236-1246
 *   Packs the kept detections of a batch as rows (x1, y1, x2, y2, conf, class) fp32, image after
 *   image, each image in descending score order.  out must hold sum(n_keep)*6 floats
 *   (<= B*cap*6); offsets (B+1) int32 receives the first row of every image and the total.
 *   If any n_keep[b] is negative (NMS reported a failure for that image) the total offsets[B] is -1:
 *   a failed image never reads as "no detections".
 * ---------------------------------------------------------------------------------------- */
int yb_pack_detections(const float* boxes, const float* scores, const int64_t* classes,
                       const int64_t* keep, const int* n_keep, int B, int cap, float* out,
                       int* offsets, void* stream);

/* ------------------------------------------------------------------------------------------
 * f-1  detection counting of eval_epoch                      train.py:993-1024, 928-958
 *   For every row of every scale: p = sigmoid(pred obj), t = target obj (both compared with
 *   conf_threshold in double, as `.item()` values are in the reference);
 *     p > conf and t > conf : IoU(decoded pred xywh, target xywh; compute_box_iou, eps 1e-6, fp32)
 *                              > (float)iou_threshold ? TP : FP
 *     p > conf only         : FP          t > conf only : FN
 *   d->img_size is the decode img_size (the reference always uses its default 640 here, :993).
 *   targets_host: HOST array of S device pointers, same layout as d->pred.
 *   counts3 (device, int64[3]) += {TP, FP, FN}: the caller zeroes it once per epoch.
 * ---------------------------------------------------------------------------------------- */
int yb_eval_counts(const yb_heads_desc* d, const float* const* targets_host, double conf_threshold,
                   double iou_threshold, long long* counts3, void* stream);

/* Per-launch CUDA-event timing for bench.py: when enabled every kernel launch of the library is
 * bracketed by two events on its stream; yb_timing_collect synchronises them, writes
 * "name count total_ms" lines into buf (NUL terminated, truncated to n) and resets. */
void yb_timing_enable(int on);
int yb_timing_collect(char* buf, size_t n);

/* Counters for tests/bench: number of kernel launches this library has enqueued (process wide). */
unsigned long long yb_launch_count(void);

/* Self-test of the batched sigmoid (tests only; synchronises the stream).  The filter and the long-row decode
 * evaluate the reference's 1/(1+expf(-x)) (train.py:758-759,773-774,1157; ATen's sigmoid) with the IEEE
 * division's own fast path but without its range-check branch.  This runs both forms on the device over
 *   [0] every float y with 2^-126 <= |y| < 2^126 : 1.0f / y against the branch-free reciprocal,
 *   [1] every non-NaN float x whose 1+expf(-x) is inside that range : the two sigmoid forms,
 * and writes the number of differing bit patterns to mismatches_host[0..1] (both must be 0). */
int yb_selftest_sigmoid(unsigned long long* mismatches_host /* [2] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLO_B200_H */
