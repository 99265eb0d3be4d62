/* rtda_lovasz.h - C ABI of the B200-native Lovasz-Softmax / confusion-matrix hot path.
 *
 * The reference (arman-rn/RTDA-SemanticSegmentation) is pure Python and has no
 * FFI of its own; these entry points are what a binding of its hot path binds:
 *
 *   rtda_lovasz_fwd / rtda_lovasz_bwd   replace LovaszSoftmax.forward
 *       (losses/lovasz_loss.py:185-200: prob_flatten :93-120, lovasz_softmax_flat
 *       :122-183, lovasz_grad :47-70, mean :19-44) and its autograd backward,
 *       as called from train_lovasz.py:102,124,282 and :108,129,292.
 *   rtda_confusion_accumulate           replaces the validation.py:98-109 slice
 *       (ignore filter + utils.py:62-89 fast_hist + `hist +=`).
 *   rtda_confusion_argmax_accumulate    the same fused with validation.py:95
 *       (argmax over the class dimension of the logits).
 *
 * Conventions (all functions):
 *   - return 0 on success, > 0 a cudaError_t, < 0 an argument error (RTDA_E_*);
 *   - no allocation, no host synchronisation, no global state besides a
 *     per-device attribute cache: every buffer (workspace included) is owned by
 *     the caller, all work is enqueued on `stream` (a cudaStream_t passed as
 *     void*), so the calls are re-entrant per (device, stream);
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - probabilities / gradients are fp32 NCHW contiguous; 16-byte aligned when HW % 4 == 0
 *     (the float4 paths), 4-byte aligned otherwise (so a batch slice of an odd-sized map works).
 *
 * Label element types (`label_kind`): RTDA_LABEL_I64 is what the reference's
 * callers pass (`labels.long()`, train_lovasz.py:66); U8 / I32 are fast paths
 * for the dataset-native masks (datasets/cityscapes.py:73-82).
 */
#ifndef RTDA_LOVASZ_H
#define RTDA_LOVASZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTDA_ABI_VERSION 2

#define RTDA_E_ARG (-1)        /* null pointer / non-positive size */
#define RTDA_E_RANGE (-2)      /* a size exceeds what the kernels index (see below) */
#define RTDA_E_WORKSPACE (-3)  /* ws_bytes smaller than rtda_lovasz_workspace_bytes */
#define RTDA_E_ALIGN (-4)      /* a pointer is not aligned as the conventions above require */

#define RTDA_LABEL_I64 0
#define RTDA_LABEL_U8 1
#define RTDA_LABEL_I32 2

#define RTDA_MAX_CLASSES 44        /* Lovasz: per-class digit histograms live in shared memory */
#define RTDA_MAX_SEGMENT (1 << 30) /* Lovasz: pixels per sort segment (B*H*W global, H*W per image) */

/* ABI version of the loaded library (== RTDA_ABI_VERSION it was built with). */
int rtda_version(void);

/* Static string for a return code of this library (cudaGetErrorString for > 0). */
const char *rtda_error_string(int code);

/* ---- Lovasz-Softmax -------------------------------------------------------
 * Segments: per_image == 0 -> one sort problem per class over all B*HW pixels
 * (what LovaszSoftmax.forward computes, lovasz_loss.py:197-199); per_image != 0
 * -> one per (image, class), averaged over images with the reference's `mean`
 * (the composition SURVEY.md 0.2 defines).  G = per_image ? B : 1 groups.
 *
 * Pixels whose label equals `ignore` (when has_ignore) or lies outside [0, C)
 * take no part (lovasz_loss.py:157-161): they are dropped before the sort and
 * their gradient is 0.  The number of out-of-range labels is reported in
 * *label_errors so a caller can raise the reference's IndexError.
 */

/* Bytes of scratch rtda_lovasz_fwd needs for this shape. */
int rtda_lovasz_workspace_bytes(int64_t B, int C, int64_t HW, int per_image, size_t *bytes);

/* Debug / test aid: byte offsets of the regions inside the workspace, in the order
 * {hist, fghist, tickets, desc32, desc64, zero_end, base, fgbase, seg_fg, codes,
 *  partial, buf_a, buf_b, total, segments S, tiles per segment T,
 *  fb_flag (u32 [S]: 1 where the segment was sorted by the LSD radix path), nb (u32 [S] buckets),
 *  cursor (u64 [S][NBK]: elements | fg elements << 32 per bucket), table, partial_b,
 *  level-1 tiles per segment T1, bucket regions per segment NBK, bucket capacity}.  offsets[24]. */
int rtda_lovasz_debug_layout(int64_t B, int C, int64_t HW, int per_image, size_t *offsets);

/* Sort path of the forward: 0 (default) = sample-partitioned bucket sort, with a per-segment,
 * device-side fallback to the LSD radix sort for key distributions it cannot split (massive exact
 * ties); 1 = LSD radix sort for every segment.  Both produce bit-identical results (descending
 * error, ties by ascending pixel index).  Process-wide; a test / benchmark aid.  The environment
 * variable RTDA_SORT_PATH=lsd selects 1 at the first call. */
int rtda_lovasz_set_sort_path(int path);

/* Where the forward leaves the gradient (`grad_mode`):
 *   RTDA_GRAD_FINAL  `grad` [B][C][HW] fp32 (the caller's NCHW layout) receives d(loss * h)/d(probas) for
 *                    h = grad_out_hint - the grad_out the caller expects to pass to the backward: 1 for a plain
 *                    loss.backward(), the loss weight, the share of a batch sharded over ranks - complete: the 1/K and 1/G factors of the two means applied with the same
 *                    fp32 divisions autograd performs, zeros at ignored pixels and in absent classes.  The
 *                    sort kernels write it directly; there is no second pass over it.  The backward is
 *                    rtda_lovasz_bwd_inplace: nothing to do for grad_out = h, one in-place multiplication
 *                    otherwise.  `applied_scale` [1] fp32 records which grad_out the buffer currently holds.
 *   RTDA_GRAD_UNIT   `grad` [G][C][n] fp32, n = B*HW/G, receives the unit gradient -g_rank * sgn(fg - p)
 *                    before the 1/K and 1/G factors; rows of absent classes and ignored pixels are not
 *                    defined (they may or may not be written: the backward never reads them).  The backward is rtda_lovasz_bwd (bit-exact for ANY grad_out) or, fused with the
 *                    softmax and cross-entropy backward, rtda_hybrid_bwd; both read `codes` as well.
 * Forward outputs:
 *   grad          see above; may be NULL (loss only, e.g. under torch.no_grad()).
 *   codes         [B*HW] u8, 4-byte aligned when HW % 4 == 0: the class of every pixel, 255 = ignored or
 *                 out of range; required with RTDA_GRAD_UNIT (its backward reads it), else may be NULL.
 *   applied_scale [1] fp32, set to grad_out_hint; required with RTDA_GRAD_FINAL, else may be NULL.
 *   loss          [1]    fp32  the scalar loss (0 when nothing is valid)
 *   class_loss    [G][C] fp32  per-class Lovasz losses (0 where absent)
 *   class_present [G][C] i32   1 where the class has a valid foreground pixel
 *   n_present     [G]    i32   K of each group
 *   group_loss    [G]    fp32  loss of each group (image); may be NULL
 *   label_errors  [1]    i32   number of labels outside [0,C) and != ignore; may be NULL
 *   packed        [4+2C] f64   partials for a batch sharded over ranks (per_image): {sum of group
 *                 losses, G, per-class sums of class_loss, per-class sums of class_present,
 *                 sum of -log p_label over the valid pixels, number of valid pixels} (the last two
 *                 are 0 from rtda_lovasz_fwd and filled by rtda_hybrid_fwd); one all-reduce(sum) of
 *                 this buffer gives the global means; may be NULL
 *
 * Non-finite probabilities: a NaN (or an infinite) error among the sorted elements of a class makes
 * that class's loss NaN (or +inf) and the means carry it into `loss`, as torch.dot does in the
 * reference (lovasz_loss.py:180).  The gradient at such elements is unspecified (finite).
 */
#define RTDA_GRAD_UNIT 0
#define RTDA_GRAD_FINAL 1
/* 1 when RTDA_GRAD_FINAL is available for this shape: always with one image per group (per_image, or B = 1),
 * where the NCHW layout [B][C][HW] is the segment layout [G][C][n]; with several images per group (global mode,
 * B > 1) when HW is a multiple of 4096 and C * B * HW < 2^30 (the sorted pairs then carry the element's NCHW
 * offset).  Otherwise rtda_lovasz_fwd returns RTDA_E_RANGE for it: use RTDA_GRAD_UNIT there. */
int rtda_lovasz_final_grad_supported(int64_t B, int C, int64_t HW, int per_image);
int rtda_lovasz_fwd(const float *probas, const void *labels, int label_kind,
                    int64_t B, int C, int64_t HW, int64_t ignore, int has_ignore, int per_image,
                    void *ws, size_t ws_bytes, int grad_mode, float grad_out_hint,
                    float *grad, uint8_t *codes, float *applied_scale, float *loss, float *class_loss,
                    int32_t *class_present, int32_t *n_present, float *group_loss,
                    int32_t *label_errors, double *packed, void *stream);

/* Backward of RTDA_GRAD_FINAL, in place: grad_probas (the forward's `grad`) becomes
 * d(loss * grad_out)/d(probas).  It holds the gradient for grad_out = *applied_scale; when grad_out equals
 * that (the plain loss.backward()) the kernel exits at once, otherwise every element is multiplied by
 * grad_out / *applied_scale and *applied_scale is updated - exact (bit-identical to the reference's fp32
 * sequence (grad_out / G) / K * g) whenever that ratio is a power of two (a loss weight of 0.5, an AMP
 * loss scale), within 3 ulp of it otherwise; rtda_lovasz_bwd is bit-exact for any grad_out.  grad_out is a
 * DEVICE scalar; `ws` is the forward's workspace (one zeroed counter in its header is used). */
int rtda_lovasz_bwd_inplace(const float *grad_out, float *grad_probas, float *applied_scale, void *ws,
                            int64_t B, int C, int64_t HW, void *stream);

/* Backward of RTDA_GRAD_UNIT: grad_probas[b][c][hw] = unit_grad * ((grad_out / G) / K_g) at valid pixels of
 * present classes, 0 elsewhere (the autograd of lovasz_loss.py:19-44,180); every element
 * of grad_probas [B][C][HW] is written.  grad_out is a DEVICE scalar; unit_grad, codes,
 * class_present and n_present are the forward's outputs. */
int rtda_lovasz_bwd(const float *grad_out, const float *unit_grad, const uint8_t *codes,
                    const int32_t *class_present, const int32_t *n_present, int64_t B, int C,
                    int64_t HW, int per_image, float *grad_probas, void *stream);

/* ---- Hybrid cross-entropy + Lovasz from logits ------------
 * What every caller computes around the loss (train_lovasz.py:98-105,120-127,280-289):
 *     loss_ce     = CrossEntropyLoss(ignore_index)(logits, labels)
 *     probas      = softmax(logits, dim=1)
 *     loss_lovasz = LovaszSoftmax(ignore)(probas, labels)
 * fused: softmax and the cross-entropy sum ride on the histogram kernel's read of the logits,
 * and the backward produces d(go_ce*loss_ce + go_lov*loss_lovasz)/d(logits) in one kernel.
 * probas [B][C][HW] fp32 is an OUTPUT of the forward (kept for the backward); n_valid [1] is the
 * number of pixels the cross-entropy averages over.  The other arguments are those of
 * rtda_lovasz_fwd / rtda_lovasz_bwd; grad_ce and grad_lovasz are DEVICE scalars. */
int rtda_hybrid_fwd(const float *logits, const void *labels, int label_kind,
                    int64_t B, int C, int64_t HW, int64_t ignore, int has_ignore, int per_image,
                    void *ws, size_t ws_bytes,
                    float *probas, float *unit_grad, uint8_t *codes, float *loss_lovasz, float *loss_ce,
                    float *class_loss, int32_t *class_present, int32_t *n_present, float *group_loss,
                    int32_t *label_errors, double *packed, int32_t *n_valid, void *stream);
int rtda_hybrid_bwd(const float *grad_ce, const float *grad_lovasz, const float *probas,
                    const float *unit_grad, const uint8_t *codes, const int32_t *class_present,
                    const int32_t *n_present, const int32_t *n_valid, int64_t B, int C, int64_t HW,
                    int per_image, float *grad_logits, void *stream);

/* ---- Confusion matrix -----------------------------------------------------
 * hist[t * n_class + p] += 1 for every pixel with label t != ignore,
 * 0 <= t < n_class and 0 <= p < n_class (validation.py:102-107 + utils.py:78-87).
 * hist is int64 [n_class * n_class] and is accumulated into, not overwritten.
 */
int rtda_confusion_accumulate(const void *labels, int label_kind, const void *preds, int pred_kind,
                              int64_t n, int n_class, int64_t ignore, int has_ignore,
                              int64_t *hist, void *stream);

/* Same, with the prediction taken as argmax over the class dimension of fp32
 * NCHW logits [B][n_class][HW] (first maximum wins, like torch.argmax). */
int rtda_confusion_argmax_accumulate(const float *logits, const void *labels, int label_kind,
                                     int64_t B, int n_class, int64_t HW, int64_t ignore, int has_ignore,
                                     int64_t *hist, void *stream);

/* ---- Per-kernel timing (benchmark aid) ---------------------------------------
 * When enabled, every kernel launch of this library is bracketed by two CUDA
 * events recorded on the launch stream.  rtda_profile_read synchronises the
 * device, ADDS the elapsed milliseconds and launch counts per stage since the
 * last read into the caller's host arrays, and recycles the events.  Off by
 * default; meant for single-threaded benchmark use (the event list is global).
 * Stages: see rtda_profile_stage_name(0 .. RTDA_PROFILE_STAGES-1). */
#define RTDA_PROFILE_STAGES 17
int rtda_profile_enable(int on);
int rtda_profile_read(double *stage_ms_host, int64_t *stage_launches_host, int n_stages);
const char *rtda_profile_stage_name(int stage);

#ifdef __cplusplus
}
#endif
#endif /* RTDA_LOVASZ_H */
