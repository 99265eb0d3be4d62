/* CPU oracle (plain C) for the Lovasz-Softmax loss and the confusion matrix.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/_build/liblovasz_oracle.so by
 * oracle/Makefile; loaded through ctypes by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg, never by the product package.
 *
 * It restates the reference algorithm (arman-rn/RTDA-SemanticSegmentation,
 * losses/lovasz_loss.py) with the tie order and integer formulation the CUDA
 * path documents, so that full-size inputs can be checked bit for bit:
 *
 *   per class c present among the valid pixels            lovasz_loss.py:142-168
 *     err_i = |[label_i == c] - p_ic| over valid pixels   lovasz_loss.py:157-171
 *     sort err descending, ties by ascending pixel index  lovasz_loss.py:173  (stable)
 *     J_r   = 1 - (G - F_r) / (G + (r+1) - F_r)           lovasz_loss.py:60-66
 *     g_r   = J_r - J_{r-1}  (g_0 = J_0)                  lovasz_loss.py:68-69
 *     loss_c = sum_r err_(r) * g_r                        lovasz_loss.py:180
 *   loss = sequential fp32 mean over present classes      lovasz_loss.py:19-44,183
 *   dL/dp_ic = -(grad_out / K) * g_rank(i) * sgn([label_i==c] - p_ic)   (autograd)
 *
 * F_r = number of foreground elements among sorted positions 0..r, G = F_last.
 * The counts are exact integers converted to fp32 once; this is bit-identical
 * to the reference's fp32 cumsum for segments shorter than 2^24 (checked by
 * oracle/pin_against_reference.py through jaccard_grad_int).  loss_c is
 * accumulated in double and rounded to fp32 once (the reference's fp32
 * torch.dot differs from it by ordinary rounding, < 1e-6 relative).
 *
 * per_image = 1 is the composition SURVEY.md 0.2 defines: the same computation
 * per image, then the reference's sequential mean over images (an image with no
 * valid present class contributes the reference's int 0 and is still counted).
 *
 * The confusion matrix follows utils.py:62-89 (fast_hist) behind the
 * validation.py:102-107 ignore filter.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t key; /* ~bits(err): ascending key == descending error */
    uint32_t pix; /* pixel index inside the group (ascending == stable tie order) */
} pair_t;

/* LSD radix sort, 4 passes of 8 bits, stable. */
static void radix_sort_pairs(pair_t *a, pair_t *tmp, size_t n) {
    for (int pass = 0; pass < 4; ++pass) {
        size_t count[257];
        memset(count, 0, sizeof(count));
        int shift = pass * 8;
        for (size_t i = 0; i < n; ++i) count[((a[i].key >> shift) & 255u) + 1]++;
        for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
        for (size_t i = 0; i < n; ++i) tmp[count[(a[i].key >> shift) & 255u]++] = a[i];
        pair_t *t = a; a = tmp; tmp = t;
    }
    /* 4 passes: result is back in the original `a` */
}

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* One segment = one (group, class).  pix_base[] lists the valid pixels of the
 * group (indices into the group's pixel range), nvalid of them.  Writes the
 * unit gradient -g*sgn into unit_grad (NCHW) for those pixels.  Returns the
 * class loss; *present = 1 when the class has foreground. */
static double segment_loss(const float *probas, const int64_t *labels, int64_t img0, int C, int64_t HW,
                           int c, const uint32_t *valid_pix, size_t nvalid, pair_t *buf, pair_t *tmp,
                           float *unit_grad, int *present) {
    size_t G = 0;
    for (size_t k = 0; k < nvalid; ++k) {
        uint32_t pix = valid_pix[k];
        int64_t b = img0 + (int64_t)(pix / (uint64_t)HW), hw = (int64_t)(pix % (uint64_t)HW);
        float p = probas[(b * C + c) * HW + hw];
        int fg = labels[b * HW + hw] == c;
        float err = fabsf((fg ? 1.0f : 0.0f) - p);
        buf[k].key = ~f2u(err);
        buf[k].pix = pix;
        G += (size_t)fg;
    }
    *present = G > 0;
    if (!G) return 0.0;
    radix_sort_pairs(buf, tmp, nvalid);
    double acc = 0.0;
    size_t F = 0;
    float j_prev = 0.0f;
    for (size_t r = 0; r < nvalid; ++r) {
        uint32_t pix = buf[r].pix;
        int64_t b = img0 + (int64_t)(pix / (uint64_t)HW), hw = (int64_t)(pix % (uint64_t)HW);
        int fg = labels[b * HW + hw] == c;
        F += (size_t)fg;
        float inter = (float)(int64_t)(G - F);
        float uni = (float)(int64_t)(G + (r + 1) - F);
        float j = 1.0f - inter / uni;
        float g = r ? j - j_prev : j;
        j_prev = j;
        uint32_t eb = ~buf[r].key;
        float err; memcpy(&err, &eb, 4);
        acc += (double)err * (double)g;
        float p = probas[(b * C + c) * HW + hw];
        float x = (fg ? 1.0f : 0.0f) - p;
        float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
        unit_grad[(b * C + c) * HW + hw] = -(g * sgn);
    }
    return acc;
}

/* Returns 0 on success.  grad, class_loss, present_out, n_present_out may be NULL.
 *   class_loss [G_n*C] (fp32-rounded per-class losses, 0 where absent)
 *   present    [G_n*C], n_present [G_n]  with G_n = per_image ? B : 1. */
int lovasz_oracle_c(const float *probas, const int64_t *labels, int64_t B, int C, int64_t HW,
                    int64_t ignore, int has_ignore, int per_image, float grad_out,
                    float *loss_out, float *grad, float *class_loss, int32_t *present_out, int32_t *n_present_out) {
    int64_t groups = per_image ? B : 1;
    int64_t imgs_per_group = per_image ? 1 : B;
    int64_t n = imgs_per_group * HW;
    if (n >= ((int64_t)1 << 31) || C < 1) return -1;
    size_t total = (size_t)(B * C * HW);
    float *unit = (float *)calloc(total ? total : 1, sizeof(float));
    uint32_t *valid_pix = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    float *scale = (float *)calloc((size_t)(groups ? groups : 1), sizeof(float));
    float *cl = (float *)calloc((size_t)C, sizeof(float));
    int *pr = (int *)calloc((size_t)C, sizeof(int));
    if (!unit || !valid_pix || !scale || !cl || !pr) return -2;

    float acc_img = 0.0f;
    for (int64_t g = 0; g < groups; ++g) {
        int64_t img0 = g * imgs_per_group;
        size_t nvalid = 0;
        for (int64_t i = 0; i < n; ++i) {                  /* lovasz_loss.py:157 valid mask */
            int64_t lab = labels[img0 * HW + i];
            int ok = !(has_ignore && lab == ignore) && lab >= 0 && lab < C;
            if (ok) valid_pix[nvalid++] = (uint32_t)i;
        }
#pragma omp parallel
        {
            pair_t *buf = (pair_t *)malloc(sizeof(pair_t) * (nvalid ? nvalid : 1));
            pair_t *tmp = (pair_t *)malloc(sizeof(pair_t) * (nvalid ? nvalid : 1));
#pragma omp for schedule(dynamic, 1)
            for (int c = 0; c < C; ++c) {
                int pres = 0;
                double l = segment_loss(probas, labels, img0, C, HW, c, valid_pix, nvalid, buf, tmp, unit, &pres);
                cl[c] = pres ? (float)l : 0.0f;
                pr[c] = pres;
            }
            free(buf);
            free(tmp);
        }
        float acc = 0.0f;
        int k = 0;
        for (int c = 0; c < C; ++c) {                      /* lovasz_loss.py:39-44: sequential, ascending class */
            if (class_loss) class_loss[g * C + c] = cl[c];
            if (present_out) present_out[g * C + c] = pr[c];
            if (!pr[c]) continue;
            acc = k ? acc + cl[c] : cl[c];
            ++k;
        }
        float loss_g = (k > 1) ? acc / (float)k : acc;     /* k == 0 -> 0 (the reference's int 0) */
        if (n_present_out) n_present_out[g] = k;
        acc_img = g ? acc_img + loss_g : loss_g;
        float s = grad_out;                                /* autograd: div backward of both means */
        if (groups > 1) s = s / (float)groups;
        if (k > 1) s = s / (float)k;
        scale[g] = k ? s : 0.0f;
    }
    float loss = (groups > 1) ? acc_img / (float)groups : acc_img;
    if (loss_out) *loss_out = loss;
    if (grad) {
        for (int64_t b = 0; b < B; ++b) {
            float s = scale[per_image ? b : 0];
            for (int64_t i = 0; i < (int64_t)C * HW; ++i) {
                float u = unit[b * C * HW + i];
                grad[b * C * HW + i] = (u == 0.0f) ? 0.0f : u * s;
            }
        }
    }
    free(unit); free(valid_pix); free(scale); free(cl); free(pr);
    return 0;
}

/* fast_hist behind the validation ignore filter: hist[n*n] += counts. */
int confusion_oracle_c(const int64_t *labels, const int64_t *preds, int64_t n, int n_class,
                       int64_t ignore, int has_ignore, int64_t *hist) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t t = labels[i], p = preds[i];
        if (has_ignore && t == ignore) continue;           /* validation.py:102 */
        if (t < 0 || t >= n_class || p < 0 || p >= n_class) continue; /* utils.py:78-83 */
        hist[t * n_class + p] += 1;                        /* utils.py:84-87 */
    }
    return 0;
}
