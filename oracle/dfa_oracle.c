/*
 * dfa_oracle.c — CPU restatement of HiP-AD's deformable feature aggregation.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_golden.py)
 * against fixtures produced by importing the reference's own torch path
 * (projects/mmdet3d_plugin/models/blocks.py:216-264) in the build container
 * (tests/golden/make_golden.py), and — on the GPU box — against the
 * reference CUDA op compiled from its own sources into oracle/_ref/.
 *
 * What is restated (all citations relative to /root/reference):
 *   projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu
 *     :13-59    bilinear_sampling              -> quad_setup(), dfa_oracle_forward()
 *     :62-126   bilinear_sampling_grad         -> dfa_oracle_backward()
 *     :129-187  deformable_aggregation_kernel  -> index math / validity in quad_setup() and the callers
 *     :190-262  deformable_aggregation_grad_kernel
 *   projects/mmdet3d_plugin/ops/src/deformable_aggregation.cpp:22-28 (layouts)
 *
 * Layouts (row-major, last index fastest):
 *   feat      [bs, num_feat, C]            float
 *   shapes    [cams, L, 2]   (h, w)        int32
 *   starts    [cams, L]      absolute row  int32
 *   loc       [bs, A, P, cams, 2] (x, y)   float, normalised
 *   weights   [bs, A, P, cams, L, G]       float
 *   out       [bs, A, C]                   float
 *
 * Arithmetic notes:
 *  - the compiled reference evaluates `loc*size - 0.5` as ONE fp32 FMA
 *    (SURVEY.md §7 "Bit-exact indices"), so fmaf() is used here; the integer
 *    corner indices and row offsets produced below are the bit-exact contract.
 *  - the bilinear value is formed in fp32 in the reference's operand order;
 *    the cross-sample accumulation (atomicAdd in the reference, order
 *    unspecified) is done in double here so the oracle is the order-free
 *    centre that both implementations are compared against.
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int valid;          /* 0 < x < 1 && 0 < y < 1 (cu:168-171) */
    int h, w;
    int h_low, w_low;
    float lh, lw, hh, hw;
    int64_t base_row;   /* b*num_feat + start[cam,l] */
    int ok[4];          /* corner in-bounds flags, order v1..v4 (cu:34-53) */
    int64_t row[4];     /* absolute feature row of each corner (valid only when ok) */
    float cw[4];        /* w1..w4 (cu:55) */
} dfa_quad;

static void quad_setup(dfa_quad *q, float loc_w, float loc_h, int h, int w,
                       int64_t base_row) {
    q->h = h; q->w = w; q->base_row = base_row;
    q->valid = (loc_w > 0.f && loc_w < 1.f && loc_h > 0.f && loc_h < 1.f);
    /* cu:180-181 as compiled: single fp32 FMA */
    const float h_im = fmaf(loc_h, (float)h, -0.5f);
    const float w_im = fmaf(loc_w, (float)w, -0.5f);
    const int h_low = (int)floorf(h_im);
    const int w_low = (int)floorf(w_im);
    const int h_high = h_low + 1, w_high = w_low + 1;
    q->h_low = h_low; q->w_low = w_low;
    q->lh = h_im - (float)h_low;
    q->lw = w_im - (float)w_low;
    q->hh = 1.f - q->lh;
    q->hw = 1.f - q->lw;
    q->ok[0] = (h_low >= 0 && w_low >= 0);
    q->ok[1] = (h_low >= 0 && w_high <= w - 1);
    q->ok[2] = (h_high <= h - 1 && w_low >= 0);
    q->ok[3] = (h_high <= h - 1 && w_high <= w - 1);
    q->row[0] = base_row + (int64_t)h_low * w + w_low;
    q->row[1] = q->row[0] + 1;
    q->row[2] = q->row[0] + w;
    q->row[3] = q->row[2] + 1;
    q->cw[0] = q->hh * q->hw; q->cw[1] = q->hh * q->lw;
    q->cw[2] = q->lh * q->hw; q->cw[3] = q->lh * q->lw;
}

/* out[b,a,c] = sum over valid (p,cam), all l: weight * bilinear  (cu:183-186) */
void dfa_oracle_forward(float *out, const float *feat, const int32_t *shapes,
                        const int32_t *starts, const float *loc,
                        const float *weights, int bs, int cams, int num_feat,
                        int C, int L, int A, int P, int G) {
    const int gd = C / G;
    const int64_t n_anchor = (int64_t)bs * A;
#pragma omp parallel
    {
        double *acc = (double *)malloc(sizeof(double) * (size_t)C);
#pragma omp for schedule(dynamic, 4)
        for (int64_t ba = 0; ba < n_anchor; ++ba) {
            const int b = (int)(ba / A);
            for (int c = 0; c < C; ++c) acc[c] = 0.0;
            for (int p = 0; p < P; ++p)
                for (int cam = 0; cam < cams; ++cam) {
                    const int64_t s = (ba * P + p) * cams + cam;
                    const float lx = loc[s * 2], ly = loc[s * 2 + 1];
                    if (!(lx > 0.f && lx < 1.f && ly > 0.f && ly < 1.f)) continue;
                    for (int l = 0; l < L; ++l) {
                        dfa_quad q;
                        const int cl = cam * L + l;
                        quad_setup(&q, lx, ly, shapes[cl * 2], shapes[cl * 2 + 1],
                                   (int64_t)b * num_feat + starts[cl]);
                        const float *wp = weights + (s * L + l) * G;
                        const float *r[4];
                        for (int k = 0; k < 4; ++k)
                            r[k] = q.ok[k] ? feat + q.row[k] * C : NULL;
                        for (int c = 0; c < C; ++c) {
                            const float v1 = r[0] ? r[0][c] : 0.f, v2 = r[1] ? r[1][c] : 0.f;
                            const float v3 = r[2] ? r[2][c] : 0.f, v4 = r[3] ? r[3][c] : 0.f;
                            const float val = q.cw[0] * v1 + q.cw[1] * v2 + q.cw[2] * v3 + q.cw[3] * v4;
                            acc[c] += (double)(val * wp[c / gd]);
                        }
                    }
                }
            float *o = out + ba * C;
            for (int c = 0; c < C; ++c) o[c] = (float)acc[c];
        }
        free(acc);
    }
}

/*
 * Gradients (cu:62-126, 190-262).  Every output buffer is fully written
 * (zeros where the reference would leave its pre-zeroed buffer untouched).
 *   g_feat [bs, num_feat, C], g_loc [bs,A,P,cams,2], g_w [bs,A,P,cams,L,G]
 */
void dfa_oracle_backward(const float *feat, const int32_t *shapes,
                         const int32_t *starts, const float *loc,
                         const float *weights, const float *grad_out,
                         float *g_feat, float *g_loc, float *g_w, int bs,
                         int cams, int num_feat, int C, int L, int A, int P,
                         int G) {
    const int gd = C / G;
    const int64_t n_anchor = (int64_t)bs * A;
    const int64_t n_sample = n_anchor * P * cams;

    /* sample-major part: g_w and g_loc */
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t s = 0; s < n_sample; ++s) {
        const int64_t ba = s / ((int64_t)P * cams);
        const int cam = (int)(s % cams);
        const int b = (int)(ba / A);
        const float lx = loc[s * 2], ly = loc[s * 2 + 1];
        double gx = 0.0, gy = 0.0;
        float *gwp = g_w + s * L * G;
        for (int i = 0; i < L * G; ++i) gwp[i] = 0.f;
        if (lx > 0.f && lx < 1.f && ly > 0.f && ly < 1.f) {
            const float *go = grad_out + ba * C;
            for (int l = 0; l < L; ++l) {
                dfa_quad q;
                const int cl = cam * L + l;
                quad_setup(&q, lx, ly, shapes[cl * 2], shapes[cl * 2 + 1],
                           (int64_t)b * num_feat + starts[cl]);
                const float *wp = weights + (s * L + l) * G;
                const float *r[4];
                for (int k = 0; k < 4; ++k)
                    r[k] = q.ok[k] ? feat + q.row[k] * C : NULL;
                for (int g = 0; g < G; ++g) {
                    double gw = 0.0;
                    for (int c = g * gd; c < (g + 1) * gd; ++c) {
                        const float v1 = r[0] ? r[0][c] : 0.f, v2 = r[1] ? r[1][c] : 0.f;
                        const float v3 = r[2] ? r[2][c] : 0.f, v4 = r[3] ? r[3][c] : 0.f;
                        const float val = q.cw[0] * v1 + q.cw[1] * v2 + q.cw[2] * v3 + q.cw[3] * v4;
                        /* cu:79-121: signs of the coordinate derivatives */
                        const float ghw = -q.hw * v1 - q.lw * v2 + q.hw * v3 + q.lw * v4; /* d/dh */
                        const float gww = -q.hh * v1 + q.hh * v2 - q.lh * v3 + q.lh * v4; /* d/dw */
                        const float top = go[c] * wp[g];               /* cu:84 */
                        gw += (double)(go[c] * val);                   /* cu:123 */
                        gx += (double)((float)q.w * gww * top);        /* cu:124 */
                        gy += (double)((float)q.h * ghw * top);        /* cu:125 */
                    }
                    gwp[l * G + g] = (float)gw;
                }
            }
        }
        g_loc[s * 2] = (float)gx;
        g_loc[s * 2 + 1] = (float)gy;
    }

    /* feature-major part: g_feat.  Parallel over channel slices so that each
     * thread owns its addresses (no races, fixed order). */
    const int64_t n_rows = (int64_t)bs * num_feat;
    double *acc = (double *)calloc((size_t)(n_rows * C), sizeof(double));
#pragma omp parallel
    {
        int nt = 1, tid = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads(); tid = omp_get_thread_num();
#endif
        const int c0 = (int)((int64_t)C * tid / nt), c1 = (int)((int64_t)C * (tid + 1) / nt);
        if (c1 > c0)
            for (int64_t s = 0; s < n_sample; ++s) {
                const float lx = loc[s * 2], ly = loc[s * 2 + 1];
                if (!(lx > 0.f && lx < 1.f && ly > 0.f && ly < 1.f)) continue;
                const int64_t ba = s / ((int64_t)P * cams);
                const int cam = (int)(s % cams);
                const int b = (int)(ba / A);
                const float *go = grad_out + ba * C;
                for (int l = 0; l < L; ++l) {
                    dfa_quad q;
                    const int cl = cam * L + l;
                    quad_setup(&q, lx, ly, shapes[cl * 2], shapes[cl * 2 + 1],
                               (int64_t)b * num_feat + starts[cl]);
                    const float *wp = weights + (s * L + l) * G;
                    for (int k = 0; k < 4; ++k) {
                        if (!q.ok[k]) continue;
                        double *dst = acc + q.row[k] * C;
                        for (int c = c0; c < c1; ++c) {
                            const float top = go[c] * wp[c / gd];
                            dst[c] += (double)(q.cw[k] * top);         /* cu:95,103,111,119 */
                        }
                    }
                }
            }
    }
    for (int64_t i = 0; i < n_rows * C; ++i) g_feat[i] = (float)acc[i];
    free(acc);
}

/*
 * Integer sampling contract, one record per (b,a,p,cam,l):
 *   idx[.., 0] = valid flag (sample inside (0,1)^2)
 *   idx[.., 1] = h_low, idx[.., 2] = w_low
 *   idx[.., 3] = level offset  = starts[cam,l]            (absolute row, per sample b excluded)
 *   idx[.., 4] = corner mask   = ok1 | ok2<<1 | ok3<<2 | ok4<<3
 *   idx[.., 5] = row of corner 1 within the sample = starts + h_low*w + w_low
 * h_low/w_low/mask/row are reported as 0 for invalid samples.
 */
void dfa_oracle_indices(int32_t *idx, const int32_t *shapes, const int32_t *starts,
                        const float *loc, int bs, int cams, int L, int A, int P) {
    const int64_t n_sample = (int64_t)bs * A * P * cams;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n_sample; ++s) {
        const int cam = (int)(s % cams);
        const float lx = loc[s * 2], ly = loc[s * 2 + 1];
        for (int l = 0; l < L; ++l) {
            dfa_quad q;
            const int cl = cam * L + l;
            quad_setup(&q, lx, ly, shapes[cl * 2], shapes[cl * 2 + 1], starts[cl]);
            int32_t *o = idx + (s * L + l) * 6;
            o[0] = q.valid;
            o[3] = starts[cl];
            if (q.valid) {
                o[1] = q.h_low; o[2] = q.w_low;
                o[4] = q.ok[0] | (q.ok[1] << 1) | (q.ok[2] << 2) | (q.ok[3] << 3);
                o[5] = (int32_t)q.row[0];
            } else {
                o[1] = o[2] = o[4] = o[5] = 0;
            }
        }
    }
}

/* Number of distinct feature rows touched by >=1 in-bounds corner of >=1 valid
 * sample (SURVEY.md §8d "U"): the algorithmic-bytes numerator of the roofline. */
int64_t dfa_oracle_unique_rows(const int32_t *shapes, const int32_t *starts,
                               const float *loc, int bs, int cams, int num_feat,
                               int L, int A, int P) {
    const int64_t n_rows = (int64_t)bs * num_feat;
    unsigned char *hit = (unsigned char *)calloc((size_t)n_rows, 1);
    const int64_t n_sample = (int64_t)bs * A * P * cams;
    for (int64_t s = 0; s < n_sample; ++s) {
        const float lx = loc[s * 2], ly = loc[s * 2 + 1];
        if (!(lx > 0.f && lx < 1.f && ly > 0.f && ly < 1.f)) continue;
        const int cam = (int)(s % cams);
        const int b = (int)(s / ((int64_t)A * P * cams));
        for (int l = 0; l < L; ++l) {
            dfa_quad q;
            const int cl = cam * L + l;
            quad_setup(&q, lx, ly, shapes[cl * 2], shapes[cl * 2 + 1],
                       (int64_t)b * num_feat + starts[cl]);
            for (int k = 0; k < 4; ++k)
                if (q.ok[k]) hit[q.row[k]] = 1;
        }
    }
    int64_t u = 0;
    for (int64_t i = 0; i < n_rows; ++i) u += hit[i];
    free(hit);
    return u;
}
