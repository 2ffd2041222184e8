/*
 * oracle/chamfer_oracle.c -- CPU restatement of the Chamfer arithmetic the reference calls.
 * TEST INFRASTRUCTURE ONLY (same rules as emd_oracle.c: never on the product path).
 *
 * PARITY UNPINNED: the arithmetic lives in third-party pytorch3d==0.7.2 (requirements.txt:9),
 * which is neither under /root/reference nor installed here.  This file restates its published
 * algorithm (pytorch3d/loss/chamfer.py chamfer_distance with the defaults the reference uses,
 * pytorch3d/ops/knn.py knn_points K=1, csrc/knn/knn_cpu.cpp / knn.cu) as recalled in
 * SURVEY.md App. B, anchored on the reference's call sites pointcloud_vision/utils.py:211 and :228:
 *   - squared L2, dist = 0; for d: diff = p1[d]-p2[d]; dist += diff*diff  (fp32)
 *   - strict '<' while scanning targets in ascending index => lowest index wins exact ties
 *   - only the first length2 targets are scanned, only the first length1 queries are filled
 *   - cham_x = sum_i dist / clamp(len,1) per cloud; loss = sum_n cham_x / max(N,1) + same for y
 *   - backward: diff = 2*g*(p1 - p2[idx]); grad_p1 += diff; grad_p2[idx] -= diff
 * mode 0 = unfused ((dx*dx + dy*dy) + dz*dz, the CPU build's order; the default everywhere),
 * mode 1 = contracted fma(dz,dz, fma(dy,dy, dx*dx)) (what nvcc makes of the same loop in knn.cu).
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#define PCL_CLONES __attribute__((target_clones("fma", "default")))
#else
#define PCL_CLONES
#endif

/* One directed K=1 search for one cloud: queries p1 (l1 valid of P1 rows), targets p2 (l2 valid). */
PCL_CLONES
static void knn1(const float *p1, long r1, int l1, int P1, const float *p2, long r2, int l2, int D, int mode,
                 float *dist, int *idx) {
    for (int i = 0; i < P1; i++) { dist[i] = 0.f; idx[i] = 0; } /* knn.py: zeros; padded rows stay 0 */
    if (l2 <= 0) return;
    for (int i = 0; i < l1; i++) {
        float best = INFINITY; int bi = 0;
        const float *a = p1 + i * r1;
        for (int j = 0; j < l2; j++) {
            const float *b = p2 + j * r2;
            float d = 0.f;
            if (mode == 0) {
                for (int c = 0; c < D; c++) { const float df = a[c] - b[c]; d = d + df * df; }
            } else {
                for (int c = 0; c < D; c++) { const float df = a[c] - b[c]; d = fmaf(df, df, d); }
            }
            if (d < best) { best = d; bi = j; }
        }
        dist[i] = best; idx[i] = bi;
    }
}

/*
 * Forward.  x (B,P1,D), y (B,P2,D) with element strides; x_len/y_len nullable (=> P1/P2).
 * Outputs: dist_x,idx_x (B*P1), dist_y,idx_y (B*P2), cham[2*n+{0,1}] = per-cloud sum_i dist / clamp(len,1)
 * for the x and y direction (double-accumulated: the per-point distances are the bit-exact quantities,
 * the scalar loss = sum_n cham / max(B,1) is formed by the caller and is a tolerance quantity).
 * Clouds are independent: callers thread over batch slices.
 */
int chamfer_oracle_forward(const float *x, long xb, long xr, const int64_t *x_len,
                           const float *y, long yb, long yr, const int64_t *y_len,
                           int B, int P1, int P2, int D, int mode,
                           float *dist_x, int *idx_x, float *dist_y, int *idx_y, double *cham /* 2*B per-cloud means */) {
    if (B < 0 || P1 < 0 || P2 < 0 || D < 1) return -1;
    for (int n = 0; n < B; n++) {
        if (x_len && (x_len[n] < 0 || x_len[n] > P1)) return -2;
        if (y_len && (y_len[n] < 0 || y_len[n] > P2)) return -2;
    }
    double *cx = cham;
    for (int w = 0; w < 2 * B; w++) {
        const int n = w >> 1, dir = w & 1;
        const int l1 = x_len ? (int)x_len[n] : P1, l2 = y_len ? (int)y_len[n] : P2;
        double s = 0.0;
        if (dir == 0) {
            knn1(x + n * xb, xr, l1, P1, y + n * yb, yr, l2, D, mode, dist_x + (size_t)n * P1, idx_x + (size_t)n * P1);
            for (int i = 0; i < l1; i++) s += dist_x[(size_t)n * P1 + i];
            cx[2 * n] = s / (double)(l1 > 1 ? l1 : 1);
        } else {
            knn1(y + n * yb, yr, l2, P2, x + n * xb, xr, l1, D, mode, dist_y + (size_t)n * P2, idx_y + (size_t)n * P2);
            for (int i = 0; i < l2; i++) s += dist_y[(size_t)n * P2 + i];
            cx[2 * n + 1] = s / (double)(l2 > 1 ? l2 : 1);
        }
    }
    return 0;
}

/*
 * Backward of loss = loss_x + loss_y for upstream gradient g (scalar).  grad_x (B,P1,D), grad_y (B,P2,D)
 * dense, overwritten.  Sequential fp32 accumulation in the order pytorch3d's CPU path applies it:
 * x-direction knn backward first (own term into grad_x, scatter into grad_y), then the y-direction.
 */
int chamfer_oracle_backward(const float *x, long xb, long xr, const int64_t *x_len,
                            const float *y, long yb, long yr, const int64_t *y_len,
                            int B, int P1, int P2, int D, const int *idx_x, const int *idx_y, float g,
                            float *grad_x, float *grad_y) {
    memset(grad_x, 0, sizeof(float) * (size_t)B * P1 * D);
    memset(grad_y, 0, sizeof(float) * (size_t)B * P2 * D);
    const float nb = (float)(B > 1 ? B : 1);
    for (int n = 0; n < B; n++) {
        const int l1 = x_len ? (int)x_len[n] : P1, l2 = y_len ? (int)y_len[n] : P2;
        const float gx = g / nb / (float)(l1 > 1 ? l1 : 1); /* d loss / d dist_x[n,i] */
        const float gy = g / nb / (float)(l2 > 1 ? l2 : 1);
        float *gxn = grad_x + (size_t)n * P1 * D, *gyn = grad_y + (size_t)n * P2 * D;
        if (l2 > 0)
            for (int i = 0; i < l1; i++) {
                const int j = idx_x[(size_t)n * P1 + i];
                for (int c = 0; c < D; c++) {
                    const float diff = 2.0f * gx * (x[n * xb + i * xr + c] - y[n * yb + j * yr + c]);
                    gxn[i * D + c] += diff; gyn[j * D + c] -= diff;
                }
            }
        if (l1 > 0)
            for (int j = 0; j < l2; j++) {
                const int i = idx_y[(size_t)n * P2 + j];
                for (int c = 0; c < D; c++) {
                    const float diff = 2.0f * gy * (y[n * yb + j * yr + c] - x[n * xb + i * xr + c]);
                    gyn[j * D + c] += diff; gxn[i * D + c] -= diff;
                }
            }
    }
    return 0;
}
