/* tools/sim_topk_cache.c -- development experiment (CPU): how many full Bid scans does a per-bidder top-K cache save?
 *
 * Prices only rise, so the value of every object for a bidder only falls.  After a full scan of bidder j keep its K best
 * objects and B = the K-th best value (an upper bound of every object outside the cache, for ever).  At a later bid of j,
 * evaluate the K cached objects at current prices: if the second largest of them is > B, (best, best index, second best)
 * are exactly what a full scan would return and the scan can be skipped.
 * Input: file of float32 [B][2][N][3] (pred, target); prints scans with / without the cache for several K.
 * Arithmetic = oracle/emd_oracle.c.   gcc -O2 -ffp-contract=off -o /tmp/sim tools/sim_topk_cache.c -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float sq3(float dx, float dy, float dz) { return fmaf(dz, dz, fmaf(dx, dx, dy * dy)); }
static inline float value(const float *x1, const float *x2, const float *price, int j, int k) {
    const float dx = x2[k * 3] - x1[j * 3], dy = x2[k * 3 + 1] - x1[j * 3 + 1], dz = x2[k * 3 + 2] - x1[j * 3 + 2];
    return (float)(3.0 - (double)sqrtf(sq3(dx, dy, dz)) - (double)price[k]);
}

#define KMAX 32
typedef struct { int n; int idx[KMAX]; float bound; } Cache;

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: sim file B N [K...]\n"); return 1; }
    FILE *f = fopen(argv[1], "rb");
    const int B = atoi(argv[2]), n = atoi(argv[3]);
    const float eps = 0.005f; const int iters = 50;
    float *data = malloc(sizeof(float) * (size_t)B * 2 * n * 3);
    if (fread(data, sizeof(float), (size_t)B * 2 * n * 3, f) != (size_t)B * 2 * n * 3) return 2;
    for (int a = 4; a <= argc; a++) {
        const int K = (a < argc) ? atoi(argv[a]) : 0;  /* last pass: K = 0 = no cache */
        long long scans = 0, hits = 0, bids = 0, mism = 0, late_scans = 0, late_bids = 0;
        long long per_iter_scans[64] = {0}, per_iter_bids[64] = {0};
        for (int b = 0; b < B; b++) {
            const float *x1 = data + (size_t)b * 2 * n * 3, *x2 = x1 + (size_t)n * 3;
            int *asg = malloc(sizeof(int) * n), *inv = malloc(sizeof(int) * n), *bid = malloc(sizeof(int) * n), *maxidx = malloc(sizeof(int) * n), *un = malloc(sizeof(int) * n);
            float *price = calloc(n, sizeof(float)), *inc = malloc(sizeof(float) * n), *maxinc = calloc(n, sizeof(float));
            Cache *C = calloc(n, sizeof(Cache));
            float *vals = malloc(sizeof(float) * n);
            for (int j = 0; j < n; j++) { asg[j] = -1; inv[j] = -1; maxidx[j] = 0; }
            for (int t = 0; t < iters; t++) {
                const int last = t == iters - 1;
                int u = 0;
                for (int j = 0; j < n; j++) if (asg[j] == -1) un[u++] = j;
                if (!u) break;
                for (int q = 0; q < u; q++) {
                    const int j = un[q];
                    bids++; per_iter_bids[t]++; if (t >= 2) late_bids++;
                    float best = -1e9f, better = -1e9f; int bi = -1;
                    int hit = 0;
                    if (K > 0 && C[j].n == K) {
                        /* cached candidates in ascending ORIGINAL index order give the reference's tie rule for free */
                        for (int c = 0; c < K; c++) {
                            const int k = C[j].idx[c];
                            const float d = value(x1, x2, price, j, k);
                            if (d > best || (d == best && k < bi)) { better = best; best = d; bi = k; }
                            else if (d > better) better = d;
                        }
                        if (better > C[j].bound) hit = 1;
                    }
                    if (!hit) {
                        scans++; per_iter_scans[t]++; if (t >= 2) late_scans++;
                        best = -1e9f; better = -1e9f; bi = -1;
                        for (int k = 0; k < n; k++) {
                            const float d = value(x1, x2, price, j, k);
                            vals[k] = d;
                            if (d > best) { better = best; best = d; bi = k; }
                            else if (d > better) better = d;
                        }
                        if (K > 0 && K <= n) { /* top-K by selection (simulation only) */
                            C[j].n = K;
                            for (int c = 0; c < K; c++) {
                                int m = -1; float mv = -3e38f;
                                for (int k = 0; k < n; k++) if (vals[k] > mv) { mv = vals[k]; m = k; }
                                C[j].idx[c] = m; vals[m] = -3e38f; C[j].bound = mv;
                            }
                        }
                    } else {
                        hits++;
                        /* verify against the full scan */
                        float b2 = -1e9f, bt2 = -1e9f; int bi2 = -1;
                        for (int k = 0; k < n; k++) {
                            const float d = value(x1, x2, price, j, k);
                            if (d > b2) { bt2 = b2; b2 = d; bi2 = k; } else if (d > bt2) bt2 = d;
                        }
                        if (b2 != best || bt2 != better || bi2 != bi) mism++;
                    }
                    bid[j] = bi; inc[j] = best - better + eps;
                    if (inc[j] > maxinc[bi]) maxinc[bi] = inc[j];
                }
                for (int q = 0; q < u; q++) {
                    const int j = un[q], o = bid[j];
                    const double x = inc[j], m = maxinc[o];
                    if (x - 1e-6 <= m && m <= x + 1e-6) maxidx[o] = j;
                }
                for (int q = 0; q < u; q++) {
                    const int j = un[q], o = bid[j];
                    if (last || maxidx[o] == j) {
                        const int prev = inv[o];
                        if (!last && prev != -1) asg[prev] = -1;
                        inv[o] = j; asg[j] = o; price[o] += inc[j]; maxinc[o] = -1e9f;
                    }
                }
            }
            free(asg); free(inv); free(bid); free(maxidx); free(un); free(price); free(inc); free(maxinc); free(C); free(vals);
        }
        printf("K=%2d: bids %lld  full scans %lld (%.1f%%)  hits %lld  mismatches %lld | after iteration 2: scans %lld of %lld bids (%.1f%%)\n", K, bids, scans,
               100.0 * scans / bids, hits, mism, late_scans, late_bids, 100.0 * late_scans / (late_bids ? late_bids : 1));
        if (K == 8 || K == 0) {
            printf("   per iteration scans/bids:");
            for (int t = 0; t < iters; t += 3) printf(" %lld/%lld", per_iter_scans[t] / B, per_iter_bids[t] / B);
            printf("\n");
        }
    }
    return 0;
}
