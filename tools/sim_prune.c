/* Development aid (CPU): how many (bidder, target) evaluations do different pruning layouts of the auction kernel
 * execute?  Replays the auction of oracle/emd_oracle.c on Morton-sorted clouds and counts, per iteration, the 32-target
 * tiles that survive the box test for groups of G neighbouring bidders (G = 32: lane-per-bidder warp, 8/4: sub-warp
 * groups, 1: warp-per-bidder), with the seed threshold (pessimistic) and the final second best (optimistic), for
 *   scheme 0: one tile bound max c over all 32 targets (the shipped kernel),
 *   scheme 1: free (never assigned, price 0) objects taken out of the tiles and packed into their own tiles.
 * Input: raw float32 file [B][2][N][3] (pred, target), written by tools/sim_prune.py.   gcc -O2 -o sim_prune sim_prune.c -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TILE 32
static unsigned spread6(unsigned v) { v = (v | (v << 8)) & 0x300F; v = (v | (v << 4)) & 0x30C3; v = (v | (v << 2)) & 0x9249; return v; }
static unsigned morton18(const float *p) {
    int q[3];
    for (int i = 0; i < 3; i++) { q[i] = (int)(p[i] * 64.f); if (q[i] < 0) q[i] = 0; if (q[i] > 63) q[i] = 63; }
    return spread6(q[0]) | (spread6(q[1]) << 1) | (spread6(q[2]) << 2);
}
typedef struct { unsigned long long key; } K;
static int cmpk(const void *a, const void *b) { unsigned long long x = ((const K *)a)->key, y = ((const K *)b)->key; return x < y ? -1 : x > y; }
static void msort(const float *src, int n, float *dst) {
    K *k = malloc(sizeof(K) * n);
    for (int i = 0; i < n; i++) k[i].key = ((unsigned long long)morton18(src + 3 * i) << 32) | (unsigned)i;
    qsort(k, n, sizeof(K), cmpk);
    for (int i = 0; i < n; i++) memcpy(dst + 3 * i, src + 3 * (k[i].key & 0xffffffffu), 12);
    free(k);
}
typedef struct { float lo[3], hi[3], cmax; int n; } Box;
static float boxd(const Box *b, const float *a) {
    float s = 0;
    for (int i = 0; i < 3; i++) { float d = fmaxf(fmaxf(b->lo[i] - a[i], a[i] - b->hi[i]), 0.f); s += d * d; }
    return sqrtf(s);
}
static int skippable(const Box *b, const float *a, float tm) { /* value upper bound of the tile < threshold */
    if (b->n == 0) return 1;
    return b->cmax - boxd(b, a) < tm;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: sim_prune file B N [verbose]\n"); return 1; }
    const int B = atoi(argv[2]), n = atoi(argv[3]), verbose = argc > 4;
    FILE *f = fopen(argv[1], "rb");
    float *raw = malloc(sizeof(float) * (size_t)B * 2 * n * 3);
    if (fread(raw, sizeof(float), (size_t)B * 2 * n * 3, f) != (size_t)B * 2 * n * 3) return 2;
    const int NT = (n + TILE - 1) / TILE, iters = 50;
    const float eps = 0.005f;
    const int Gs[4] = {32, 8, 4, 1};
    /* totals[scheme][thr][g] */
    double tot[2][2][4] = {{{0}}}, alg = 0, exact_needed = 0, exact_home = 0, home_tiles = 0;
    for (int c = 0; c < B; c++) {
        float *x1 = malloc(12 * n), *x2 = malloc(12 * n);
        msort(raw + ((size_t)c * 2) * n * 3, n, x1);
        msort(raw + ((size_t)c * 2 + 1) * n * 3, n, x2);
        int *asg = malloc(4 * n), *inv = malloc(4 * n), *bid = malloc(4 * n), *maxidx = malloc(4 * n), *un = malloc(4 * n);
        float *price = calloc(n, 4), *inc = malloc(4 * n), *maxinc = calloc(n, 4), *T2 = malloc(4 * n), *tms = malloc(4 * n), *tmh = malloc(4 * n);
        int(*seed)[4] = malloc(16 * n);
        for (int j = 0; j < n; j++) { asg[j] = inv[j] = -1; for (int s = 0; s < 4; s++) seed[j][s] = -1; }
        Box *bx = malloc(sizeof(Box) * NT), *ba = malloc(sizeof(Box) * NT), *bf = malloc(sizeof(Box) * NT);
        int *freel = malloc(4 * n);
        for (int t = 0; t < iters; t++) {
            int u = 0;
            for (int j = 0; j < n; j++) if (asg[j] < 0) un[u++] = j;
            if (!u) break;
            alg += (double)u * n;
            /* boxes: scheme 0 all targets; scheme 1 assigned-only tiles + free tiles */
            int nf = 0;
            for (int k = 0; k < n; k++) if (inv[k] < 0) freel[nf++] = k;
            const int NF = (nf + TILE - 1) / TILE;
            for (int tl = 0; tl < NT; tl++) {
                Box *b0 = &bx[tl], *b1 = &ba[tl];
                for (int i = 0; i < 3; i++) { b0->lo[i] = b1->lo[i] = 3e38f; b0->hi[i] = b1->hi[i] = -3e38f; }
                b0->cmax = b1->cmax = -1e30f; b0->n = b1->n = 0;
                for (int k = tl * TILE; k < n && k < (tl + 1) * TILE; k++) {
                    for (int pass = 0; pass < 2; pass++) {
                        if (pass == 1 && inv[k] < 0) continue;
                        Box *b = pass ? b1 : b0;
                        for (int i = 0; i < 3; i++) { b->lo[i] = fminf(b->lo[i], x2[3 * k + i]); b->hi[i] = fmaxf(b->hi[i], x2[3 * k + i]); }
                        b->cmax = fmaxf(b->cmax, 3.f - price[k]); b->n++;
                    }
                }
            }
            for (int tl = 0; tl < NF; tl++) {
                Box *b = &bf[tl];
                for (int i = 0; i < 3; i++) { b->lo[i] = 3e38f; b->hi[i] = -3e38f; }
                b->cmax = 3.f; b->n = 0;
                for (int q = tl * TILE; q < nf && q < (tl + 1) * TILE; q++) {
                    const int k = freel[q];
                    for (int i = 0; i < 3; i++) { b->lo[i] = fminf(b->lo[i], x2[3 * k + i]); b->hi[i] = fmaxf(b->hi[i], x2[3 * k + i]); }
                    b->n++;
                }
            }
            /* bids */
            for (int q = 0; q < u; q++) {
                const int j = un[q];
                const float *a = x1 + 3 * j;
                /* first-iteration seeds: Morton rank neighbours */
                if (seed[j][0] < 0 && n >= 4) { int k1 = j < 1 ? 1 : j > n - 3 ? n - 3 : j; seed[j][0] = k1; seed[j][1] = k1 - 1; seed[j][2] = k1 + 1; seed[j][3] = k1 + 2; }
                float sv[4]; int ns = 0;
                for (int s = 0; s < 4; s++) {
                    const int k = seed[j][s]; int dup = k < 0;
                    for (int s2 = 0; s2 < s; s2++) dup |= seed[j][s2] == k;
                    if (dup) continue;
                    const float dx = x2[3 * k] - a[0], dy = x2[3 * k + 1] - a[1], dz = x2[3 * k + 2] - a[2];
                    sv[ns++] = 3.f - sqrtf(dx * dx + dy * dy + dz * dz) - price[k];
                }
                float hi = -1e9f, lo = -1e9f;
                for (int s = 0; s < ns; s++) { if (sv[s] > hi) { lo = hi; hi = sv[s]; } else if (sv[s] > lo) lo = sv[s]; }
                tms[j] = lo - 2e-6f;
                {   /* threshold after a pre-pass over the tile of the previous best object (or of the own Morton rank) */
                    float h1 = hi, h2 = lo;
                    const int ht = (seed[j][0] >= 0 ? seed[j][0] : j) / TILE;
                    for (int k = ht * TILE; k < n && k < (ht + 1) * TILE; k++) {
                        int dup = 0;
                        for (int s2 = 0; s2 < 4; s2++) dup |= seed[j][s2] == k;
                        if (dup) continue;
                        const float dx = x2[3 * k] - a[0], dy = x2[3 * k + 1] - a[1], dz = x2[3 * k + 2] - a[2];
                        const float v = 3.f - sqrtf(dx * dx + dy * dy + dz * dz) - price[k];
                        if (v > h1) { h2 = h1; h1 = v; } else if (v > h2) h2 = v;
                    }
                    tmh[j] = h2 - 2e-6f;
                }
                float v4[4] = {-1e9f, -1e9f, -1e9f, -1e9f}; int k4[4] = {-1, -1, -1, -1};
                for (int k = 0; k < n; k++) {
                    const float dx = x2[3 * k] - a[0], dy = x2[3 * k + 1] - a[1], dz = x2[3 * k + 2] - a[2];
                    const float v = (float)(3.0 - (double)sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy))) - (double)price[k]);
                    int p = 4;
                    while (p > 0 && v > v4[p - 1]) p--;
                    if (p < 4) { for (int s = 3; s > p; s--) { v4[s] = v4[s - 1]; k4[s] = k4[s - 1]; } v4[p] = v; k4[p] = k; }
                    if (v >= tms[j]) exact_needed += 1;
                    if (v >= tmh[j]) exact_home += 1;
                }
                bid[j] = k4[0]; inc[j] = v4[0] - v4[1] + eps; T2[j] = v4[1] - 2e-6f;
                for (int s = 0; s < 4; s++) seed[j][s] = k4[s];
                if (inc[j] > maxinc[k4[0]]) maxinc[k4[0]] = inc[j];
            }
            /* survival counts */
            double it_cnt[2][2][4] = {{{0}}};
            for (int gi = 0; gi < 4; gi++) {
                const int G = Gs[gi];
                for (int q0 = 0; q0 < u; q0 += G) {
                    const int q1 = q0 + G < u ? q0 + G : u;
                    for (int thr = 0; thr < 2; thr++) {
                        const float *th = thr ? T2 : tms;
                        for (int tl = 0; tl < NT; tl++) {
                            int s0 = 1, s1 = 1;
                            for (int q = q0; q < q1 && (s0 || s1); q++) {
                                const int j = un[q];
                                if (s0 && !skippable(&bx[tl], x1 + 3 * j, th[j])) s0 = 0;
                                if (s1 && !skippable(&ba[tl], x1 + 3 * j, th[j])) s1 = 0;
                            }
                            if (!s0) it_cnt[0][thr][gi] += (double)(q1 - q0) * TILE;
                            if (!s1) it_cnt[1][thr][gi] += (double)(q1 - q0) * TILE;
                        }
                        for (int tl = 0; tl < NF; tl++) {
                            int s1 = 1;
                            for (int q = q0; q < q1 && s1; q++) if (!skippable(&bf[tl], x1 + 3 * un[q], th[un[q]])) s1 = 0;
                            if (!s1) it_cnt[1][thr][gi] += (double)(q1 - q0) * TILE;
                        }
                    }
                }
            }
            for (int q0 = 0; q0 < u; q0 += 32) {
                const int q1 = q0 + 32 < u ? q0 + 32 : u;
                for (int tl = 0; tl < NT; tl++) {
                    int s0 = 1;
                    for (int q = q0; q < q1 && s0; q++) if (!skippable(&bx[tl], x1 + 3 * un[q], tmh[un[q]])) s0 = 0;
                    if (!s0) home_tiles += (double)(q1 - q0) * TILE;
                }
            }
            for (int s = 0; s < 2; s++) for (int thr = 0; thr < 2; thr++) for (int gi = 0; gi < 4; gi++) tot[s][thr][gi] += it_cnt[s][thr][gi];
            if (verbose && c == 0) {
                float pm = 0; for (int k = 0; k < n; k++) pm = fmaxf(pm, price[k]);
                printf("it %2d U %4d free %4d pmax %.3f | frac of U*N executed: s0 seed", t, u, nf, pm);
                for (int gi = 0; gi < 4; gi++) printf(" %.3f", it_cnt[0][0][gi] / ((double)u * n));
                printf(" | s0 final"); for (int gi = 0; gi < 4; gi++) printf(" %.3f", it_cnt[0][1][gi] / ((double)u * n));
                printf(" | s1 seed"); for (int gi = 0; gi < 4; gi++) printf(" %.3f", it_cnt[1][0][gi] / ((double)u * n));
                printf(" | s1 final"); for (int gi = 0; gi < 4; gi++) printf(" %.3f", it_cnt[1][1][gi] / ((double)u * n));
                printf("\n");
            }
            /* GetMax + Assign */
            for (int q = 0; q < u; q++) { const int j = un[q], o = bid[j]; if ((double)inc[j] - 1e-6 <= (double)maxinc[o] && (double)maxinc[o] <= (double)inc[j] + 1e-6) maxidx[o] = j; }
            for (int q = 0; q < u; q++) {
                const int j = un[q], o = bid[j];
                if (t == iters - 1 || maxidx[o] == j) {
                    const int prev = inv[o];
                    if (t != iters - 1 && prev != -1) asg[prev] = -1;
                    inv[o] = j; asg[j] = o; price[o] += inc[j]; maxinc[o] = -1e9f;
                }
            }
        }
    }
    printf("algorithmic pairs %.4g; candidates above the seed threshold (exact evaluations needed) %.4g = %.4f\n", alg, exact_needed, exact_needed / alg);
    printf("with a pre-pass over the tile of the previous best object: candidates above the threshold %.4f of the pairs, executed/algorithmic (G=32) %.4f\n", exact_home / alg, home_tiles / alg);
    for (int s = 0; s < 2; s++)
        for (int thr = 0; thr < 2; thr++) {
            printf("scheme %d (%s) threshold %-5s: executed/algorithmic for G=32,8,4,1:", s, s ? "free objects in own tiles" : "shipped", thr ? "final" : "seed");
            for (int gi = 0; gi < 4; gi++) printf(" %.4f", tot[s][thr][gi] / alg);
            printf("\n");
        }
    return 0;
}
