/*
 * oracle/fps_oracle.c -- CPU restatement of farthest point sampling.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference calls third-party CUDA code that is neither under /root/reference nor
 * installed: pointnet2_ops._ext.furthest_point_sampling (models/pointnet2_utils.py:6,89-90; unpinned git URL in
 * requirements.txt:8) and pytorch3d.ops.sample_farthest_points (utils.py:10,90; models/pointmlp.py:158).
 * What IS in the reference is the torch algorithm kept as a comment in models/pointnet2_utils.py:64-86, which this
 * file follows: distance = 1e10; loop { centroids[i] = farthest; dist = sum((xyz - centroid)**2, -1);
 * distance = min(distance, dist); farthest = argmax(distance) }, with the start index 0 of pointnet2_ops /
 * pytorch3d (random_start_point=False) instead of the comment's torch.randint, and the first maximum on ties.
 * skip_origin != 0 reproduces pointnet2_ops' quirk of never selecting points with x*x+y*y+z*z <= 1e-3 (recalled from
 * its sampling_gpu.cu; cannot be verified here).
 */
#include <stdint.h>
#include <stdlib.h>

int fps_oracle(const float *xyz, long bs, long rs, int b, int n, int npoint, const int *start, int skip_origin, int *idx) {
    if (b < 0 || n < 1 || npoint < 0) return -1;
    float *mind = (float *)malloc(sizeof(float) * (size_t)n);
    if (!mind) return -2;
    for (int i = 0; i < b; i++) {
        const float *p = xyz + i * bs;
        int *out = idx + (size_t)i * npoint;
        for (int k = 0; k < n; k++) mind[k] = 1e10f;
        int last = start ? start[i] : 0;
        for (int j = 0; j < npoint; j++) {
            out[j] = last;
            const float lx = p[last * rs + 0], ly = p[last * rs + 1], lz = p[last * rs + 2];
            float best = -1.f;
            int besti = 0;
            for (int k = 0; k < n; k++) {
                const float x = p[k * rs + 0], y = p[k * rs + 1], z = p[k * rs + 2];
                if (skip_origin && ((x * x + y * y) + z * z) <= 1e-3f) continue;
                const float dx = x - lx, dy = y - ly, dz = z - lz;
                const float d = (dx * dx + dy * dy) + dz * dz;
                const float m = d < mind[k] ? d : mind[k];
                mind[k] = m;
                if (m > best) { best = m; besti = k; }
            }
            last = besti;
        }
    }
    free(mind);
    return 0;
}
