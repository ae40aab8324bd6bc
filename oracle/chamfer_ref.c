/* ORACLE (test infrastructure, not product code): CPU restatement of the chamfer3D forward/backward.
 *
 * Follows /root/reference/external/chamfer3D/chamfer3D.cu:12-134 (NmDistanceKernel) and :155-174
 * (NmDistanceGradKernel). Bit-level definition taken from the sm_100a SASS of that file under
 * nvcc 12.9: with (x,y,z) = candidate - query,
 *      d = fmaf(z, z, fmaf(x, x, y*y))
 * candidates are scanned in 512-wide tiles; inside a tile the first candidate initialises the running
 * best and a later one replaces it iff d < best (strict); the tile winner replaces the stored result
 * iff it is the first tile or stored > winner (strict). For finite inputs this is "minimum distance,
 * lowest candidate index among exact ties".
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (see oracle/Makefile). -ffp-contract=off keeps
 * y*y a separately rounded product; the two fmaf() calls are explicit.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define SC_TILE 512

static int g_threads = 0; /* 0 = all online cores (libgomp is not in the image, so plain pthreads) */
void sc_oracle_set_threads(int t) { g_threads = t; }
int sc_oracle_get_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct { int b, n, m; const float *q, *c; float *dist; int32_t *idx; long lo, hi; } job_t;

static void *direction_worker(void *arg)
{
    const job_t *J = (const job_t *)arg;
    const int n = J->n, m = J->m;
    const float *q = J->q, *c = J->c;
    float *dist = J->dist;
    int32_t *idx = J->idx;
    for (long w = J->lo; w < J->hi; ++w) {
        {
            const int i = (int)(w / n), j = (int)(w % n);
            const float qx = q[((long)i * n + j) * 3 + 0];
            const float qy = q[((long)i * n + j) * 3 + 1];
            const float qz = q[((long)i * n + j) * 3 + 2];
            float stored = 0.0f;
            int32_t stored_i = 0;
            for (int k0 = 0; k0 < m; k0 += SC_TILE) {
                const int k1 = (k0 + SC_TILE < m) ? k0 + SC_TILE : m;
                float best = 0.0f;
                int32_t best_i = 0;
                for (int k = k0; k < k1; ++k) {
                    const float x = c[((long)i * m + k) * 3 + 0] - qx;
                    const float y = c[((long)i * m + k) * 3 + 1] - qy;
                    const float z = c[((long)i * m + k) * 3 + 2] - qz;
                    const float d = fmaf(z, z, fmaf(x, x, y * y));
                    if (k == k0 || d < best) { best = d; best_i = k; }
                }
                if (k0 == 0 || stored > best) { stored = best; stored_i = best_i; }
            }
            if (m > 0) { dist[(long)i * n + j] = stored; idx[(long)i * n + j] = stored_i; }
        }
    }
    return 0;
}

static void one_direction(int b, int n, const float *q, int m, const float *c, float *dist, int32_t *idx)
{
    const long total = (long)b * n;
    int T = sc_oracle_get_threads();
    if (T > total) T = total > 0 ? (int)total : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * T);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * T);
    for (int t = 0; t < T; ++t) {
        job_t J = { b, n, m, q, c, dist, idx, total * t / T, total * (t + 1) / T };
        jobs[t] = J;
        pthread_create(&th[t], 0, direction_worker, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
    free(th);
    free(jobs);
}

/* xyz1 [b,n,3], xyz2 [b,m,3] -> dist1/idx1 [b,n], dist2/idx2 [b,m]. Returns 1 like the reference's binding. */
int sc_oracle_chamfer_forward(const float *xyz1, const float *xyz2, int b, int n, int m,
                              float *dist1, float *dist2, int32_t *idx1, int32_t *idx2)
{
    one_direction(b, n, xyz1, m, xyz2, dist1, idx1);
    one_direction(b, m, xyz2, n, xyz1, dist2, idx2);
    return 1;
}

static void grad_direction(int b, int n, const float *p1, int m, const float *p2, const float *gd,
                           const int32_t *idx, float *g1, float *g2)
{
    /* sequential on purpose: the reference uses float atomics, so its summation order is unspecified;
       the parity test for backward uses a tolerance. */
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const long a = ((long)i * n + j) * 3;
            const long t = ((long)i * m + idx[(long)i * n + j]) * 3;
            const float g = gd[(long)i * n + j] * 2;
            for (int c = 0; c < 3; ++c) {
                const float v = g * (p1[a + c] - p2[t + c]);
                g1[a + c] += v;
                g2[t + c] += -v;
            }
        }
}

/* gradxyz1/gradxyz2 must be zeroed by the caller (as in the reference). */
int sc_oracle_chamfer_backward(const float *xyz1, const float *xyz2, int b, int n, int m,
                               const float *graddist1, const float *graddist2,
                               const int32_t *idx1, const int32_t *idx2, float *gradxyz1, float *gradxyz2)
{
    grad_direction(b, n, xyz1, m, xyz2, graddist1, idx1, gradxyz1, gradxyz2);
    grad_direction(b, m, xyz2, n, xyz1, graddist2, idx2, gradxyz2, gradxyz1);
    return 1;
}
