/*
 * rt3d_oracle.c -- CPU restatement of the reference's per-frame geometric path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product (librt3d.so) never links it.
 *
 * Each function follows the reference lines cited above it (paths relative to
 * /root/reference).  Parity status:
 *   - B0/B1/T1/C1/W0/V1/S1/centroid: pinned against the reference's own source,
 *     imported unmodified on torch-CPU (oracle/ref_import.py, tests/golden/).
 *   - O1 (statistical outlier removal) and V2 (Open3D tensor voxel_down_sample)
 *     restate the published algorithm of the un-vendored dependency
 *     open3d==0.18.0 (requirements.txt:64): PARITY UNPINNED for those two --
 *     Open3D cannot be installed here; O1 is cross-checked against an
 *     independent scipy cKDTree restatement only.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off  (see oracle/Makefile).
 * -ffp-contract=off matters: every fused multiply-add below is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RT3D_ORACLE_VERSION 1

int rt3d_oracle_version(void) { return RT3D_ORACLE_VERSION; }

void rt3d_oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

int rt3d_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* arithmetic flavour of `tensor / python_float`:
 *   0 = IEEE true division by fl32(b) (torch-CPU, NumPy)
 *   1 = multiply by float(1.0 / double(b)) (torch-CUDA's div-by-scalar kernel; measured
 *       against torch 2.11 on a B200 -- the reciprocal is taken in double, rounded once) */
static inline float div_flavour(float a, float b, float inv_b, int recip) {
    return recip ? a * inv_b : a / b;
}

/* ------------------------------------------------------------------ B0 ----
 * torch.nonzero(mask, as_tuple=False)   2cam/vision_pipeline_utils.py:296
 * mask: (H,W) plane, dtype 0 = uint8, 1 = float32.  Row-major [row, col].   */
int64_t rt3d_oracle_mask_nonzero(const void *mask, int dtype, int H, int W, int64_t *out_idx) {
    int64_t n = 0;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            size_t i = (size_t)r * W + c;
            int nz = dtype == 0 ? (((const uint8_t *)mask)[i] != 0)
                                : (((const float *)mask)[i] != 0.0f); /* NaN != 0 is true, as in torch */
            if (nz) {
                if (out_idx) { out_idx[2 * n] = r; out_idx[2 * n + 1] = c; }
                ++n;
            }
        }
    return n;
}

/* ------------------------------------------------------------------ B1 ----
 * convert_mask_to_3d_points   2cam/vision_pipeline_utils.py:19-36
 * idx (K,2) int64 [row, col]; keeps d > 0 and finite (:26); x,y,z per :32-34. */
int64_t rt3d_oracle_backproject_indices(const int64_t *idx, int64_t K, const float *depth, int H, int W,
                                        double cx_, double cy_, double fx_, double fy_, int recip, float *out_xyz) {
    (void)H;
    const float cx = (float)cx_, cy = (float)cy_, fx = (float)fx_, fy = (float)fy_;
    const float inv_fx = (float)(1.0 / fx_), inv_fy = (float)(1.0 / fy_);
    int64_t n = 0;
    for (int64_t i = 0; i < K; ++i) {
        int64_t v = idx[2 * i], u = idx[2 * i + 1];
        float d = depth[(size_t)v * W + u];
        if (!(d > 0.0f) || isnan(d) || isinf(d)) continue;
        float x = div_flavour(((float)u - cx) * d, fx, inv_fx, recip);
        float y = div_flavour(((float)v - cy) * d, fy, inv_fy, recip);
        out_xyz[3 * n] = x; out_xyz[3 * n + 1] = y; out_xyz[3 * n + 2] = d;
        ++n;
    }
    return n;
}

/* ------------------------------------------------------------------ T1 ----
 * torch.mm(P, R.T) + t  (:301)  ==  torch.mm(R, P.T).T + t  (:239-240)
 * torch-CPU (MKL sgemm, rows >= 11) rounds as  fma(z,R2, fma(y,R1, x*R0)) then a
 * separately rounded +t  [measured here against the reference].              */
void rt3d_oracle_transform(const float *P, int64_t n, const float *R, const float *t, float *out) {
    for (int64_t i = 0; i < n; ++i) {
        float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
        for (int j = 0; j < 3; ++j) {
            float acc = x * R[3 * j];
            acc = fmaf(y, R[3 * j + 1], acc);
            acc = fmaf(z, R[3 * j + 2], acc);
            out[3 * i + j] = acc + t[j];
        }
    }
}

/* T1d: float64 variant  np.dot(R64, P.T).T + t64   2cam/2cams_mask_cpu.py:595,618
 * (order of the three products inside dgemm is BLAS-specific; compared with a
 * 1e-12 tolerance, never bit-exact).                                         */
void rt3d_oracle_transform_f64(const float *P, int64_t n, const double *R, const double *t, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        double x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
        for (int j = 0; j < 3; ++j) {
            double acc = x * R[3 * j];
            acc = fma(y, R[3 * j + 1], acc);
            acc = fma(z, R[3 * j + 2], acc);
            out[3 * i + j] = acc + t[j];
        }
    }
}

/* ------------------------------------------------------------------ C1 ----
 * crop_point_cloud_gpu   2cam/vision_pipeline_utils.py:78-84
 * closed intervals, bounds compared as fp32 scalars, order preserved.
 * box = {xlo,xhi,ylo,yhi,zlo,zhi}.                                           */
int64_t rt3d_oracle_crop(const float *P, int64_t n, const float *box, float *out) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = P + 3 * i;
        if (p[0] >= box[0] && p[0] <= box[1] && p[1] >= box[2] && p[1] <= box[3] &&
            p[2] >= box[4] && p[2] <= box[5]) {
            out[3 * m] = p[0]; out[3 * m + 1] = p[1]; out[3 * m + 2] = p[2];
            ++m;
        }
    }
    return m;
}

/* ------------------------------------------------------------------ W0 ----
 * workspace ingest   2cam/vision_pipeline_utils.py:233-238
 * rows of an (n, stride) fp32 image (stride 4 = ZED XYZ + padding channel);
 * keeps rows whose first three channels are all finite.                      */
int64_t rt3d_oracle_xyz_ingest(const float *xyz, int64_t n, int stride, float *out) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = xyz + (size_t)stride * i;
        if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2])) {
            out[3 * m] = p[0]; out[3 * m + 1] = p[1]; out[3 * m + 2] = p[2];
            ++m;
        }
    }
    return m;
}

/* ------------------------------------------------------------------ V1 ----
 * downsample_point_cloud_gpu   2cam/vision_pipeline_utils.py:40-44
 * q = rint(p / v) * v (round half to even), then torch.unique(dim=0):
 * lexicographically sorted distinct rows (x major).  -0.0 == +0.0.           */
static int cmp_row3f(const void *a, const void *b) {
    const float *p = (const float *)a, *q = (const float *)b;
    for (int k = 0; k < 3; ++k) {
        if (p[k] < q[k]) return -1;
        if (p[k] > q[k]) return 1;
    }
    return 0;
}

void rt3d_oracle_voxel_snap(const float *P, int64_t n, double v_, int recip, float *out) {
    const float v = (float)v_, inv_v = (float)(1.0 / v_);
    for (int64_t i = 0; i < 3 * n; ++i) out[i] = rintf(div_flavour(P[i], v, inv_v, recip)) * v;
}

int64_t rt3d_oracle_voxel_snap_unique(const float *P, int64_t n, double v, int recip, float *out) {
    if (n == 0) return 0;
    float *q = (float *)malloc(sizeof(float) * 3 * (size_t)n);
    rt3d_oracle_voxel_snap(P, n, v, recip, q);
    qsort(q, (size_t)n, 3 * sizeof(float), cmp_row3f);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (i == 0 || cmp_row3f(q + 3 * i, q + 3 * (i - 1)) != 0) {
            memcpy(out + 3 * m, q + 3 * i, 3 * sizeof(float));
            ++m;
        }
    }
    free(q);
    return m;
}

/* integer voxel indices rint(p / v) as int32 triples (what the packed keys hold) */
void rt3d_oracle_voxel_index(const float *P, int64_t n, double v_, int recip, int32_t *out) {
    const float v = (float)v_, inv_v = (float)(1.0 / v_);
    for (int64_t i = 0; i < 3 * n; ++i) out[i] = (int32_t)rintf(div_flavour(P[i], v, inv_v, recip));
}

/* ------------------------------------------------------------------ V2 ----
 * downsample_point_cloud   2cam/2cams_mask_cpu.py:60-65
 * -> open3d 0.18.0 t::geometry::PointCloud::VoxelDownSample(voxel, "mean")
 * [published algorithm; un-vendored]:  key = floor(p / v) in the input dtype,
 * positions cast to Float32 and summed per key (IndexAdd_), divided by the
 * Float32 count, cast back to the input dtype.  Open3D's row order is the hash
 * buffer order (arbitrary); we emit voxels sorted by (kx,ky,kz) and accumulate
 * in input order.  dtype: 0 = float32 input, 1 = float64 input.
 * out_keys (m,3) int64, out_xyz (m,3) double, out_cnt (m) int32.             */
typedef struct { int64_t k[3]; int64_t i; } v2_item;
static int cmp_v2(const void *a, const void *b) {
    const v2_item *p = (const v2_item *)a, *q = (const v2_item *)b;
    for (int d = 0; d < 3; ++d) {
        if (p->k[d] < q->k[d]) return -1;
        if (p->k[d] > q->k[d]) return 1;
    }
    return (p->i > q->i) - (p->i < q->i);
}

int64_t rt3d_oracle_voxel_centroid(const void *P, int dtype, int64_t n, double voxel, int64_t *out_keys,
                                   double *out_xyz, int32_t *out_cnt) {
    if (n == 0) return 0;
    v2_item *it = (v2_item *)malloc(sizeof(v2_item) * (size_t)n);
    const float v32 = (float)voxel;
    for (int64_t i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) {
            if (dtype == 0) it[i].k[d] = (int64_t)floorf(((const float *)P)[3 * i + d] / v32);
            else            it[i].k[d] = (int64_t)floor(((const double *)P)[3 * i + d] / voxel);
        }
        it[i].i = i;
    }
    qsort(it, (size_t)n, sizeof(v2_item), cmp_v2);
    int64_t m = 0, s = 0;
    while (s < n) {
        int64_t e = s;
        float acc[3] = {0.f, 0.f, 0.f};
        while (e < n && it[e].k[0] == it[s].k[0] && it[e].k[1] == it[s].k[1] && it[e].k[2] == it[s].k[2]) {
            for (int d = 0; d < 3; ++d) {
                float p = dtype == 0 ? ((const float *)P)[3 * it[e].i + d] : (float)((const double *)P)[3 * it[e].i + d];
                acc[d] = acc[d] + p;
            }
            ++e;
        }
        float cnt = (float)(e - s);
        for (int d = 0; d < 3; ++d) {
            out_keys[3 * m + d] = it[s].k[d];
            out_xyz[3 * m + d] = (double)(acc[d] / cnt);
        }
        if (out_cnt) out_cnt[m] = (int32_t)(e - s);
        ++m;
        s = e;
    }
    free(it);
    return m;
}

/* ------------------------------------------------------------ centroid ----
 * calculate_centroid = np.mean(pc, axis=0)   2cam/vision_pipeline_utils.py:74-75
 * NumPy reduces axis 0 of a C-contiguous (N,3) array by sequential row adds in
 * the array's dtype, then divides by N  [measured here].                     */
void rt3d_oracle_centroid_f32(const float *P, int64_t n, float *out) {
    float s[3] = {0.f, 0.f, 0.f};
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) s[d] = s[d] + P[3 * i + d];
    for (int d = 0; d < 3; ++d) out[d] = s[d] / (float)n;
}
void rt3d_oracle_centroid_f64(const double *P, int64_t n, double *out) {
    double s[3] = {0., 0., 0.};
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) s[d] = s[d] + P[3 * i + d];
    for (int d = 0; d < 3; ++d) out[d] = s[d] / (double)n;
}

/* ------------------------------------------------------------------ O1 ----
 * filter_outliers_sor   2cam/vision_pipeline_utils.py:48-58
 * -> open3d 0.18.0 geometry::PointCloud::RemoveStatisticalOutliers
 * [published algorithm; un-vendored], all in fp64:
 *   kNN(k = nb_neighbors, query included) with nanoflann L2 metric
 *       d2 = ((dx*dx + dy*dy) + dz*dz),  results ascending;
 *   avg_i = (sum_j sqrt(d2_ij), ascending, from 0.0) / k'   (k' = min(k, N));
 *   valid = #points with a non-empty result (= N for N >= 1);
 *   mu    = (sum over avg_i > 0, index order, from 0.0) / valid;
 *   sigma = sqrt( sum_{avg_i>0} (avg_i-mu)*(avg_i-mu) / (valid-1) );
 *   keep  = avg_i > 0 && avg_i < mu + std_ratio*sigma.
 * The kNN here is exact (the k smallest d2 values); nanoflann's branch
 * pruning can differ on last-ulp near-ties -- part of "parity unpinned".     */
typedef struct {
    double lo[3];
    double h, inv_h;
    int dim[3];
    int64_t *cell_start; /* ncell+1 */
    int64_t *order;      /* n, point ids grouped by cell */
} grid_t;

static inline void heap_push(double *heap, int *sz, int k, double v) {
    /* max-heap of the k smallest values */
    if (*sz < k) {
        int i = (*sz)++;
        heap[i] = v;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (heap[p] < heap[i]) { double t = heap[p]; heap[p] = heap[i]; heap[i] = t; i = p; } else break;
        }
    } else if (v < heap[0]) {
        heap[0] = v;
        int i = 0;
        for (;;) {
            int l = 2 * i + 1, r = l + 1, m = i;
            if (l < k && heap[l] > heap[m]) m = l;
            if (r < k && heap[r] > heap[m]) m = r;
            if (m == i) break;
            double t = heap[m]; heap[m] = heap[i]; heap[i] = t; i = m;
        }
    }
}

static int cmp_dbl(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

static inline double dist2(const double *a, const double *b) {
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
}

static double avg_from_heap(double *heap, int sz) {
    qsort(heap, (size_t)sz, sizeof(double), cmp_dbl);
    double s = 0.0;
    for (int j = 0; j < sz; ++j) s = s + sqrt(heap[j]);
    return s / (double)sz;
}

static void grid_build(grid_t *g, const double *P, int64_t n, double h_hint) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            if (P[3 * i + d] < lo[d]) lo[d] = P[3 * i + d];
            if (P[3 * i + d] > hi[d]) hi[d] = P[3 * i + d];
        }
    double ext[3], emax = 0;
    for (int d = 0; d < 3; ++d) { ext[d] = hi[d] - lo[d]; if (ext[d] > emax) emax = ext[d]; }
    double h = h_hint;
    if (!(h > 0)) {
        /* surface-like clouds: spacing ~ sqrt(area / n); cell = 3 spacings */
        double e0 = ext[0], e1 = ext[1], e2 = ext[2], t;
        if (e0 < e1) { t = e0; e0 = e1; e1 = t; }
        if (e1 < e2) { t = e1; e1 = e2; e2 = t; }
        if (e0 < e1) { t = e0; e0 = e1; e1 = t; }
        double area = e0 * (e1 > 0 ? e1 : e0);
        h = 3.0 * sqrt(area / (double)(n > 0 ? n : 1));
    }
    if (!(h > 0)) h = 1.0;
    for (;;) { /* cap the dense table at 2^25 cells */
        double cells = 1;
        for (int d = 0; d < 3; ++d) cells *= floor(ext[d] / h) + 1;
        if (cells <= (double)(1 << 25)) break;
        h *= 1.5;
    }
    g->h = h; g->inv_h = 1.0 / h;
    int64_t ncell = 1;
    for (int d = 0; d < 3; ++d) { g->lo[d] = lo[d]; g->dim[d] = (int)floor(ext[d] / h) + 1; ncell *= g->dim[d]; }
    g->cell_start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
    g->order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t *cid = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) {
        int c[3];
        for (int d = 0; d < 3; ++d) {
            c[d] = (int)floor((P[3 * i + d] - lo[d]) * g->inv_h);
            if (c[d] < 0) c[d] = 0;
            if (c[d] >= g->dim[d]) c[d] = g->dim[d] - 1;
        }
        cid[i] = ((int64_t)c[0] * g->dim[1] + c[1]) * g->dim[2] + c[2];
        g->cell_start[cid[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; ++c) g->cell_start[c + 1] += g->cell_start[c];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell);
    memcpy(fill, g->cell_start, sizeof(int64_t) * (size_t)ncell);
    for (int64_t i = 0; i < n; ++i) g->order[fill[cid[i]]++] = i;
    free(fill); free(cid);
}

static void grid_free(grid_t *g) { free(g->cell_start); free(g->order); }

static inline void grid_cell_of(const grid_t *g, const double *p, int *c) {
    for (int d = 0; d < 3; ++d) {
        c[d] = (int)floor((p[d] - g->lo[d]) * g->inv_h);
        if (c[d] < 0) c[d] = 0;
        if (c[d] >= g->dim[d]) c[d] = g->dim[d] - 1;
    }
}

/* average distance to the k nearest neighbours (self included), exact */
void rt3d_oracle_knn_avg(const double *P, int64_t n, int k, int brute, double *avg) {
    if (n == 0) return;
    int kk = (int64_t)k < n ? k : (int)n;
    if (brute) {
#pragma omp parallel
        {
            double *heap = (double *)malloc(sizeof(double) * (size_t)kk);
#pragma omp for schedule(dynamic, 64)
            for (int64_t i = 0; i < n; ++i) {
                int sz = 0;
                for (int64_t j = 0; j < n; ++j) heap_push(heap, &sz, kk, dist2(P + 3 * i, P + 3 * j));
                avg[i] = avg_from_heap(heap, sz);
            }
            free(heap);
        }
        return;
    }
    grid_t g;
    grid_build(&g, P, n, 0.0);
    int maxdim = g.dim[0] > g.dim[1] ? g.dim[0] : g.dim[1];
    if (g.dim[2] > maxdim) maxdim = g.dim[2];
#pragma omp parallel
    {
        double *heap = (double *)malloc(sizeof(double) * (size_t)kk);
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            const double *q = P + 3 * i;
            int c[3], sz = 0;
            grid_cell_of(&g, q, c);
            for (int r = 0; r <= maxdim; ++r) {
                /* visit the shell of cells at Chebyshev distance exactly r */
                for (int x = c[0] - r; x <= c[0] + r; ++x) {
                    if (x < 0 || x >= g.dim[0]) continue;
                    int ax = abs(x - c[0]);
                    for (int y = c[1] - r; y <= c[1] + r; ++y) {
                        if (y < 0 || y >= g.dim[1]) continue;
                        int ay = abs(y - c[1]);
                        for (int z = c[2] - r; z <= c[2] + r; ++z) {
                            if (z < 0 || z >= g.dim[2]) continue;
                            if (ax != r && ay != r && abs(z - c[2]) != r) continue; /* interior: done in earlier rings */
                            int64_t cell = ((int64_t)x * g.dim[1] + y) * g.dim[2] + z;
                            for (int64_t s = g.cell_start[cell]; s < g.cell_start[cell + 1]; ++s)
                                heap_push(heap, &sz, kk, dist2(q, P + 3 * g.order[s]));
                        }
                    }
                }
                /* every unvisited point is farther than r*h (minus rounding slack) */
                if (sz == kk) {
                    double reach = (double)r * g.h * (1.0 - 1e-9);
                    if (heap[0] <= reach * reach) break;
                }
            }
            avg[i] = avg_from_heap(heap, sz);
        }
        free(heap);
    }
    grid_free(&g);
}

/* stats[0]=mu, stats[1]=sigma, stats[2]=threshold, stats[3]=valid.  Returns #kept. */
int64_t rt3d_oracle_sor_threshold(const double *avg, int64_t n, double std_ratio, uint8_t *keep, double *stats) {
    if (n == 0) return 0;
    int64_t valid = n; /* every query finds at least itself */
    double mu = 0.0;
    for (int64_t i = 0; i < n; ++i) if (avg[i] > 0) mu = mu + avg[i];
    mu /= (double)valid;
    double sq = 0.0;
    for (int64_t i = 0; i < n; ++i) sq = sq + (avg[i] > 0 ? (avg[i] - mu) * (avg[i] - mu) : 0.0);
    double sigma = sqrt(sq / (double)(valid - 1));
    double thr = mu + std_ratio * sigma;
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) {
        keep[i] = (avg[i] > 0 && avg[i] < thr) ? 1 : 0;
        kept += keep[i];
    }
    if (stats) { stats[0] = mu; stats[1] = sigma; stats[2] = thr; stats[3] = (double)valid; }
    return kept;
}

/* ---- O1, third opinion: the k-NN search as nanoflann performs it ------------
 * Open3D 0.18.0 builds `nanoflann::KDTreeEigenMatrixAdaptor<..., metric_L2>` with leaf_max_size 15
 * (cpp/open3d/geometry/KDTreeFlann.cpp, SetRawData: `KDTreeSingleIndexAdaptorParams(15)`) and answers SearchKNN with
 * `index_->knnSearch` (nanoflann.hpp, bundled v1.5.x).  That source is not available offline; what follows restates
 * its published algorithm from memory, function for function:
 *   buildIndex / divideTree   recursive split until a node holds <= 15 points; the node box is re-tightened from its
 *                             children on the way up
 *   middleSplit_              cut dimension = largest point spread among the dimensions whose box span is within
 *                             (1 - 1e-5) of the largest span; cut value = box middle clamped to the points' range;
 *                             planeSplit's three-way partition (<, ==, >) and the lim1 / lim2 / count/2 balance rule
 *   findNeighbors             computeInitialDistances (per-dimension distance of the query to the root box), then
 *   searchLevel               leaf: worst distance read ONCE before the loop, `dist < worst` admits a point;
 *                             inner node: nearer child first, the farther one only if
 *                             mindistsq + cut_dist - dists[idx] <= worstDist()   (eps = 0)
 *   KNNResultSet::addPoint    insertion into the ascending list (no index tie-break: NANOFLANN_FIRST_MATCH is off)
 *   L2_Adaptor::evalMetric    ((0 + dx*dx) + dy*dy) + dz*dz for three dimensions
 * The pruning bound is a sum of rounded terms, so on last-ulp near-ties it could in principle drop a point an exact
 * search keeps; tests/test_oracle_golden.py compares the two on lattice clouds with massive ties, on the golden
 * clouds and on a 160 k-point workspace.  Parity with Open3D itself stays unpinned (it cannot be installed here). */
#define NF_LEAF 15
typedef struct { double lo, hi; } nf_interval;
typedef struct nf_node {
    int child1, child2;            /* -1: leaf */
    int64_t left, right;           /* leaf: range in vind */
    int divfeat; double divlow, divhigh;
} nf_node;
typedef struct {
    const double *P; int64_t n; int64_t *vind;
    nf_node *nodes; int n_nodes, cap_nodes;
    nf_interval root_bbox[3];
} nf_tree;

static int nf_new_node(nf_tree *t) {
    if (t->n_nodes == t->cap_nodes) { t->cap_nodes = t->cap_nodes ? 2 * t->cap_nodes : 1024; t->nodes = (nf_node *)realloc(t->nodes, sizeof(nf_node) * (size_t)t->cap_nodes); }
    return t->n_nodes++;
}
static void nf_minmax(const nf_tree *t, const int64_t *ind, int64_t count, int d, double *mn, double *mx) {
    *mn = *mx = t->P[3 * ind[0] + d];
    for (int64_t i = 1; i < count; ++i) { double v = t->P[3 * ind[i] + d]; if (v < *mn) *mn = v; if (v > *mx) *mx = v; }
}
static void nf_plane_split(const nf_tree *t, int64_t *ind, int64_t count, int cutfeat, double cutval, int64_t *lim1, int64_t *lim2) {
    int64_t left = 0, right = count - 1;
    for (;;) {
        while (left <= right && t->P[3 * ind[left] + cutfeat] < cutval) ++left;
        while (right && left <= right && t->P[3 * ind[right] + cutfeat] >= cutval) --right;
        if (left > right || !right) break;
        int64_t tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp;
        ++left; --right;
    }
    *lim1 = left;
    right = count - 1;
    for (;;) {
        while (left <= right && t->P[3 * ind[left] + cutfeat] <= cutval) ++left;
        while (right && left <= right && t->P[3 * ind[right] + cutfeat] > cutval) --right;
        if (left > right || !right) break;
        int64_t tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp;
        ++left; --right;
    }
    *lim2 = left;
}
static void nf_middle_split(const nf_tree *t, int64_t *ind, int64_t count, int64_t *index, int *cutfeat, double *cutval, const nf_interval *bbox) {
    const double EPS = 0.00001;
    double max_span = bbox[0].hi - bbox[0].lo;
    for (int i = 1; i < 3; ++i) { double span = bbox[i].hi - bbox[i].lo; if (span > max_span) max_span = span; }
    double max_spread = -1;
    *cutfeat = 0;
    for (int i = 0; i < 3; ++i) {
        double span = bbox[i].hi - bbox[i].lo;
        if (span > (1 - EPS) * max_span) {
            double mn, mx;
            nf_minmax(t, ind, count, i, &mn, &mx);
            double spread = mx - mn;
            if (spread > max_spread) { *cutfeat = i; max_spread = spread; }
        }
    }
    double split_val = (bbox[*cutfeat].lo + bbox[*cutfeat].hi) / 2;
    double mn, mx;
    nf_minmax(t, ind, count, *cutfeat, &mn, &mx);
    if (split_val < mn) *cutval = mn; else if (split_val > mx) *cutval = mx; else *cutval = split_val;
    int64_t lim1, lim2;
    nf_plane_split(t, ind, count, *cutfeat, *cutval, &lim1, &lim2);
    if (lim1 > count / 2) *index = lim1; else if (lim2 < count / 2) *index = lim2; else *index = count / 2;
}
static int nf_divide(nf_tree *t, int64_t left, int64_t right, nf_interval *bbox) {
    int id = nf_new_node(t);
    if (right - left <= NF_LEAF) {
        t->nodes[id].child1 = t->nodes[id].child2 = -1;
        t->nodes[id].left = left; t->nodes[id].right = right;
        for (int i = 0; i < 3; ++i) { bbox[i].lo = bbox[i].hi = t->P[3 * t->vind[left] + i]; }
        for (int64_t k = left + 1; k < right; ++k)
            for (int i = 0; i < 3; ++i) { double v = t->P[3 * t->vind[k] + i]; if (bbox[i].lo > v) bbox[i].lo = v; if (bbox[i].hi < v) bbox[i].hi = v; }
        return id;
    }
    int64_t idx; int cutfeat; double cutval;
    nf_middle_split(t, t->vind + left, right - left, &idx, &cutfeat, &cutval, bbox);
    nf_interval lb[3], rb[3];
    memcpy(lb, bbox, sizeof lb); memcpy(rb, bbox, sizeof rb);
    lb[cutfeat].hi = cutval;
    int c1 = nf_divide(t, left, left + idx, lb);
    rb[cutfeat].lo = cutval;
    int c2 = nf_divide(t, left + idx, right, rb);
    nf_node *nd = &t->nodes[id];     /* (re-read: the array may have moved) */
    nd->child1 = c1; nd->child2 = c2; nd->divfeat = cutfeat; nd->divlow = lb[cutfeat].hi; nd->divhigh = rb[cutfeat].lo;
    for (int i = 0; i < 3; ++i) { bbox[i].lo = lb[i].lo < rb[i].lo ? lb[i].lo : rb[i].lo; bbox[i].hi = lb[i].hi > rb[i].hi ? lb[i].hi : rb[i].hi; }
    return id;
}
typedef struct { double *dists; int capacity, count; } nf_result;
static inline double nf_worst(const nf_result *r) { return r->dists[r->capacity - 1]; }
static inline void nf_add(nf_result *r, double dist) {
    int i;
    for (i = r->count; i > 0; --i) {
        if (r->dists[i - 1] > dist) { if (i < r->capacity) r->dists[i] = r->dists[i - 1]; }
        else break;
    }
    if (i < r->capacity) r->dists[i] = dist;
    if (r->count < r->capacity) r->count++;
}
static void nf_search(const nf_tree *t, nf_result *res, const double *q, int node, double mindistsq, double *dists) {
    const nf_node *nd = &t->nodes[node];
    if (nd->child1 < 0) {
        double worst = nf_worst(res);
        for (int64_t i = nd->left; i < nd->right; ++i) {
            double d = dist2(q, t->P + 3 * t->vind[i]);
            if (d < worst) nf_add(res, d);
        }
        return;
    }
    int idx = nd->divfeat;
    double val = q[idx], diff1 = val - nd->divlow, diff2 = val - nd->divhigh, cut_dist;
    int best, other;
    if (diff1 + diff2 < 0) { best = nd->child1; other = nd->child2; cut_dist = (val - nd->divhigh) * (val - nd->divhigh); }
    else { best = nd->child2; other = nd->child1; cut_dist = (val - nd->divlow) * (val - nd->divlow); }
    nf_search(t, res, q, best, mindistsq, dists);
    double dst = dists[idx];
    mindistsq = mindistsq + cut_dist - dst;
    dists[idx] = cut_dist;
    if (mindistsq <= nf_worst(res)) nf_search(t, res, q, other, mindistsq, dists);
    dists[idx] = dst;
}
/* mean distance to the k nearest neighbours (self included) through the nanoflann-style tree */
void rt3d_oracle_knn_avg_kdtree(const double *P, int64_t n, int k, double *avg) {
    if (n == 0) return;
    nf_tree t;
    memset(&t, 0, sizeof t);
    t.P = P; t.n = n;
    t.vind = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) t.vind[i] = i;
    for (int i = 0; i < 3; ++i) t.root_bbox[i].lo = t.root_bbox[i].hi = P[i];
    for (int64_t j = 1; j < n; ++j)
        for (int i = 0; i < 3; ++i) { double v = P[3 * j + i]; if (v < t.root_bbox[i].lo) t.root_bbox[i].lo = v; if (v > t.root_bbox[i].hi) t.root_bbox[i].hi = v; }
    nf_interval bbox[3];
    memcpy(bbox, t.root_bbox, sizeof bbox);
    nf_divide(&t, 0, n, bbox);
    const int kk = (int64_t)k < n ? k : (int)n;
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)kk);
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            nf_result res = {buf, kk, 0};
            buf[kk - 1] = DBL_MAX;
            const double *q = P + 3 * i;
            double dists[3] = {0, 0, 0}, distsq = 0;
            for (int d = 0; d < 3; ++d) {       /* computeInitialDistances */
                if (q[d] < t.root_bbox[d].lo) { dists[d] = (q[d] - t.root_bbox[d].lo) * (q[d] - t.root_bbox[d].lo); distsq += dists[d]; }
                if (q[d] > t.root_bbox[d].hi) { dists[d] = (q[d] - t.root_bbox[d].hi) * (q[d] - t.root_bbox[d].hi); distsq += dists[d]; }
            }
            nf_search(&t, &res, q, 0, distsq, dists);
            double s = 0.0;                     /* Open3D: sqrt each, accumulate ascending from 0.0, / size */
            for (int j = 0; j < res.count; ++j) s += sqrt(buf[j]);
            avg[i] = res.count ? s / (double)res.count : -1.0;
        }
        free(buf);
    }
    free(t.vind); free(t.nodes);
}

int64_t rt3d_oracle_sor(const double *P, int64_t n, int k, double std_ratio, int brute, uint8_t *keep,
                        double *avg_out, double *stats) {
    if (n == 0) return 0;
    double *avg = avg_out ? avg_out : (double *)malloc(sizeof(double) * (size_t)n);
    rt3d_oracle_knn_avg(P, n, k, brute, avg);
    int64_t kept = rt3d_oracle_sor_threshold(avg, n, std_ratio, keep, stats);
    if (!avg_out) free(avg);
    return kept;
}

/* ------------------------------------------------------------------ S1 ----
 * subtract_point_clouds_gpu   2cam/vision_pipeline_utils.py:168-187
 * keep workspace row i  iff  min_j ||w_i - o_j|| > thr, thr as an fp32 scalar.
 * The reference evaluates this through torch.cdist's matrix-multiply form in
 * fp32 (abs error ~1.7e-5 [SURVEY]); the contract here is the exact predicate
 * on the actual fp32 coordinates, evaluated in fp64:
 *       keep  iff  min_j d2_ij > (double)thr32 * (double)thr32.
 * near[i] (optional) flags rows whose min distance is within `band` of thr --
 * the only rows on which the reference's rounding noise can disagree.        */
int64_t rt3d_oracle_subtract(const float *Wp, int64_t m, const float *Op, int64_t n, float thr,
                             uint8_t *keep, double band, uint8_t *near) {
    const double t2 = (double)thr * (double)thr;
    int64_t kept = 0;
#pragma omp parallel for schedule(static) reduction(+ : kept)
    for (int64_t i = 0; i < m; ++i) {
        double w[3] = {Wp[3 * i], Wp[3 * i + 1], Wp[3 * i + 2]};
        double best = INFINITY;
        for (int64_t j = 0; j < n; ++j) {
            double o[3] = {Op[3 * j], Op[3 * j + 1], Op[3 * j + 2]};
            double d2 = dist2(w, o);
            if (d2 < best) {
                best = d2;
                if (!near && best <= t2) break; /* already removed */
            }
        }
        keep[i] = best > t2;
        kept += keep[i];
        if (near) near[i] = fabs(sqrt(best) - (double)thr) <= band;
    }
    return kept;
}

/* S1 again, for clouds too large for the O(M N) loop above (HD1080 at 2 mm: 1e6 x 5e5): the same exact
 * predicate  keep <=> min_j |w - o_j| > thr  (fp64 on the fp32 coordinates, 2cam/vision_pipeline_utils.py:175-182),
 * with the object points binned into a dense grid of edge thr/6 over their bounding box.  A cell is skipped only
 * when its box is farther than thr by a safe margin, so every point that could be within thr is still compared
 * exactly; tests/test_oracle_golden.py checks it against rt3d_oracle_subtract.                                  */
typedef struct { int64_t cell; int64_t idx; } cell_item;
static int cell_item_cmp(const void *a, const void *b) {
    const cell_item *x = (const cell_item *)a, *y = (const cell_item *)b;
    if (x->cell != y->cell) return x->cell < y->cell ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
int64_t rt3d_oracle_subtract_grid(const float *Wp, int64_t m, const float *Op, int64_t n, float thr, uint8_t *keep) {
    const double t = (double)thr, t2 = t * t;
    if (n == 0) { for (int64_t i = 0; i < m; ++i) keep[i] = 1; return m; }
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t j = 0; j < n; ++j)
        for (int a = 0; a < 3; ++a) { double v = Op[3 * j + a]; if (v < lo[a]) lo[a] = v; if (v > hi[a]) hi[a] = v; }
    double h = t / 6.0;
    int dim[3];
    for (;;) {   /* at most 2^21 cells */
        double cells = 1;
        for (int a = 0; a < 3; ++a) { dim[a] = h > 0 ? (int)floor((hi[a] - lo[a]) / h) + 1 : 1; cells *= dim[a]; }
        if (h > 0 && cells <= 2097152.0) break;
        h = h > 0 ? h * 2 : 1e-3;
    }
    const int64_t ncell = (int64_t)dim[0] * dim[1] * dim[2];
    cell_item *items = (cell_item *)malloc((size_t)n * sizeof(cell_item));
    int64_t *start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
    float *sorted = (float *)malloc((size_t)n * 3 * sizeof(float));
    if (!items || !start || !sorted) { free(items); free(start); free(sorted); return -1; }
    for (int64_t j = 0; j < n; ++j) {
        int c[3];
        for (int a = 0; a < 3; ++a) {
            c[a] = (int)floor(((double)Op[3 * j + a] - lo[a]) / h);
            if (c[a] < 0) c[a] = 0;
            if (c[a] >= dim[a]) c[a] = dim[a] - 1;
        }
        items[j].cell = ((int64_t)c[0] * dim[1] + c[1]) * dim[2] + c[2];
        items[j].idx = j;
    }
    qsort(items, (size_t)n, sizeof(cell_item), cell_item_cmp);
    for (int64_t j = 0; j < n; ++j) {
        start[items[j].cell + 1]++;
        memcpy(sorted + 3 * j, Op + 3 * items[j].idx, 3 * sizeof(float));
    }
    for (int64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    const int reach = (int)ceil(t / h) + 1;
    const double skip2 = t2 * (1.0 + 1e-9) + 1e-18;
    int64_t kept = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : kept)
    for (int64_t i = 0; i < m; ++i) {
        const double w[3] = {Wp[3 * i], Wp[3 * i + 1], Wp[3 * i + 2]};
        int c0[3], c1[3], outside = 0;
        for (int a = 0; a < 3; ++a) {
            if (w[a] < lo[a] - t * (1.0 + 1e-9) - 1e-12 || w[a] > hi[a] + t * (1.0 + 1e-9) + 1e-12) outside = 1;
            const int c = (int)floor((w[a] - lo[a]) / h);
            c0[a] = c - reach < 0 ? 0 : c - reach;
            c1[a] = c + reach >= dim[a] ? dim[a] - 1 : c + reach;
        }
        int hit = 0;
        if (!outside) {
            for (int x = c0[0]; x <= c1[0] && !hit; ++x)
                for (int y = c0[1]; y <= c1[1] && !hit; ++y)
                    for (int z = c0[2]; z <= c1[2] && !hit; ++z) {
                        const int64_t cell = ((int64_t)x * dim[1] + y) * dim[2] + z;
                        if (start[cell] == start[cell + 1]) continue;
                        const int cc[3] = {x, y, z};
                        double dmin2 = 0;
                        for (int a = 0; a < 3; ++a) {   /* the cell's box, one ulp-proof margin wider than its points */
                            const double bl = lo[a] + cc[a] * h - 1e-9, bh = lo[a] + (cc[a] + 1) * h + 1e-9;
                            const double d = w[a] < bl ? bl - w[a] : (w[a] > bh ? w[a] - bh : 0.0);
                            dmin2 += d * d;
                        }
                        if (dmin2 > skip2) continue;
                        for (int64_t j = start[cell]; j < start[cell + 1] && !hit; ++j) {
                            const double o[3] = {sorted[3 * j], sorted[3 * j + 1], sorted[3 * j + 2]};
                            hit = dist2(w, o) <= t2;
                        }
                    }
        }
        keep[i] = !hit;
        kept += keep[i];
    }
    free(items); free(start); free(sorted);
    return kept;
}

/* S2: KD-tree radius variant  2cam/2cams_mask_cpu.py:246-270 -- a workspace point
 * is removed iff some object point has d2 < r*r in fp64 (strict; ties kept).  */
int64_t rt3d_oracle_subtract_radius(const double *Wp, int64_t m, const double *Op, int64_t n, double r, uint8_t *keep) {
    const double r2 = r * r;
    int64_t kept = 0;
#pragma omp parallel for schedule(static) reduction(+ : kept)
    for (int64_t i = 0; i < m; ++i) {
        int hit = 0;
        for (int64_t j = 0; j < n && !hit; ++j) hit = dist2(Wp + 3 * i, Op + 3 * j) < r2;
        keep[i] = !hit;
        kept += keep[i];
    }
    return kept;
}
