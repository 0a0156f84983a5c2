/*
 * rt3d.h -- C ABI of librt3d.so: the B200-native (sm_100a) geometric post-processing path of
 * janfrischi/RT-3D-Reconstruction, `2cam/vision_pipeline_utils.py`.
 *
 * The reference has no FFI of its own -- its boundary is the set of module-level Python
 * functions in 2cam/vision_pipeline_utils.py.  Each entry point below names the reference
 * function (file:line, relative to the reference root) whose arithmetic it replaces; the
 * Python drop-in (`rt-3d-reconstruction_b200/vision_pipeline_utils.py`) binds them with ctypes
 * and INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - Plain C: pointers + sizes, no torch / C++ types.  Unless marked HOST, every pointer is a
 *    device pointer valid on the current CUDA device.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises
 *    (rt3d_frame_batch_init excepted), reads no environment variable, and is re-entrant per stream; the
 *    only process-wide state is a mutex-guarded cache of "fast division verified for divisor b" results,
 *    consulted by rt3d_frame_batch_init alone.  Callers own all memory,
 *    including the scratch arena (size it with rt3d_scratch_bytes; contents need no initialisation).
 *  - Result counts are written to device memory (`out_count`, int32) so that a frame can be
 *    chained or graph-captured without a host round trip.  Output buffers must hold the worst
 *    case (the input item count).
 *  - Return value: RT3D_OK or a negative rt3d_status; rt3d_last_error() gives the text (thread
 *    local).  Device-side faults (key range overflow, scratch look-back timeout) are reported
 *    in the int32 status word at the very start of the scratch arena (RT3D_DEV_* bits).
 *  - Point clouds are row-major (N,3).  fp32 unless a dtype argument says otherwise.
 *  - `flags` selects the arithmetic flavour of `tensor / python_float` in the reference:
 *      RT3D_DIV_IEEE   true IEEE division      (torch-CPU / NumPy; what the golden vectors hold)
 *      RT3D_DIV_RECIP  multiply by float(1.0/double(b))  (torch-CUDA's div-by-scalar kernel, i.e. the
 *                                               reference as deployed on a GPU; measured on torch 2.11)
 */
#ifndef RT3D_H_
#define RT3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT3D_ABI_VERSION 2

#if defined(__GNUC__)
#define RT3D_API __attribute__((visibility("default")))
#else
#define RT3D_API
#endif

typedef void *rt3d_stream_t; /* cudaStream_t */

typedef enum {
    RT3D_OK = 0,
    RT3D_EINVAL = -1,   /* bad argument */
    RT3D_ECUDA = -2,    /* CUDA runtime error at launch */
    RT3D_ESCRATCH = -3, /* scratch arena too small */
    RT3D_ERANGE = -4    /* size beyond what the kernels index (n >= 2^30) */
} rt3d_status;

enum { RT3D_U8 = 0, RT3D_F32 = 1, RT3D_F64 = 2 };     /* element types */
enum { RT3D_DIV_IEEE = 0, RT3D_DIV_RECIP = 1 };        /* arithmetic flavour (flags bit 0) */
/* rt3d_frame_config.flags bits 2-3: which project kernel streams the depth maps.  AUTO picks the tensor-map (TMA)
 * producer/consumer pipeline for long batches of images whose width is a multiple of 16 and at most 2 instance masks,
 * the per-lane cp.async pipeline otherwise; both run the same per-tile arithmetic, so results never depend on it. */
enum { RT3D_PROJECT_AUTO = 0, RT3D_PROJECT_TMA = 4, RT3D_PROJECT_CP_ASYNC = 8, RT3D_PROJECT_MASK = 12 };
/* rt3d_frame_config.flags bits 1 and 4: the outlier filter's occupancy bit-plane neighbour search in front of the grid
 * search (same results either way).  Neither bit: automatic (on where its four extra launches pay: obj_capacity >=
 * 1 M points and voxel >= 3 mm). */
enum { RT3D_SOR_GRID_ONLY = 2, RT3D_SOR_BITPLANES = 16 };

/* device-side status bits OR-ed into the int32 at the start of a scratch arena */
enum {
    RT3D_DEV_OK = 0,
    RT3D_DEV_KEY_RANGE = 1, /* a voxel index fell outside the packable range (non-finite input?) */
    RT3D_DEV_LOOKBACK = 2,  /* decoupled look-back gave up waiting (should never happen) */
    RT3D_DEV_CAPACITY = 4   /* an output or intermediate buffer was too small */
};

/* operations for rt3d_scratch_bytes */
enum {
    RT3D_OP_SELECT = 0,        /* mask_nonzero, backproject_*, xyz_ingest, crop: n = items            */
    RT3D_OP_VOXEL_UNIQUE = 1,  /* voxel_snap_unique: n = points                                        */
    RT3D_OP_VOXEL_CENTROID = 2,/* voxel_centroid: n = points                                           */
    RT3D_OP_SOR = 3,           /* sor: n = points                                                      */
    RT3D_OP_SUBTRACT = 4,      /* subtract: n = workspace points, m = object points                    */
    RT3D_OP_KEYS_UNIQUE = 5    /* keys_sort_unique: n = keys                                           */
};

RT3D_API int rt3d_abi_version(void);
RT3D_API const char *rt3d_last_error(void);
/* HOST query: bytes of scratch an operation needs for n (and m) items. */
RT3D_API size_t rt3d_scratch_bytes(int op, int64_t n, int64_t m);

/* ---- mask pre-processing in front of the back-projection
 * erode_mask_gpu(mask, kernel_size=3)                      2cam/vision_pipeline_utils.py:12-15
 *   conv2d(mask, ones(k,k), padding=k//2) > 0 as 0/1 fp32: a box *dilation* (sic).  out is
 *   (H + 1 - k % 2... i.e. H + 2*(k/2) - k + 1) x (W + 2*(k/2) - k + 1), the conv2d output shape.   */
RT3D_API int rt3d_box_dilate(const void *mask, int mask_dtype, int H, int W, int k, float *out, rt3d_stream_t stream);
/* erode_mask(mask, iterations=1)                           2cam/2cams_mask_cpu.py:54-57
 *   cv2.erode with a k x k ones kernel (k = 12 there), anchor (k/2, k/2), pixels beyond the border
 *   ignored.  uint8 in/out; tmp (H*W bytes) is needed when iterations > 1.                          */
RT3D_API int rt3d_box_erode(const uint8_t *mask, int H, int W, int k, int iterations, uint8_t *out, uint8_t *tmp,
                            rt3d_stream_t stream);

/* ---- B0  torch.nonzero(mask, as_tuple=False)            2cam/vision_pipeline_utils.py:296
 * mask (H,W) of mask_dtype (RT3D_U8 or RT3D_F32; non-zero incl. NaN counts).  out_idx (K,2)
 * int64 [row, col] in row-major order.                                                         */
RT3D_API int rt3d_mask_nonzero(const void *mask, int mask_dtype, int H, int W, int64_t *out_idx, int32_t *out_count,
                      void *scratch, size_t scratch_bytes, rt3d_stream_t stream);

/* ---- B1  convert_mask_to_3d_points(mask_indices, depth_map, cx, cy, fx, fy)     :19-36
 * idx (K,2) int64 [row, col]; intr = HOST doubles {cx, cy, fx, fy} (the Python floats, unrounded:
 * the torch-CUDA flavour needs them in double).  Keeps d > 0 and finite, output in input order.
 *                                                                                */
RT3D_API int rt3d_backproject_indices(const int64_t *idx, int64_t K, const float *depth, int H, int W, const double *intr,
                             int flags, float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes,
                             rt3d_stream_t stream);

/* ---- B0+B1+T1+C1 fused: one pass over a depth image                   :296, :19-36, :301, :78-84
 * mask may be NULL (every pixel; the synthetic stand-in for the ZED XYZ workspace cloud).
 * Rt = HOST {R row-major (9), t (3)} or NULL (camera frame).  box = HOST {xlo,xhi,ylo,yhi,zlo,zhi}
 * closed intervals or NULL.  Output order = row-major pixel order.                             */
RT3D_API int rt3d_backproject_mask(const float *depth, const void *mask, int mask_dtype, int H, int W, const double *intr,
                          const float *Rt, const float *box, int flags, float *out_xyz, int32_t *out_count,
                          void *scratch, size_t scratch_bytes, rt3d_stream_t stream);

/* ---- W0+T1+C1  workspace ingest of process_point_clouds                          :233-245
 * xyz: n rows of `stride` floats (4 = ZED MEASURE.XYZ F32_C4, 3 = packed).  finite_filter != 0
 * drops rows with a non-finite x/y/z (:235-238); Rt / box as above (NULL to skip).  With
 * finite_filter = 0, Rt = NULL this is crop_point_cloud_gpu (:78-84).                          */
RT3D_API int rt3d_xyz_select(const float *xyz, int64_t n, int stride, int finite_filter, const float *Rt, const float *box,
                    float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes, rt3d_stream_t stream);

/* ---- T1  torch.mm(P, R.T) + t   :301 / :239-240 -- fma chain over k then a rounded +t.        */
RT3D_API int rt3d_transform(const float *xyz, int64_t n, const float *Rt, float *out_xyz, rt3d_stream_t stream);
/* ---- T1d np.dot(R64, P.T).T + t64   2cam/2cams_mask_cpu.py:595,618 (fp32 in, fp64 out)        */
RT3D_API int rt3d_transform_f64(const float *xyz, int64_t n, const double *Rt, double *out_xyz, rt3d_stream_t stream);

/* ---- V1  downsample_point_cloud_gpu(point_cloud, voxel_size)                        :40-44
 * q = rint(p / v) * v, lexicographically sorted distinct rows.  n_dev (optional) overrides n
 * with a device-resident count (n is then the capacity).  out_keys (optional) receives the
 * packed voxel keys of the rows, out_layout (optional, 8 x int32 device) how they were packed:
 * {min_x,min_y,min_z,bits_x,bits_y,bits_z,0,0}.                                                */
RT3D_API int rt3d_voxel_snap_unique(const float *xyz, int64_t n, const int32_t *n_dev, double voxel, int flags, float *out_xyz,
                           uint64_t *out_keys, int32_t *out_layout, int32_t *out_count, void *scratch,
                           size_t scratch_bytes, rt3d_stream_t stream);

/* ---- V2  downsample_point_cloud (Open3D tensor voxel_down_sample, "mean")  2cam/2cams_mask_cpu.py:60-65
 * key = floor(p / v) in the input dtype (RT3D_F32 / RT3D_F64); per-voxel mean of the positions
 * cast to fp32, accumulated in input order; output in the input dtype, voxels sorted by key
 * (Open3D's own order is its hash-buffer order).  out_cnt (optional) = points per voxel.       */
RT3D_API int rt3d_voxel_centroid(const void *xyz, int dtype, int64_t n, double voxel, void *out_xyz, int32_t *out_cnt,
                        int32_t *out_count, void *scratch, size_t scratch_bytes, rt3d_stream_t stream);

/* ---- O1  filter_outliers_sor(point_cloud, nb_neighbors=20, std_ratio=1.5)            :48-58
 * Open3D remove_statistical_outlier in fp64 (see oracle/rt3d_oracle.c for the restated algorithm).
 * xyz of dtype RT3D_F32 / RT3D_F64.  out_keep (n) uint8; out_xyz (n,3) fp64 compacted survivors in
 * input order; out_stats (optional, 4 doubles) = {mu, sigma, threshold, valid}; out_avg (optional,
 * n doubles) = mean k-NN distance per point.  1 <= k <= 64.                                    */
RT3D_API int rt3d_sor(const void *xyz, int dtype, int64_t n, int k, double std_ratio, uint8_t *out_keep, double *out_xyz,
             int32_t *out_count, double *out_stats, double *out_avg, void *scratch, size_t scratch_bytes,
             rt3d_stream_t stream);

/* ---- calculate_centroid = np.mean(pc, axis=0)                                         :74-75
 * sequential sum in the array's dtype then / n (NumPy's order for a C-contiguous (N,3) array).  */
RT3D_API int rt3d_centroid(const void *xyz, int dtype, int64_t n, void *out3, rt3d_stream_t stream);

/* ---- S1  subtract_point_clouds_gpu(workspace_pc, objects_pc, distance_threshold)   :168-187
 * keeps workspace row i iff min_j |w_i - o_j| > thr, evaluated exactly (fp64) on the fp32
 * coordinates; order preserved.  out_keep optional.  n == 0 is the caller's IndexError.        */
RT3D_API int rt3d_subtract(const float *ws, int64_t m, const float *obj, int64_t n, float thr, uint8_t *out_keep,
                  float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes, rt3d_stream_t stream);

/* ---- S2  subtract_point_clouds(workspace_pc, objects_pc, distance_threshold=0.01)   2cam/2cams_mask_cpu.py:246-270
 * the KD-tree variant: a workspace row is removed iff search_radius_vector_3d(row, radius) finds an object
 * point, i.e. d^2 < radius^2 STRICTLY (nanoflann's radius search), so ties are kept -- the opposite of S1.
 * Same exact fp64 evaluation on the fp32 coordinates, same buffers as rt3d_subtract.                  */
RT3D_API int rt3d_subtract_radius(const float *ws, int64_t m, const float *obj, int64_t n, double radius, uint8_t *out_keep,
                                  float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes,
                                  rt3d_stream_t stream);

/* ---- voxel-key set operations (optional box-wide fusion over all-gathered keys)
 * sort + unique of n packed keys (n_dev optional device count).                                */
RT3D_API int rt3d_keys_sort_unique(const uint64_t *keys, int64_t n, const int32_t *n_dev, int key_bits, uint64_t *out_keys,
                          int32_t *out_count, void *scratch, size_t scratch_bytes, rt3d_stream_t stream);
/* per-call keys (+ their layout) -> canonical keys: 21 bits per axis, biased by 2^20, x major -- comparable
 * across calls and GPUs (what gets all-gathered).  out_layout (optional, 8 x int32 device) receives the
 * canonical layout for rt3d_keys_decode; status (device int32) collects RT3D_DEV_KEY_RANGE.               */
RT3D_API int rt3d_keys_rebase(const uint64_t *keys, int64_t n, const int32_t *n_dev, const int32_t *layout,
                              uint64_t *out_keys, int32_t *out_layout, int32_t *status, rt3d_stream_t stream);
/* keys -> lattice points k * v using a layout written by rt3d_voxel_snap_unique.               */
RT3D_API int rt3d_keys_decode(const uint64_t *keys, int64_t n, const int32_t *n_dev, const int32_t *layout, float voxel,
                     float *out_xyz, rt3d_stream_t stream);

/* ---- batched frame front-end: the whole geometric path of one 2cams.py loop
 * iteration (2cam/2cams.py:165-208: process_point_clouds, process_masks, fuse_point_clouds,
 * subtract_point_clouds) for F frames x C cameras in one call, device resident, no host round trip.
 *
 * Workspace cloud of a camera = every valid depth pixel, back-projected (:19-36), transformed (:239-240),
 * cropped to `box` (:78-84, :241-243) and voxelised (:40-44) -- the synthetic stand-in for the ZED XYZ
 * measure; a frame's workspace is the concatenation of its cameras' voxel sets (:248).  Object o of
 * camera c = process_masks (:292-304) of instance mask o (an object without a valid pixel yields no
 * cloud, :300).  Subtraction (:168-187) runs per frame against the frame's fused objects (all fused clouds
 * stacked, as fuse_point_clouds :313-317 does); a frame without any object point keeps its whole workspace.
 *
 * Fusion, two modes:
 *  RT3D_FUSE_BY_INDEX   instance mask o of every camera is the same physical object (a tracker has paired them):
 *      the 1-vs-1 branch of fuse_point_clouds_centroid (:123-126) -- the cameras' clouds of object o are
 *      stacked in camera order and filtered with remove_statistical_outlier; a pair whose camera-0 cloud is
 *      empty is passed through unfiltered (:162-163).  Any C <= 8.  One output slot per (frame, object).
 *  RT3D_FUSE_CENTROID   the whole of fuse_point_clouds_centroid (:88-165) for C == 2, per frame: clouds are
 *      grouped by class (`class_rank`); a class seen once by each camera is stacked and filtered (:123-126);
 *      otherwise every camera-1 cloud is filtered (:131), matched greedily to the nearest camera-2 cloud
 *      by centroid distance (centroid of the FILTERED camera-1 cloud in fp64, of the UNFILTERED camera-2
 *      cloud in fp32, strict < best and < fuse_thr, :136-147), stacked with the filtered match without a
 *      final filter (:152); the match leaves the candidate list only if it equals its filtered copy (:155),
 *      i.e. when the filter dropped nothing; what remains of camera 2 is appended unfiltered (:162-163).
 *      2 * n_obj output slots per frame, filled in the reference's list order (classes by ascending rank),
 *      unused slots empty.                                                                             */
enum { RT3D_FUSE_BY_INDEX = 0, RT3D_FUSE_CENTROID = 1 };
/* kind of an output cloud (obj_meta[.][0]) */
enum {
    RT3D_CLOUD_EMPTY = 0,     /* unused slot */
    RT3D_CLOUD_PAIR = 1,      /* filter(stack(camera clouds))                       (:123-126; float64 in the reference) */
    RT3D_CLOUD_MATCHED = 2,   /* stack(filter(camera-1 cloud), filter(match))       (:152; float64)                      */
    RT3D_CLOUD_UNMATCHED = 3, /* filter(camera-1 cloud)                             (:160; float64)                      */
    RT3D_CLOUD_LEFTOVER = 4   /* camera-2 cloud as is                               (:162-163; float32 in the reference) */
};

typedef struct {
    int32_t F, C, H, W, n_obj;   /* frames, cameras per frame (<= 8), image size (W % 4 == 0), masks per camera */
    double intr[8][4];           /* per camera {cx, cy, fx, fy} */
    float Rt[8][12];             /* per camera {R row-major, t} */
    float box[6];                /* workspace crop {xlo,xhi,ylo,yhi,zlo,zhi} */
    double voxel;
    int32_t flags;               /* RT3D_DIV_* | RT3D_SOR_GRID_ONLY / RT3D_SOR_BITPLANES | RT3D_PROJECT_* */
    double max_range;            /* |coordinate| bound of object points in metres (key width); beyond -> RT3D_DEV_KEY_RANGE */
    int32_t sor_k;               /* nb_neighbors (<= 64) */
    double sor_std_ratio;
    float sub_thr;               /* subtraction distance threshold */
    int64_t ws_capacity;         /* rows of ws_xyz / kept_xyz */
    int64_t obj_capacity;        /* object pixels per batch the intermediate buffers can hold */
    int32_t fuse_mode;           /* RT3D_FUSE_* */
    double fuse_thr;             /* distance_threshold of fuse_point_clouds_centroid (RT3D_FUSE_CENTROID) */
    int32_t mask_dtype;          /* RT3D_U8 (0/1 bytes) or RT3D_F32 (the ultralytics masks.data planes, :295; non-zero = set) */
    int64_t fused_capacity;      /* rows of obj_xyz; 0 = obj_capacity.  RT3D_FUSE_CENTROID can emit a camera-2 cloud more than once */
    int32_t ws_sor;              /* != 0: every frame's fused workspace (its cameras' voxel sets concatenated) goes through
                                    remove_statistical_outlier(sor_k, sor_std_ratio) before the subtraction, as
                                    2cam/2cams_mask_cpu.py:530 does; ws_xyz / ws_off still report the unfiltered sets */
    int32_t plan;                /* written by rt3d_frame_batch_init (decisions taken once per configuration); 0 before */
} rt3d_frame_config;

/* output slots per frame: n_obj (RT3D_FUSE_BY_INDEX) or 2 * n_obj (RT3D_FUSE_CENTROID) */
RT3D_API int rt3d_frame_batch_slots(const rt3d_frame_config *cfg);
RT3D_API size_t rt3d_frame_batch_scratch_bytes(const rt3d_frame_config *cfg);
/* once per scratch arena and configuration: zeroes the occupancy bitmaps (every run leaves them empty again), builds
 * the per-camera depth-interval cull table, qualifies the fast IEEE division for the configuration's divisors, reads
 * back whether the cull table reaches beyond that division's verified depth range, and records those decisions in
 * cfg->plan (run() reads them from there: no process-global state, no environment).  Synchronises `stream` once. */
RT3D_API int rt3d_frame_batch_init(rt3d_frame_config *cfg, void *scratch, size_t scratch_bytes, rt3d_stream_t stream);
/* depth [F*C][H*W] fp32; masks [F*C][n_obj][H*W] of cfg->mask_dtype (NULL when n_obj == 0);
 * class_rank [F*C*n_obj] int32 (RT3D_FUSE_CENTROID; NULL otherwise): equal rank = same class, classes are visited
 *   in ascending rank, 0 <= rank < 2 * n_obj (the Python front-end ranks the class ids in the reference's set order).
 * ws_xyz (ws_capacity,3) + ws_off [F*C+1]: per-image sorted voxel sets, image-major (frame f = images f*C..f*C+C-1).
 * obj_xyz (fused_capacity,3) fp64 + obj_off [F*slots+1]: fused object clouds, (frame, slot)-major.
 * obj_meta (optional) [F*slots][4] int32: {RT3D_CLOUD_* kind, class rank (object index for RT3D_FUSE_BY_INDEX),
 *   rows of the first part, rows of the second part (RT3D_CLOUD_MATCHED: the filtered match)}.
 * kept_xyz (ws_capacity,3) + kept_off [F+1]: subtracted workspace per frame.
 * counts [8] = {device status bits, object pixels, object voxels, fused rows, workspace voxels, kept points,
 *   outlier-filter queries answered on the occupancy bit planes (diagnostic), workspace rows surviving ws_sor}. */
RT3D_API int rt3d_frame_batch_run(const rt3d_frame_config *cfg, const float *depth, const void *masks, const int32_t *class_rank,
                                  float *ws_xyz, int32_t *ws_off, double *obj_xyz, int32_t *obj_off, int32_t *obj_meta,
                                  float *kept_xyz, int32_t *kept_off, int32_t *counts, void *scratch, size_t scratch_bytes,
                                  rt3d_stream_t stream);

/* Per-stage device timing of a batch (CUDA events recorded on the launching stream between the stages
 * A project, B extract, C object sort/unique, D fuse (SOR), E subtract).  read() waits for the last event
 * and returns the five stage durations in milliseconds. */
typedef struct rt3d_stage_timer rt3d_stage_timer;
RT3D_API rt3d_stage_timer *rt3d_stage_timer_create(void);
RT3D_API void rt3d_stage_timer_destroy(rt3d_stage_timer *timer);
RT3D_API int rt3d_stage_timer_read(rt3d_stage_timer *timer, float *ms5);
RT3D_API int rt3d_frame_batch_run_timed(const rt3d_frame_config *cfg, const float *depth, const void *masks,
                                        const int32_t *class_rank, float *ws_xyz, int32_t *ws_off, double *obj_xyz,
                                        int32_t *obj_off, int32_t *obj_meta, float *kept_xyz, int32_t *kept_off,
                                        int32_t *counts, void *scratch, size_t scratch_bytes, rt3d_stream_t stream,
                                        rt3d_stage_timer *timer);

#ifdef __cplusplus
}
#endif
#endif /* RT3D_H_ */
