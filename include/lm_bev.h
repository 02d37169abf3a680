/* lm_bev.h -- C-ABI of the B200-native BEV rasteriser (liblm_bev.so).
 *
 * What this boundary replaces.  The reference has NO FFI and no in-tree rasteriser: its
 * BEV images come from an external tool (reference README.md:171-172) and enter the repo
 * as files read by PIL (reference baseline/datasets/laserlane_proposals.py:85-98) and by
 * the inverse map (reference baseline/utils/coor_img2pc.py:185-193).  The entry points
 * below are what a ctypes/cffi stub inside the reference would bind to produce those
 * files -- or the in-memory sample['proj'] tensor -- on a B200.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - every *_dev pointer is DEVICE memory owned by the caller (PyTorch allocates and
 *     frees it).  The library never allocates device memory and never synchronises the
 *     device: it only enqueues kernels / memsets on the stream passed in
 *     (a cudaStream_t carried as void*; NULL = legacy default stream).
 *   - return value: 0 = OK; <0 = argument error detected before any launch (LM_ERR_*);
 *     >0 = the cudaError_t of a failed launch.  lm_bev_last_error() gives the message
 *     for the calling thread.  Re-entrant: no global mutable state besides that
 *     thread-local message; no environment variables are read (tuning knobs are fields of
 *     lm_bev_tuning, passed through a plan).
 *   - device-side conditions that cannot be known before launch (workspace chunk pool
 *     exhausted, per-cell count above the u32-sum limit) are reported in lm_bev_stats.error,
 *     which lives at the start of the workspace.
 */
#ifndef LM_BEV_H
#define LM_BEV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_BEV_ABI_VERSION 3

/* u8 image channels.  Index 1 of a 3-channel cropped_tiff must be an elevation channel:
 * reference baseline/utils/coor_img2pc.py:150 reads img[row, col, 1] as height. */
enum {
    LM_CH_MAX_I   = 0,  /* max quantised intensity                                     */
    LM_CH_MEAN_I  = 1,  /* (sum_i + count/2) / count                                   */
    LM_CH_MIN_Z   = 2,  /* min quantised height                                        */
    LM_CH_MAX_Z   = 3,  /* max quantised height                                        */
    LM_CH_MEAN_Z  = 4,  /* (sum_z + count/2) / count                                   */
    LM_CH_DENSITY = 5,  /* min(count, 255): >= 1 in every occupied cell, so an occupied */
                        /* pixel is never all-zero (coor_img2pc.py:78,106 empty rule)  */
    LM_CH__COUNT  = 6
};

/* raw u32 accumulator planes, [LM_ACC_PLANES][height][width]; used to merge strip halos */
enum {
    LM_ACC_COUNT = 0, LM_ACC_SUM_I = 1, LM_ACC_SUM_Z = 2,
    LM_ACC_MAX_I = 3, LM_ACC_MIN_Z = 4 /* 0xFFFFFFFF where count==0 */, LM_ACC_MAX_Z = 5,
    LM_ACC_PLANES = 6
};

enum {
    LM_ALGO_BINNED = 0, /* product path: bin -> index -> per-tile shared-memory reduce  */
    LM_ALGO_DIRECT = 1, /* global-atomic accumulate + finalize; cross-check path        */
    LM_ALGO_SWEEP  = 2  /* EXPERIMENTAL single-pass sweep: records are routed through L2-resident        */
                        /* mailboxes to column-owner CTAs (no record pool in HBM: DRAM traffic = 1.0 x   */
                        /* algorithmic bytes), with the BINNED kernels queued behind it as the exact     */
                        /* fall-back (they return at once unless the sweep gave up).  Same results on    */
                        /* every input; slower than BINNED today (instruction-bound).  Needs  */
                        /* lm_bev_workspace_init once per workspace.                                     */
};

enum {
    LM_OK = 0,
    LM_ERR_INVALID = -1,     /* NULL / out-of-range argument                            */
    LM_ERR_WORKSPACE = -2,   /* workspace too small or misaligned                       */
    LM_ERR_UNSUPPORTED = -3  /* raster too large for one call (shard it by row window)  */
};

/* lm_bev_stats.error bits */
enum {
    LM_DEV_ERR_POOL = 1,      /* chunk pool exhausted (cannot happen with the size      */
                              /* lm_bev_workspace_bytes returns)                        */
    LM_DEV_ERR_CELL_OVERFLOW = 2 /* a cell received >= 2^24 points: u32 sums may wrap   */
};

/* Geometry and channel list.  Field names follow the keys of the sidecar
 * cropped_tiff_param/<stem>.txt (reference baseline/utils/io_utils.py:125-150):
 *   row = floor((x - bev_img_offset[0]) / img_reso[0])   (inverse of coor_img2pc.py:136-139)
 *   col = floor((y - bev_img_offset[1]) / img_reso[1])
 *   zq  = clamp(rint((z - local_min_ele) / ele_reso), 0, 255)   (inverse of :150)
 *   iq  = (clamp(I, inten_min, inten_max) - inten_min) * 255 / (inten_max - inten_min)
 *         (clip range: reference baseline/datasets/laserlane_proposals.py:626-628)
 * all float steps are single IEEE binary32 operations; a point is kept iff
 * row0 <= row < row0+height and col0 <= col < col0+width (never clamped), and is stored
 * at (row-row0, col-col0).  |row0|+height and |col0|+width must stay below 2^24.       */
typedef struct lm_bev_params {
    int32_t height, width;        /* window size in cells                               */
    int32_t row0, col0;           /* window origin in the global grid                   */
    float   bev_img_offset[2];
    float   img_reso[2];
    float   local_min_ele;
    float   ele_reso;
    int32_t inten_min, inten_max; /* 0 <= inten_min < inten_max <= 65535                */
    int32_t n_channels;           /* 1..4 u8 channels                                   */
    int32_t channels[4];          /* LM_CH_*                                            */
} lm_bev_params;

/* Output buffers (device).  Any pointer may be NULL = not wanted, but at least one must
 * be set.  Every cell of every requested buffer is written exactly once per call.      */
typedef struct lm_bev_outputs {
    uint8_t  *image_dev;    /* [height][width][n_channels] u8 (the PNG pixel layout)    */
    uint16_t *count16_dev;  /* [height][width] u16 = min(count, 65535)                  */
    float    *proj_dev;     /* [n_channels][height][width] f32 = u8/255: the loader's   */
                            /* to_tensor(...).float() (laserlane_proposals.py:88-89)    */
    uint32_t *acc_dev;      /* [LM_ACC_PLANES][height][width] raw accumulators          */
    int32_t   acc_band;     /* <=0: all six planes, all rows.  >0 (halo mode, with another */
                            /* output requested): only tiles touching rows [0,band) or    */
                            /* [height-band,height) are written, and only the planes the  */
                            /* requested channels need are accumulated (others read empty) */
    int32_t   reserved;
} lm_bev_outputs;

/* First bytes of the workspace after a call (read it back with a D2H copy if wanted). */
typedef struct lm_bev_stats {
    uint32_t error;         /* LM_DEV_ERR_* bits                                        */
    uint32_t n_chunks;      /* record chunks used by the binned path                    */
    uint64_t n_valid;       /* points that fell inside the window                       */
    uint32_t n_tiles;       /* shared-memory tiles of the binned path                   */
    uint32_t ct_overflow;   /* 1: the compact-table pass met more tiles per CTA than its  */
                            /* table holds and the direct-indexed kernels redid the call  */
                            /* (informational: the result is exact either way)            */
    uint32_t reserved[2];
} lm_bev_stats;

int         lm_bev_abi_version(void);
const char *lm_bev_last_error(void);

/* Bytes of device workspace lm_bev_rasterize needs for n_points points (256-B aligned).
 * out: the output set the call will use (only which pointers are non-NULL matters: it fixes the
 * shared-memory tile size); NULL = an upper bound valid for every output set.            */
int lm_bev_workspace_bytes(const lm_bev_params *p, int64_t n_points, int algo,
                           const lm_bev_outputs *out, size_t *bytes);

/* Prepare a freshly allocated workspace (sized by lm_bev_workspace_bytes with the SAME p, n_points, algo,
 * out) for LM_ALGO_SWEEP: the sweep's mailboxes keep state between calls.  Enqueues one kernel; call it
 * once per workspace (again after the workspace memory was used for something else).  A workspace that
 * was never initialised is detected and simply always takes the fall-back.  No-op for other algos.  */
int lm_bev_workspace_init(const lm_bev_params *p, int64_t n_points, int algo, const lm_bev_outputs *out,
                          void *workspace_dev, size_t workspace_bytes, void *stream);

/* Byte offset, inside an LM_ALGO_SWEEP workspace of workspace_bytes bytes, of four uint32 {magic, cooldown, n_failed, n_ok}: how many
 * calls on this workspace were done by the sweep (n_ok) and how many fell back (n_failed); cooldown > 0 =
 * the sweep is switched off for that many further calls after a failure.  Diagnostics only.          */
int lm_bev_sweep_state_offset(size_t workspace_bytes, size_t *offset);

/* points_dev: n_points packed records (x, y, z, intensity) of 4 x f32 = one 16-byte float4
 * each, 16-byte aligned, in the raster's local frame (the LAS read offset and the sidecar
 * rotation/translation already removed; intensity = the LAS u16 value as a float, the
 * record convention of read_las, reference baseline/datasets/laserlane_proposals.py:618-636). */
int lm_bev_rasterize(const lm_bev_params *p, const float *points_dev, int64_t n_points, int algo,
                     void *workspace_dev, size_t workspace_bytes,
                     const lm_bev_outputs *out, void *stream);

/* Same call, restricted to some pipeline stages (for per-kernel timing with events between
 * the stages; running BIN, INDEX, REDUCE back to back on one stream == lm_bev_rasterize).
 * LM_ALGO_DIRECT: BIN = init + accumulate, REDUCE = finalize, INDEX = nothing.            */
/* LM_ALGO_SWEEP: SWEEP = the single-pass kernel (it zeroes the tables and leaves its verdict there, so a stage-split
 * sequence starts with it); BIN, INDEX, REDUCE = the fall-back kernels behind it.                          */
enum { LM_STAGE_BIN = 1, LM_STAGE_INDEX = 2, LM_STAGE_REDUCE = 4, LM_STAGE_SWEEP = 8, LM_STAGE_ALL = 15 };
int lm_bev_rasterize_stages(const lm_bev_params *p, const float *points_dev, int64_t n_points, int algo,
                            void *workspace_dev, size_t workspace_bytes,
                            const lm_bev_outputs *out, void *stream, int stages);

/* ---- Plan: a stream of equally-shaped calls.  Holds what is fixed (geometry, output set, algorithm, tuning) and,
 * with tuning.use_graph, a CUDA graph of the launch sequence that is replayed while a call repeats the previous
 * call's arguments (same buffers, same n_points: the steady state of a pipeline with preallocated buffers) -- one
 * graph launch instead of a memset and 4..6 kernel launches, which is what bounds the small configs.
 * The library reads no environment variables: every knob is a field here (0 = default).  A plan is used by one
 * thread at a time; it owns no device memory (a capture stream and the graph only).                          */
typedef struct lm_bev_tuning {
    int32_t bin_ctas_per_sm;   /* bin_points CTAs per SM (default: what fits, 3 on rasters wider than 36 tiles)  */
    int32_t red_ctas_per_sm;   /* reduce_tiles CTAs per SM (default: what fits)                                  */
    int32_t tile_h_log2;       /* shared-memory tile height 2^5..2^7 rows (default: by plane count / raster size) */
    int32_t max_tiles;         /* tiles per launch; smaller values force the in-call row-window loop (tests)     */
    int32_t stream_hint;       /* 1: the point stream is loaded with an L2 evict-first policy                    */
    int32_t use_graph;         /* 1: lm_bev_plan_rasterize replays a captured graph while the arguments repeat   */
    int32_t bin_compact_table; /* bin_points keeps its per-tile append state in a per-CTA hash table (4 CTAs per SM for any tile */
                               /* count): 0 = when direct indexing would run < 3 CTAs per SM, 1 = always, -1 = never         */
    int32_t reserved;
} lm_bev_tuning;
typedef struct lm_bev_plan lm_bev_plan;

/* out_set: which outputs the calls will request (pointer NULL-ness + acc_band; the pointers themselves are not kept). */
int lm_bev_plan_create(const lm_bev_params *p, int64_t max_points, int algo, const lm_bev_outputs *out_set,
                       const lm_bev_tuning *tuning /* NULL = defaults */, lm_bev_plan **plan);
int lm_bev_plan_workspace_bytes(const lm_bev_plan *plan, size_t *bytes);
int lm_bev_plan_init_workspace(lm_bev_plan *plan, void *workspace_dev, size_t workspace_bytes, void *stream);
int lm_bev_plan_rasterize(lm_bev_plan *plan, const float *points_dev, int64_t n_points, void *workspace_dev,
                          size_t workspace_bytes, const lm_bev_outputs *out, void *stream);
int lm_bev_plan_destroy(lm_bev_plan *plan);

/* Batched call (BASELINE.json configs[4]: on-the-fly rasterisation of a DataLoader batch into
 * sample['proj'], reference baseline/models/pcencoder/postprojector.py:79-82).  n_samples clouds
 * are rasterised onto n_samples rasters of the SAME shape p (height, width, resolutions, channels,
 * intensity range); sample s takes its origin and window shift from geoms[s] instead of p.  Up to
 * 32 samples share one set of launches (they are stacked along the rows internally), so a batch
 * of 1152^2 crops runs at the speed of one large raster instead of n_samples small ones.
 *   points_dev[s], n_points[s]   HOST arrays of device pointers / counts (one cloud per sample)
 *   out->image_dev  [n_samples][height][width][n_channels] u8
 *   out->count16_dev [n_samples][height][width] u16
 *   out->proj_dev   [n_samples][n_channels][height][width] f32   (= torch.stack of the loader's tensors)
 *   out->acc_dev    must be NULL
 * Results are bit-identical to n_samples separate lm_bev_rasterize calls.                        */
typedef struct lm_bev_sample_geom {
    float   bev_img_offset[2];
    float   local_min_ele;
    int32_t row0, col0;
    int32_t reserved;
} lm_bev_sample_geom;

int lm_bev_workspace_bytes_batch(const lm_bev_params *p, int32_t n_samples, int64_t n_points_total,
                                 const lm_bev_outputs *out, size_t *bytes);
int lm_bev_rasterize_batch(const lm_bev_params *p, int32_t n_samples, const lm_bev_sample_geom *geoms,
                           const float *const *points_dev, const int64_t *n_points,
                           void *workspace_dev, size_t workspace_bytes,
                           const lm_bev_outputs *out, void *stream);

/* dst = merge(dst, src) over rows [0,rows) of two accumulator sets whose planes are
 * dst_plane_stride / src_plane_stride ELEMENTS apart (count, sums: add; max: max; min: min).
 * This is the halo-merge law of strip sharding.                 */
int lm_bev_acc_merge(uint32_t *dst_acc_dev, int64_t dst_plane_stride,
                     const uint32_t *src_acc_dev, int64_t src_plane_stride,
                     int32_t rows, int32_t width, void *stream);

/* Accumulators -> requested outputs for rows [row_begin,row_end) of the window.  acc_dev
 * is [LM_ACC_PLANES][p->height][p->width]; outputs are full-window buffers.            */
int lm_bev_finalize(const lm_bev_params *p, const uint32_t *acc_dev,
                    int32_t row_begin, int32_t row_end, const lm_bev_outputs *out, void *stream);

/* The two calls above in one launch, for the halo band of a strip: rows [row_begin,row_end) of acc_dev
 * ([LM_ACC_PLANES][p->height][p->width]) <- merge with the neighbour's planes, then finished into out.
 * Only the planes whose bit is set in plane_mask (bit k = LM_ACC_* plane k) were sent: recv_dev is
 * [popcount(plane_mask)][row_end-row_begin][p->width], ascending plane order.                      */
int lm_bev_merge_finalize(const lm_bev_params *p, uint32_t *acc_dev, int32_t row_begin, int32_t row_end,
                          const uint32_t *recv_dev, int32_t plane_mask, const lm_bev_outputs *out, void *stream);

/* Cut a [height][width][c] u8 mosaic into non-overlapping tile x tile crops (row-major crop
 * order, ragged edges zero-filled = empty cells): crops_dev is [n_crops][tile][tile][c].
 * tile = 1152 for cropped_tiff (reference configs/Proj_polyline_fpn_vit_vertex_2.py:38).  */
int lm_bev_crop_tiles(const uint8_t *image_dev, int32_t height, int32_t width, int32_t c,
                      int32_t tile, uint8_t *crops_dev, void *stream);

/* Self-test of the kernels' division: they compute a / divisor (divisor = img_reso, ele_reso)
 * with a 3-operation exact sequence instead of the generic IEEE division.  This runs that
 * sequence against __fdiv_rn for ALL 2^32 binary32 dividends: out2_dev[0] = mismatches (must be
 * 0), out2_dev[1] = dividends covered by the fast path (the rest take __fdiv_rn itself).     */
int lm_bev_selftest_div(float divisor, unsigned long long *out2_dev, void *stream);

/* Self-test of the reduce kernel's mean: (sum + count/2) / count is computed with a float
 * reciprocal when count <= 4095.  Checks it against the integer division for every count in
 * [1, 4095] and every sum in [0, 255*count]: out2_dev[0] = mismatches (must be 0),
 * out2_dev[1] = pairs checked.                                                              */
int lm_bev_selftest_mean(unsigned long long *out2_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LM_BEV_H */
