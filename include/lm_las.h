/* lm_las.h -- C-ABI of the LAS-record front end of liblm_bev.so (SURVEY.md section 8f, rank 2:
 * "LAS decode on GPU").
 *
 * What this boundary replaces.  The reference reads clouds with laspy on the host:
 *   las = laspy.read(path); np.stack((las.x, las.y, las.z)); las.intensity
 * (reference baseline/datasets/laserlane_proposals.py:618-636, same code in
 * laserlane_proposals_ego.py:622-640), i.e. it scales every int32 coordinate to float64 on
 * one CPU core and stacks four float64 columns.  Here the point-data block of the
 * (uncompressed) LAS file is copied to the device as it lies on disk and decoded there:
 * either to packed float4 records (lm_las_decode) or straight into the rasteriser, with no
 * float copy of the cloud in HBM at all (lm_bev_rasterize_las).
 *
 * The arithmetic, one IEEE binary64 operation per step (no FMA), in this order:
 *   world = X * scale + offset                   laspy's las.x / las.y / las.z
 *   d     = (world - las_read_offset) - t        inverse of reference
 *                                                baseline/utils/coor_img2pc.py:172,175-177
 *   p     = rot * d,  p[k] = (rot[3k]*d0 + rot[3k+1]*d1) + rot[3k+2]*d2
 *                                                rot = R(q)^T: inverse of the quaternion
 *                                                rotation of coor_img2pc.py:163-171
 *   out   = (float)p.x, (float)p.y, (float)p.z, (float)intensity_u16
 * las_read_offset, t (= las_rotation_trans_quan[0:3]) and q (= [3:7], w x y z) are the sidecar
 * values of cropped_tiff_param/<stem>.txt (reference baseline/utils/io_utils.py:125-150).
 * The caller converts q to the row-major matrix rot once on the host
 * (lanemapping_b200/sidecar.py::quat_to_matrix(q).T); identity = {1,0,0, 0,1,0, 0,0,1}.
 *
 * Conventions are those of lm_bev.h: caller-owned device memory, nothing allocated, nothing
 * synchronised, kernels enqueued on the stream passed in, 0 / negative / positive returns.
 */
#ifndef LM_LAS_H
#define LM_LAS_H

#include <stddef.h>
#include <stdint.h>

#include "lm_bev.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lm_las_xform {
    int32_t record_length;      /* bytes per point data record (LAS header offset 105), 14..100 */
    int32_t reserved;
    double  scale[3];           /* LAS header X/Y/Z scale factor  (offset 131)                  */
    double  offset[3];          /* LAS header X/Y/Z offset        (offset 155)                  */
    double  las_read_offset[3]; /* sidecar                                                      */
    double  translation[3];     /* sidecar las_rotation_trans_quan[0:3]                         */
    double  rot[9];             /* row-major R(q)^T                                             */
} lm_las_xform;

/* records_dev: n_points * record_length bytes, 16-byte aligned; the buffer must be readable up to
 * the next multiple of 16 bytes past its end (any cudaMalloc / torch allocation is).
 * points_dev:  n_points packed float4 (x, y, z, intensity), 16-byte aligned: exactly the input
 * of lm_bev_rasterize.                                                                          */
int lm_las_decode(const uint8_t *records_dev, int64_t n_points, const lm_las_xform *x,
                  float *points_dev, void *stream);

/* lm_bev_rasterize (LM_ALGO_BINNED) with the decode fused into the first kernel: the records are
 * staged in shared memory by TMA bulk copies and decoded there.  Same workspace size
 * (lm_bev_workspace_bytes(p, n_points, LM_ALGO_BINNED, out, ...)), same outputs, bit-identical to
 * lm_las_decode followed by lm_bev_rasterize.  LM_ERR_UNSUPPORTED if the staged records do not fit
 * in shared memory next to the tile state (very long records or > ~10000 tiles): decode first. */
int lm_bev_rasterize_las(const lm_bev_params *p, const uint8_t *records_dev, int64_t n_points,
                         const lm_las_xform *x, void *workspace_dev, size_t workspace_bytes,
                         const lm_bev_outputs *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LM_LAS_H */
