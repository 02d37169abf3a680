/* lm_post.h -- C-ABI of the stages on either side of the rasteriser that the reference runs as
 * per-pixel / per-vertex Python loops.  Both have an in-tree
 * reference implementation, so parity is PINNED: tests/golden/ holds inputs and outputs produced
 * by running the reference's own functions (tests/golden/make_golden.py).
 *
 *   lm_bev_img2pc        BEV pixel polylines -> LAS world coordinates
 *                        = reference baseline/utils/coor_img2pc.py:127-183
 *                          (transform_coordinate_from_img_2_pc, with the roi branch of
 *                          modify_empty_pixel_elevation :97-122 and LeastSuqare :59-73)
 *   lm_label_endpoint_map  Gaussian endpoint heat map of one 1152 x 1152 label crop
 *                        = reference data/convert_data.py:248-317,357-361
 *                          (gaussian + get_endpoint_maps_per_batch(merge_endp_map=True) * 255)
 *   lm_label_polylines   the semantic / instance / orientation label rasters
 *                        = reference data/convert_data.py:319-356 (cv2.line, 1 px, 8-connected)
 *   lm_proj_color_jitter the loader's colour augmentation + normalisation on the GPU
 *                        = reference baseline/datasets/laserlane_proposals.py:95-96,255-264
 *
 * Conventions are those of lm_bev.h (caller-owned device memory, stream passed in, 0 / <0 / >0).
 */
#ifndef LM_POST_H
#define LM_POST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Sidecar values of one crop (reference baseline/utils/io_utils.py:125-150) as float64, plus the
 * two quaternion forms rotateByQuanternion3D derives on the host (coor_img2pc.py:38-53):
 * quat = las_rotation_trans_quan[3:7] (w x y z), quat_inv = conj(quat) / |quat| (sic: not |quat|^2). */
typedef struct lm_img2pc_params {
    double img_reso[2];
    double bev_img_offset[2];
    double ele_reso;
    double local_min_ele;
    double translation[3];      /* las_rotation_trans_quan[0:3] */
    double quat[4];
    double quat_inv[4];
    double las_read_offset[3];
} lm_img2pc_params;

/* A batch of n_crops crops of equal shape; one CTA per crop.
 *   images_dev  [n_crops][height][width][channels] u8, IN/OUT: the elevation (channel 1) of empty
 *               pixels under polyline vertices is filled in place with the mean over the smallest
 *               window that holds a non-empty pixel, in vertex order -- exactly what the reference
 *               does to its working copy (coor_img2pc.py:97-122), including the uint8 truncation.
 *   seqs_dev    [n_crops][n_lines][max_len][2] f64 (row, col) pixel coordinates, zero padded
 *   lens_dev    [n_crops][n_lines] i32 vertices per polyline (0 = unused line)
 *   params_dev  [n_crops]
 *   world_dev   [n_crops][n_lines][max_len][3] f64, every entry written (padding included, as upstream)
 * Vertices must lie inside the image (the reference raises IndexError otherwise; here they are
 * clamped).  A vertex on an empty pixel of an all-empty image is left unfilled (upstream loops forever). */
int lm_bev_img2pc(uint8_t *images_dev, int32_t n_crops, int32_t height, int32_t width, int32_t channels,
                  const double *seqs_dev, const int32_t *lens_dev, int32_t n_lines, int32_t max_len,
                  const lm_img2pc_params *params_dev, double *world_dev, void *stream);

/* Endpoint heat map of a label crop: out[r][c] = saturate_u8(rint(255 * max over lanes and over the
 * lane's two end points (r0, c0) of exp(-((r-r0)^2 + (c-c0)^2) / (2 sigma^2)))), sigma = 2, for end
 * points strictly inside the 20 px border; the end-point pixels themselves are exactly 255.
 *   starts_dev, ends_dev  [n_lines][2] f64 (row, col) first / last vertex of each lane
 *   out_dev               [height][width] u8  (the label PNG labels/sparse_endp/<stem>.png)
 * The 50 grey levels that are not zero (d^2 <= 49) are computed on the HOST with libm's exp, exactly
 * as the reference's math.exp / np.float32 / cv2 saturate_cast chain does, and handed to the kernel as
 * a table: parity is exact, no device transcendental is involved.                               */
int lm_label_endpoint_map(const double *starts_dev, const double *ends_dev, int32_t n_lines,
                          int32_t height, int32_t width, uint8_t *out_dev, void *stream);

/* The three polyline label rasters.  Segments are drawn in order with cv2.line semantics (8-connected
 * Bresenham, both end points included, clipped to the image); where segments overlap the LAST one wins,
 * as in the sequential reference.
 *   seqs_dev      [n_lines][max_len][2] f64 (row, col); truncated to int like tuple(map(int, pt))
 *   lens_dev      [n_lines] i32
 *   semantic_dev  [n_lines] i32 lane class (1 -> 128, else 255, convert_data.py:331-334)
 *   instance_dev  [n_lines] i32 instance id
 *   orient_dev    [n_lines][max_len] i32 orientation bin of segment (v, v+1)
 *   out_*         [height][width] u8 each; scratch_dev: [height][width] u32 workspace            */
int lm_label_polylines(const double *seqs_dev, const int32_t *lens_dev, const int32_t *semantic_dev,
                       const int32_t *instance_dev, const int32_t *orient_dev, int32_t n_lines, int32_t max_len,
                       int32_t height, int32_t width, uint8_t *out_semantic_dev, uint8_t *out_instance_dev,
                       uint8_t *out_orient_dev, uint32_t *scratch_dev, void *stream);

/* Loader fusion: the colour augmentation the reference applies to the
 * loaded image on a DataLoader worker (baseline/datasets/laserlane_proposals.py:95-96,255-264:
 * torchvision ColorJitter(brightness, contrast, saturation) in a random order, then
 * Normalize(mean=[0.5], std=[0.5])), applied in place to the rasteriser's proj tensor on the GPU.
 * The random draws stay on the host (torchvision ColorJitter.get_params: same RNG stream as the
 * reference); this call only does the arithmetic, in float32 and in torchvision's operation order:
 *   gray = 0.2989 r + 0.587 g + 0.114 b
 *   brightness: x = clamp(b x, 0, 1)      saturation: x = clamp(s x + (1 - s) gray, 0, 1)
 *   contrast:   x = clamp(c x + (1 - c) mean(gray), 0, 1)       normalise: x = (x - mean) / std
 * Parity: float32, within 1e-6 absolute of torchvision on the CPU (the only difference is the
 * summation order of mean(gray)).                                                                 */
typedef struct lm_jitter {
    int32_t order[4];      /* permutation of {0 brightness, 1 contrast, 2 saturation, 3 hue (ignored)} */
    float   brightness, contrast, saturation;   /* factors; < 0 = skip that operation             */
    float   reserved;
} lm_jitter;

#define LM_JITTER_PARTIALS 64
/* proj_dev [n_samples][3][height][width] f32 in place; jitter [n_samples] HOST array;
 * scratch_dev [n_samples][LM_JITTER_PARTIALS] f64 device workspace.                              */
int lm_proj_color_jitter(float *proj_dev, int32_t n_samples, int32_t height, int32_t width,
                         const lm_jitter *jitter, float norm_mean, float norm_std,
                         double *scratch_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LM_POST_H */
