// lm_post.cu -- the stages around the rasteriser that the reference runs as Python loops
// (include/lm_post.h): BEV pixel polylines -> LAS world coordinates, and the label rasters.
// All float arithmetic is binary64 in the reference's own operation order (compile with
// -fmad=false), so the results match numpy bit for bit; everything else is integer.
#include "lm_post.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "lm_host.h"

namespace {

constexpr int POST_THREADS = 256;

// ------------------------------------------------------------------------------------------
// lm_bev_img2pc  (reference baseline/utils/coor_img2pc.py:127-183)
// ------------------------------------------------------------------------------------------
// Python int() / ndarray.astype(int) of a float64 pixel coordinate: truncation toward zero
__device__ __forceinline__ int trunc_px(double v, int n) {
    const int i = (int)v;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

struct Quat {
    double w, x, y, z;
};
// multiplyQuanternion, coor_img2pc.py:22-30 (each line evaluated left to right)
__device__ __forceinline__ Quat quat_mul(const Quat a, const Quat b) {
    Quat o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return o;
}

__global__ void __launch_bounds__(POST_THREADS) img2pc_kernel(uint8_t *images, int H, int W, int C,
                                                              const double *__restrict__ seqs,
                                                              const int *__restrict__ lens, int L, int V,
                                                              const lm_img2pc_params *__restrict__ params,
                                                              double *__restrict__ world) {
    const int crop = blockIdx.x, tid = threadIdx.x;
    uint8_t *img = images + (size_t)crop * H * W * C;
    const double *sq = seqs + (size_t)crop * L * V * 2;
    const int *ln = lens + (size_t)crop * L;
    double *out = world + (size_t)crop * L * V * 3;
    const lm_img2pc_params P = params[crop];
    __shared__ unsigned long long s_tot[POST_THREADS / 32], s_g[POST_THREADS / 32];
    __shared__ unsigned int s_val[POST_THREADS / 32];

    // ---- A: modify_empty_pixel_elevation, roi branch (:97-122).  Sequential over the vertices (a
    //      filled pixel is non-empty for every later search), cooperative inside one search.
    for (int l = 0; l < L; ++l) {
        const int n = min(max(ln[l], 0), V);
        for (int k = 0; k < n; ++k) {
            const int ph = trunc_px(sq[(l * V + k) * 2 + 0], H), pw = trunc_px(sq[(l * V + k) * 2 + 1], W);
            unsigned int s = 0;
            for (int c = 0; c < C; ++c) s += img[((size_t)ph * W + pw) * C + c];
            if ((ph == 0 && pw == 0) || s > 1u) continue;                     // :106 (uniform over the CTA)
            for (int step = 1;; ++step) {
                const int r0 = max(ph - step, 0), r1 = min(ph + step, H);       // [pt-step, pt+step): as upstream
                const int c0 = max(pw - step, 0), c1 = min(pw + step, W);
                const int ww = c1 - c0, cnt = (r1 - r0) * ww;
                unsigned long long tot = 0, g = 0;
                unsigned int val = 0;
                for (int i = tid; i < cnt; i += POST_THREADS) {
                    const uint8_t *px = img + ((size_t)(r0 + i / ww) * W + (c0 + i % ww)) * C;
                    unsigned int ps = 0;
                    for (int c = 0; c < C; ++c) ps += px[c];
                    tot += ps;
                    val += ps > 0u;
                    g += px[1];
                }
                for (int o = 16; o; o >>= 1) {
                    tot += __shfl_xor_sync(0xffffffffu, tot, o);
                    g += __shfl_xor_sync(0xffffffffu, g, o);
                    val += __shfl_xor_sync(0xffffffffu, val, o);
                }
                if ((tid & 31) == 0) { s_tot[tid >> 5] = tot; s_g[tid >> 5] = g; s_val[tid >> 5] = val; }
                __syncthreads();
                tot = 0; g = 0; val = 0;
                for (int w = 0; w < POST_THREADS / 32; ++w) { tot += s_tot[w]; g += s_g[w]; val += s_val[w]; }
                __syncthreads();
                if (tot > 0) {
                    // img[pt_h, pt_w, 1] = sum(G) / valid_num : float64 quotient stored into uint8 (truncates)
                    if (tid == 0) img[((size_t)ph * W + pw) * C + 1] = (uint8_t)((double)g / (double)val);
                    break;
                }
                if (r0 == 0 && c0 == 0 && r1 == H && c1 == W) break;           // all-empty image: upstream never returns
            }
            __syncthreads();                                                   // the fill is visible to the next search
        }
    }
    __syncthreads();

    // ---- B: pixel -> local frame (:136-139, :150), every entry incl. the zero padding
    for (int i = tid; i < L * V; i += POST_THREADS) {
        const double r = sq[i * 2 + 0], c = sq[i * 2 + 1];
        const int ph = trunc_px(r, H), pw = trunc_px(c, W);
        out[i * 3 + 0] = r * P.img_reso[0] + P.bev_img_offset[0];
        out[i * 3 + 1] = c * P.img_reso[1] + P.bev_img_offset[1];
        out[i * 3 + 2] = (double)img[((size_t)ph * W + pw) * C + 1] * P.ele_reso + P.local_min_ele;
    }
    __syncthreads();

    // ---- C: least-squares line through each polyline's elevations (:154-159, LeastSuqare :59-73):
    //      Python's sum() adds left to right, so one thread per polyline does the same
    for (int l = tid; l < L; l += POST_THREADS) {
        const int n = min(max(ln[l], 0), V);
        if (n <= 0) continue;
        double *z = out + (size_t)l * V * 3 + 2;
        double sxy = 0.0, sy = 0.0;
        long long sx = 0, sxx = 0;
        for (int k = 0; k < n; ++k) {
            const double y = z[k * 3];
            sxy = sxy + (double)k * y;
            sy = sy + y;
            sx += k;
            sxx += (long long)k * k;
        }
        const double p = (double)n * sxy - (double)sx * sy;
        const long long q = (long long)n * sxx - sx * sx;
        const double w = q == 0 ? 0.0 : p / (double)q;                         // abs(q) < EPS on an integer
        double sb = 0.0;
        for (int k = 0; k < n; ++k) sb = sb + (z[k * 3] - w * (double)k);
        const double b = sb / (double)n;
        for (int k = 0; k < n; ++k) z[k * 3] = w * (double)k + b;
    }
    __syncthreads();

    // ---- D: rotate by the quaternion, add the translation, add the LAS read offset (:163-177)
    const Quat q = {P.quat[0], P.quat[1], P.quat[2], P.quat[3]};
    const Quat qi = {P.quat_inv[0], P.quat_inv[1], P.quat_inv[2], P.quat_inv[3]};
    for (int i = tid; i < L * V; i += POST_THREADS) {
        const Quat v = {0.0, out[i * 3 + 0], out[i * 3 + 1], out[i * 3 + 2]};
        const Quat r = quat_mul(quat_mul(q, v), qi);
        out[i * 3 + 0] = (r.x + P.translation[0]) + P.las_read_offset[0];
        out[i * 3 + 1] = (r.y + P.translation[1]) + P.las_read_offset[1];
        out[i * 3 + 2] = (r.z + P.translation[2]) + P.las_read_offset[2];
    }
}

// ------------------------------------------------------------------------------------------
// lm_label_endpoint_map  (reference data/convert_data.py:248-317, 357-361)
// ------------------------------------------------------------------------------------------
constexpr int ENDP_CLIP = 20;            // clip_width = kernel_size * 5, kernel_size = 4
constexpr int ENDP_LUT = 64;             // grey level by squared distance; zero from d^2 = 50 on
struct EndpLut {
    uint8_t v[ENDP_LUT];
};

__global__ void __launch_bounds__(POST_THREADS) endpoint_map_kernel(const double *__restrict__ starts,
                                                                    const double *__restrict__ ends, int n_lines,
                                                                    int H, int W, EndpLut lut, uint8_t *__restrict__ out) {
    extern __shared__ int s_pts[];       // [2 * n_lines][2] integer end points, (-1, -1) = not drawn
    for (int i = threadIdx.x; i < 2 * n_lines; i += blockDim.x) {
        const int l = i >> 1;
        const double a0 = starts[l * 2], a1 = starts[l * 2 + 1], b0 = ends[l * 2], b1 = ends[l * 2 + 1];
        const bool lane = !(fabs(b0 - a0) < 1e-3 && fabs(b1 - a1) < 1e-3);          // :267-270 "no lane instance"
        const double p0 = (i & 1) ? b0 : a0, p1 = (i & 1) ? b1 : a1;
        const bool inside = p0 > ENDP_CLIP && p0 < (H - ENDP_CLIP) && p1 > ENDP_CLIP && p1 < (W - ENDP_CLIP);
        s_pts[i * 2 + 0] = lane && inside ? (int)p0 : -1;
        s_pts[i * 2 + 1] = lane && inside ? (int)p1 : -1;
    }
    __syncthreads();
    const size_t total = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / W), c = (int)(i % W);
        unsigned int best = 0;
        for (int k = 0; k < 2 * n_lines; ++k) {
            const int pr = s_pts[k * 2], pc = s_pts[k * 2 + 1];
            if (pr < 0) continue;
            const int dr = r - pr, dc = c - pc;
            if (abs(dr) > 7 || abs(dc) > 7) continue;
            const int d2 = dr * dr + dc * dc;
            if (d2 < ENDP_LUT) best = max(best, (unsigned int)lut.v[d2]);
        }
        out[i] = (uint8_t)best;
    }
}

// ------------------------------------------------------------------------------------------
// lm_label_polylines  (reference data/convert_data.py:319-356)
// ------------------------------------------------------------------------------------------
// one thread per segment walks cv::LineIterator (8-connected, left to right) and stamps its drawing
// order into the scratch raster; the later segment wins, as when the lines are drawn one by one
__global__ void __launch_bounds__(POST_THREADS) polyline_stamp_kernel(const double *__restrict__ seqs,
                                                                      const int *__restrict__ lens, int L, int V,
                                                                      int H, int W, uint32_t *__restrict__ scratch) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L * V) return;
    const int l = s / V, v = s - l * V;
    if (v + 1 >= min(max(lens[l], 0), V)) return;
    // pt = tuple(map(int, pt[::-1])): x = col, y = row, truncated toward zero
    int x1 = (int)seqs[(l * V + v) * 2 + 1], y1 = (int)seqs[(l * V + v) * 2 + 0];
    const int x2 = (int)seqs[(l * V + v + 1) * 2 + 1], y2 = (int)seqs[(l * V + v + 1) * 2 + 0];
    int dx = x2 - x1, dy = y2 - y1, delta_x = 1, delta_y = 1;
    if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }       // left_to_right: start from the left end point
    if (dy < 0) { dy = -dy; delta_y = -1; }
    const bool vert = dy > dx;
    if (vert) { int t = dx; dx = dy; dy = t; t = delta_x; delta_x = delta_y; delta_y = t; }
    int err = dx - (dy + dy);
    const int plus_delta = dx + dx, minus_delta = -(dy + dy);
    // major-axis move every step, minor-axis move when err < 0
    const int major_x = vert ? 0 : delta_x, major_y = vert ? delta_x : 0;
    const int minor_x = vert ? delta_y : 0, minor_y = vert ? 0 : delta_y;
    const uint32_t stamp = (uint32_t)s + 1u;
    int x = x1, y = y1;
    for (int i = 0; i <= dx; ++i) {
        if (x >= 0 && x < W && y >= 0 && y < H) atomicMax(&scratch[(size_t)y * W + x], stamp);
        const int mask = err < 0 ? -1 : 0;
        err += minus_delta + (plus_delta & mask);
        x += major_x + (minor_x & mask);
        y += major_y + (minor_y & mask);
    }
}

__device__ __forceinline__ uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void __launch_bounds__(POST_THREADS) polyline_resolve_kernel(const uint32_t *__restrict__ scratch,
                                                                        const int *__restrict__ semantic,
                                                                        const int *__restrict__ instance,
                                                                        const int *__restrict__ orient, int V, size_t total,
                                                                        uint8_t *__restrict__ o_sem, uint8_t *__restrict__ o_ins,
                                                                        uint8_t *__restrict__ o_ori) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t st = scratch[i];
        uint8_t a = 0, b = 0, c = 0;
        if (st) {
            const int s = (int)(st - 1u), l = s / V;
            a = semantic[l] == 1 ? 128 : 255;                     // convert_data.py:331-334
            b = sat_u8(instance[l]);                              // cv2 colour scalar: saturate_cast<uchar>
            c = sat_u8(orient[s]);
        }
        if (o_sem) o_sem[i] = a;
        if (o_ins) o_ins[i] = b;
        if (o_ori) o_ori[i] = c;
    }
}

// ------------------------------------------------------------------------------------------
// lm_proj_color_jitter  (reference baseline/datasets/laserlane_proposals.py:255-264; torchvision
// transforms/_functional_tensor.py: adjust_brightness / adjust_contrast / adjust_saturation / _blend)
// ------------------------------------------------------------------------------------------
constexpr int JIT_MAX = 32;              // samples per launch (the factors travel as kernel parameters)
struct JitTab {
    int order[JIT_MAX][4];
    float b[JIT_MAX], c[JIT_MAX], s[JIT_MAX];
    float omc[JIT_MAX], oms[JIT_MAX];    // (float)(1.0 - factor), the subtraction done in binary64 like Python
};

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ float gray_of(float r, float g, float b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
// ratio * img1 + (1 - ratio) * img2, clamped
__device__ __forceinline__ float blend(float x, float ratio, float one_minus, float other) {
    return clamp01(__fadd_rn(__fmul_rn(ratio, x), __fmul_rn(one_minus, other)));
}
// the operations of `order` from position `from` up to (not including) the contrast step, or all of
// them when the mean is known; returns the position of the contrast step it stopped at (4 = done)
__device__ __forceinline__ int jitter_ops(const JitTab &t, int smp, int from, bool have_mean, float mean, float &r, float &g,
                                          float &b) {
    for (int k = from; k < 4; ++k) {
        const int op = t.order[smp][k];
        if (op == 0 && t.b[smp] >= 0.0f) {
            r = blend(r, t.b[smp], 0.0f, 0.0f); g = blend(g, t.b[smp], 0.0f, 0.0f); b = blend(b, t.b[smp], 0.0f, 0.0f);
        } else if (op == 1 && t.c[smp] >= 0.0f) {
            if (!have_mean) return k;
            r = blend(r, t.c[smp], t.omc[smp], mean); g = blend(g, t.c[smp], t.omc[smp], mean); b = blend(b, t.c[smp], t.omc[smp], mean);
        } else if (op == 2 && t.s[smp] >= 0.0f) {
            const float y = gray_of(r, g, b);
            r = blend(r, t.s[smp], t.oms[smp], y); g = blend(g, t.s[smp], t.oms[smp], y); b = blend(b, t.s[smp], t.oms[smp], y);
        }
    }
    return 4;
}

// pass 1: per-sample partial sums of gray(state before the contrast step), fixed slices, binary64
__global__ void __launch_bounds__(POST_THREADS) jitter_mean_kernel(const float *__restrict__ proj, size_t cells,
                                                                   const __grid_constant__ JitTab tab, double *__restrict__ partial) {
    const int smp = blockIdx.y, part = blockIdx.x;
    const float *base = proj + (size_t)smp * 3 * cells;
    const size_t per = (cells + LM_JITTER_PARTIALS - 1) / LM_JITTER_PARTIALS;
    const size_t lo = part * per, hi = lo + per < cells ? lo + per : cells;
    double acc = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += POST_THREADS) {
        float r = base[i], g = base[cells + i], b = base[2 * cells + i];
        jitter_ops(tab, smp, 0, false, 0.0f, r, g, b);
        acc += (double)gray_of(r, g, b);
    }
    __shared__ double s_acc[POST_THREADS];
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    for (int o = POST_THREADS / 2; o; o >>= 1) {          // fixed tree: deterministic
        if ((int)threadIdx.x < o) s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)smp * LM_JITTER_PARTIALS + part] = s_acc[0];
}

// pass 2: all operations in order, then the normalisation, in place
__global__ void __launch_bounds__(POST_THREADS) jitter_apply_kernel(float *__restrict__ proj, size_t cells,
                                                                    const __grid_constant__ JitTab tab,
                                                                    const double *__restrict__ partial, float nmean, float nstd) {
    const int smp = blockIdx.y;
    float *base = proj + (size_t)smp * 3 * cells;
    double sum = 0.0;
    for (int k = 0; k < LM_JITTER_PARTIALS; ++k) sum += partial[(size_t)smp * LM_JITTER_PARTIALS + k];
    const float mean = (float)(sum / (double)cells);
    for (size_t i = blockIdx.x * (size_t)POST_THREADS + threadIdx.x; i < cells; i += (size_t)gridDim.x * POST_THREADS) {
        float r = base[i], g = base[cells + i], b = base[2 * cells + i];
        jitter_ops(tab, smp, 0, true, mean, r, g, b);
        base[i] = __fdiv_rn(__fsub_rn(r, nmean), nstd);
        base[cells + i] = __fdiv_rn(__fsub_rn(g, nmean), nstd);
        base[2 * cells + i] = __fdiv_rn(__fsub_rn(b, nmean), nstd);
    }
}

int failf(int code, const char *msg) { return lm_fail_msg(code, msg); }

}  // namespace

extern "C" {

int lm_bev_img2pc(uint8_t *images_dev, int32_t n_crops, int32_t height, int32_t width, int32_t channels,
                  const double *seqs_dev, const int32_t *lens_dev, int32_t n_lines, int32_t max_len,
                  const lm_img2pc_params *params_dev, double *world_dev, void *stream) {
    if (n_crops < 0 || height <= 0 || width <= 0 || n_lines < 0 || max_len < 0) return failf(-1, "img2pc: negative size");
    if (channels < 2 || channels > 4) return failf(-1, "img2pc: channels must be 2..4 (index 1 is the elevation)");
    if (n_crops == 0 || n_lines == 0 || max_len == 0) return 0;
    if (!images_dev || !seqs_dev || !lens_dev || !params_dev || !world_dev) return failf(-1, "img2pc: NULL buffer");
    if ((long long)n_lines * max_len >= (1ll << 28)) return failf(-3, "img2pc: more than 2^28 vertices per crop");
    img2pc_kernel<<<n_crops, POST_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        images_dev, height, width, channels, seqs_dev, lens_dev, n_lines, max_len, params_dev, world_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : lm_cuda_fail(e, "img2pc launch");
}

int lm_label_endpoint_map(const double *starts_dev, const double *ends_dev, int32_t n_lines, int32_t height,
                          int32_t width, uint8_t *out_dev, void *stream) {
    if (n_lines < 0 || height <= 0 || width <= 0) return failf(-1, "endpoint_map: negative size");
    if (!out_dev || (n_lines > 0 && (!starts_dev || !ends_dev))) return failf(-1, "endpoint_map: NULL buffer");
    if (n_lines > 4096) return failf(-3, "endpoint_map: more than 4096 lanes");
    // grey level by squared distance, computed with the host libm exactly as the reference does:
    // np.float32(math.exp(-d2 / (2 * sigma**2))) * 255 -> cv2's saturate_cast<uchar> (round half to even)
    EndpLut lut;
    const double sigma = 4 / 2.0;
    for (int d2 = 0; d2 < ENDP_LUT; ++d2) {
        const float h = (float)exp(-(double)d2 / (2 * sigma * sigma));
        const double v = nearbyint((double)h * 255.0);
        lut.v[d2] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    if (lut.v[50] != 0) return failf(-3, "endpoint_map: heat map support assumption violated");
    const size_t total = (size_t)height * width;
    int grid = (int)((total + POST_THREADS - 1) / POST_THREADS);
    const int cap = lm_sm_count() * 8;
    if (grid > cap) grid = cap;
    endpoint_map_kernel<<<grid, POST_THREADS, (size_t)(n_lines > 0 ? n_lines : 1) * 4 * sizeof(int),
                          reinterpret_cast<cudaStream_t>(stream)>>>(starts_dev, ends_dev, n_lines, height, width, lut, out_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : lm_cuda_fail(e, "endpoint_map launch");
}

int lm_label_polylines(const double *seqs_dev, const int32_t *lens_dev, const int32_t *semantic_dev,
                       const int32_t *instance_dev, const int32_t *orient_dev, int32_t n_lines, int32_t max_len,
                       int32_t height, int32_t width, uint8_t *out_semantic_dev, uint8_t *out_instance_dev,
                       uint8_t *out_orient_dev, uint32_t *scratch_dev, void *stream) {
    if (n_lines < 0 || max_len < 0 || height <= 0 || width <= 0) return failf(-1, "polylines: negative size");
    if (!scratch_dev || (!out_semantic_dev && !out_instance_dev && !out_orient_dev)) return failf(-1, "polylines: NULL buffer");
    if (n_lines > 0 && max_len > 0 && (!seqs_dev || !lens_dev || !semantic_dev || !instance_dev || !orient_dev))
        return failf(-1, "polylines: NULL table");
    if ((long long)n_lines * max_len >= (1ll << 31) - 1) return failf(-3, "polylines: too many segments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t total = (size_t)height * width;
    cudaError_t e = cudaMemsetAsync(scratch_dev, 0, total * sizeof(uint32_t), st);
    if (e != cudaSuccess) return lm_cuda_fail(e, "polylines memset");
    const int segs = n_lines * max_len;
    if (segs > 0)
        polyline_stamp_kernel<<<(segs + POST_THREADS - 1) / POST_THREADS, POST_THREADS, 0, st>>>(seqs_dev, lens_dev, n_lines,
                                                                                              max_len, height, width, scratch_dev);
    int grid = (int)((total + POST_THREADS - 1) / POST_THREADS);
    const int cap = lm_sm_count() * 8;
    if (grid > cap) grid = cap;
    polyline_resolve_kernel<<<grid, POST_THREADS, 0, st>>>(scratch_dev, semantic_dev, instance_dev, orient_dev,
                                                          max_len > 0 ? max_len : 1, total, out_semantic_dev,
                                                          out_instance_dev, out_orient_dev);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : lm_cuda_fail(e, "polylines launch");
}

int lm_proj_color_jitter(float *proj_dev, int32_t n_samples, int32_t height, int32_t width, const lm_jitter *jitter,
                         float norm_mean, float norm_std, double *scratch_dev, void *stream) {
    if (n_samples < 0 || height <= 0 || width <= 0) return failf(-1, "color_jitter: negative size");
    if (n_samples == 0) return 0;
    if (!proj_dev || !jitter || !scratch_dev) return failf(-1, "color_jitter: NULL buffer");
    if (!(norm_std > 0.0f)) return failf(-1, "color_jitter: norm_std must be positive");
    for (int s = 0; s < n_samples; ++s) {
        int seen = 0;
        for (int k = 0; k < 4; ++k) {
            if (jitter[s].order[k] < 0 || jitter[s].order[k] > 3) return failf(-1, "color_jitter: order must be a permutation of 0..3");
            seen |= 1 << jitter[s].order[k];
        }
        if (seen != 15) return failf(-1, "color_jitter: order must be a permutation of 0..3");
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t cells = (size_t)height * width;
    int gx = (int)((cells + POST_THREADS - 1) / POST_THREADS);
    const int cap = lm_sm_count() * 8;
    if (gx > cap) gx = cap;
    for (int s0 = 0; s0 < n_samples; s0 += JIT_MAX) {
        const int nb = n_samples - s0 < JIT_MAX ? n_samples - s0 : JIT_MAX;
        JitTab tab = {};
        for (int s = 0; s < nb; ++s) {
            const lm_jitter &j = jitter[s0 + s];
            for (int k = 0; k < 4; ++k) tab.order[s][k] = j.order[k];
            tab.b[s] = j.brightness; tab.c[s] = j.contrast; tab.s[s] = j.saturation;
            tab.omc[s] = (float)(1.0 - (double)j.contrast);
            tab.oms[s] = (float)(1.0 - (double)j.saturation);
        }
        float *p = proj_dev + (size_t)s0 * 3 * cells;
        double *part = scratch_dev + (size_t)s0 * LM_JITTER_PARTIALS;
        jitter_mean_kernel<<<dim3(LM_JITTER_PARTIALS, nb), POST_THREADS, 0, st>>>(p, cells, tab, part);
        jitter_apply_kernel<<<dim3(gx, nb), POST_THREADS, 0, st>>>(p, cells, tab, part, norm_mean, norm_std);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : lm_cuda_fail(e, "color_jitter launch");
}

}  // extern "C"
