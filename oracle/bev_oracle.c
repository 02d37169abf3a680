/* CPU oracle for the BEV rasterisation hot path, plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent restatement of the frozen forward spec (DESIGN.md section 2,
 * lanemapping_b200/spec.py) beside oracle/bev_oracle.py: one scalar loop over the points, no
 * numpy idiom shared with the Python oracle, so that an error in either shows up as a mismatch
 * between them (tests/test_c_oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library; the product (lanemapping_b200) never does.
 *
 * PARITY UNPINNED upstream: the reference has no forward rasteriser (README.md:171-172 defers to an
 * external tool).  What constrains the arithmetic below, all in /root/reference:
 *   baseline/utils/coor_img2pc.py:136-139   X = row*reso0 + off0, Y = col*reso1 + off1  (row <-> x)
 *   baseline/utils/coor_img2pc.py:150       Z = G*ele_reso + local_min_ele              (inverse of zq)
 *   baseline/utils/coor_img2pc.py:78,106    empty cell <=> all-zero pixel
 *   baseline/datasets/laserlane_proposals.py:626-628   intensity clip [800, 33000]
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off, no fast-math: every float operation below is
 * one IEEE binary32 operation, exactly as in the CUDA kernel and the numpy oracle).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    int32_t height, width, row0, col0;
    float off0, off1, reso0, reso1, local_min_ele, ele_reso;
    int32_t inten_min, inten_max;
} lmo_spec;

enum { ACC_COUNT = 0, ACC_SUM_I, ACC_SUM_Z, ACC_MAX_I, ACC_MIN_Z, ACC_MAX_Z, ACC_PLANES };
enum { CH_MAX_I = 0, CH_MEAN_I, CH_MIN_Z, CH_MAX_Z, CH_MEAN_Z, CH_DENSITY };

int lmo_abi_version(void) { return 1; }

/* acc: uint32 [6][H][W], zero-initialised here (min_z plane to 0xFFFFFFFF).  Returns the number
 * of points that fell inside the window. */
int64_t lmo_accumulate(const float *pts, int64_t n, const lmo_spec *s, uint32_t *acc)
{
    const int64_t cells = (int64_t)s->height * s->width;
    memset(acc, 0, sizeof(uint32_t) * ACC_PLANES * (size_t)cells);
    memset(acc + ACC_MIN_Z * cells, 0xFF, sizeof(uint32_t) * (size_t)cells);
    const float lo_r = (float)s->row0, hi_r = (float)(s->row0 + s->height);
    const float lo_c = (float)s->col0, hi_c = (float)(s->col0 + s->width);
    const int32_t span = s->inten_max - s->inten_min;
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float x = pts[4 * i], y = pts[4 * i + 1], z = pts[4 * i + 2], in = pts[4 * i + 3];
        /* step 1: keys.  sub, div, floor -- each one binary32 operation; NaN fails every test */
        const float dx = x - s->off0, dy = y - s->off1;
        const float rf = floorf(dx / s->reso0), cf = floorf(dy / s->reso1);
        if (!(rf >= lo_r && rf < hi_r && cf >= lo_c && cf < hi_c)) continue;
        const int64_t cell = (int64_t)((int32_t)rf - s->row0) * s->width + ((int32_t)cf - s->col0);
        /* step 2: height, round half to even (default rounding mode), NaN -> 0, clamp [0, 255] */
        const float dz = z - s->local_min_ele;
        const float zf = nearbyintf(dz / s->ele_reso);
        uint32_t zq = 0;
        if (zf > 0.0f) zq = zf >= 255.0f ? 255u : (uint32_t)zf;
        /* step 3: intensity, clip then truncate then integer map (NaN -> inten_min, as fmax/fmin) */
        float ic = in;
        if (!(ic >= (float)s->inten_min)) ic = (float)s->inten_min;
        if (ic > (float)s->inten_max) ic = (float)s->inten_max;
        const uint32_t iq = (uint32_t)(((int64_t)((int32_t)ic - s->inten_min) * 255) / span);
        /* step 4: integer accumulators */
        acc[ACC_COUNT * cells + cell] += 1u;
        acc[ACC_SUM_I * cells + cell] += iq;
        acc[ACC_SUM_Z * cells + cell] += zq;
        if (iq > acc[ACC_MAX_I * cells + cell]) acc[ACC_MAX_I * cells + cell] = iq;
        if (zq > acc[ACC_MAX_Z * cells + cell]) acc[ACC_MAX_Z * cells + cell] = zq;
        if (zq < acc[ACC_MIN_Z * cells + cell]) acc[ACC_MIN_Z * cells + cell] = zq;
        ++kept;
    }
    return kept;
}

/* step 5: channels.  image: u8 [H][W][nch]; count16: u16 [H][W] or NULL. */
int lmo_finalize(const uint32_t *acc, const lmo_spec *s, const int32_t *channels, int32_t nch,
                 uint8_t *image, uint16_t *count16)
{
    const int64_t cells = (int64_t)s->height * s->width;
    for (int64_t c = 0; c < cells; ++c) {
        const uint64_t cnt = acc[ACC_COUNT * cells + c];
        const uint64_t safe = cnt ? cnt : 1;
        for (int32_t k = 0; k < nch; ++k) {
            uint64_t v;
            switch (channels[k]) {
            case CH_MAX_I:   v = acc[ACC_MAX_I * cells + c]; break;
            case CH_MEAN_I:  v = (acc[ACC_SUM_I * cells + c] + cnt / 2) / safe; break;
            case CH_MIN_Z:   v = cnt ? acc[ACC_MIN_Z * cells + c] : 0; break;
            case CH_MAX_Z:   v = acc[ACC_MAX_Z * cells + c]; break;
            case CH_MEAN_Z:  v = (acc[ACC_SUM_Z * cells + c] + cnt / 2) / safe; break;
            case CH_DENSITY: v = cnt < 255 ? cnt : 255; break;
            default: return -1;
            }
            image[c * nch + k] = (uint8_t)v;
        }
        if (count16) count16[c] = (uint16_t)(cnt < 65535 ? cnt : 65535);
    }
    return 0;
}
