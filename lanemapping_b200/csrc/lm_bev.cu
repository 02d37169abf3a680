// lm_bev.cu -- hand-written sm_100a kernels + the C-ABI of include/lm_bev.h.
//
// Pipeline of the product path (LM_ALGO_BINNED), all integer after the per-point keys:
//
//   bin_points_kernel     one coalesced float4 pass over the packed point records: per point
//                         the cell key and the quantised intensity/height are packed into ONE
//                         32-bit record [cell-in-tile:14 | iq:8 | zq:8]; every CTA counting-sorts
//                         its batch by tile in shared memory and appends each tile's run to a
//                         CTA-private chunk of that tile (chunks come from a bump-allocated pool),
//                         so global writes are contiguous runs and no pre-count pass is needed.
//   scan_tiles/index      tiny: per-tile chunk lists from the chunk side table.
//   reduce_tiles_kernel   persistent CTAs, one tile at a time: stream the tile's chunks, reduce
//                         into a shared-memory accumulator tile with integer atomics
//                         (count/sum add, max, min), derive the u8 channels in place and write the
//                         tile out coalesced (u8 HWC image, u16 count plane, f32 CHW proj, raw acc).
//
// HBM traffic: 16 B/pt read + 4 B/pt written + 4 B/pt read + output  (DESIGN.md section 4).
// No tensor cores: nothing here is a dense contraction.
//
// Bit-exactness: all float steps are single IEEE-754 binary32 operations (__fsub_rn/__fdiv_rn,
// floorf, rintf); compile with -fmad=false and never with -use_fast_math.
#include "lm_bev.h"
#include "lm_las.h"
#include "lm_dev.cuh"
#include "lm_host.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <new>
#include <vector>

namespace {
extern thread_local char g_err[512];
}
// shared with the other translation units of the library (csrc/lm_host.h)
int lm_fail_msg(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int lm_cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
int lm_sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

namespace {

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------
constexpr int TILE_W_LOG2 = 7;            // shared-memory tile: 128 cols x (128 | 64) rows
constexpr int TILE_W = 1 << TILE_W_LOG2;
#ifndef LM_CHUNK_LOG2
#define LM_CHUNK_LOG2 10
#endif
constexpr int CHUNK_RECS = 1 << LM_CHUNK_LOG2;   // records per pool chunk (allocation unit of bin_points)
constexpr int PIECE_RECS = 512;                  // records per work unit of reduce_tiles (2 KB)
constexpr int HALVES = CHUNK_RECS / PIECE_RECS;  // pieces per chunk
static_assert(CHUNK_RECS % PIECE_RECS == 0 && HALVES >= 1, "chunk = whole pieces");
constexpr int IDX_ID_BITS = 23;                  // index entry: piece id | (count-1) << 23
#ifndef LM_BIN_THREADS
#define LM_BIN_THREADS 256
#endif
#ifndef LM_BIN_PPT
#define LM_BIN_PPT 4
#endif
#ifndef LM_BIN_MIN_CTAS
#define LM_BIN_MIN_CTAS 4
#endif
#ifndef LM_BIN_STAGES
#define LM_BIN_STAGES 2      // TMA stage buffers of bin_points (tuning builds: 3 or 4 keep more loads in flight per CTA)
#endif
constexpr int BIN_STAGES = LM_BIN_STAGES;
static_assert(BIN_STAGES >= 2 && BIN_STAGES <= 4, "bin_points stages its batches through 2..4 buffers");
constexpr int BIN_THREADS = LM_BIN_THREADS;
constexpr int BIN_PPT = LM_BIN_PPT;       // points per thread per batch
constexpr int BIN_BATCH = BIN_THREADS * BIN_PPT;
constexpr int MAX_BIN_CTAS = 148 * (LM_BIN_MIN_CTAS + 1);   // most bin CTAs ever launched (B200: 148 SMs)
constexpr int RED_THREADS = 512;
constexpr int RED_MIN_CTAS = 2;          // shared-memory tiles are sized so that two CTAs fit per SM
constexpr int MAX_TILES = 15000;          // bin_points keeps 4 * (1 + NSLOT) B of append state per tile in shared memory
constexpr uint32_t INVALID_U32 = 0xFFFFFFFFu;
constexpr long long SCAN_IN_BIN_MAX_POINTS = 32ll << 20;   // calls up to this size fold scan_tiles into bin_points' last CTA

// accumulator planes held in shared memory by reduce_tiles (bit mask)
enum : int { M_CNT = 1, M_SUMI = 2, M_SUMZ = 4, M_MAXI = 8, M_MINZ = 16, M_MAXZ = 32, M_ALL = 63 };

__host__ __device__ constexpr int popc6(int m) {
    return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1) + ((m >> 4) & 1) + ((m >> 5) & 1);
}
// index of plane `bit` inside the packed plane list of `mask`
__host__ __device__ constexpr int plane_of(int mask, int bit) { return popc6(mask & (bit - 1)); }

struct KParams {
    int H, W, row0, col0;
    float off0, off1, reso0, reso1, zmin, zreso;
    float row_lo, row_hi, col_lo, col_hi;
    float rreso0, rreso1, rzreso;   // RN(1/reso): operands of the exact 3-op division
    int fast_div;                   // all three divisors in the range div_const is proven for
    int imin, imax;
    uint32_t imagic, ishift;        // n/d == umulhi(n << 8, imagic) >> ishift for n < 2^24
    int nch;
    int ch[4];
    int tile_h_log2, tiles_x, tiles_y, T;
    int band;       // > 0: raw accumulators are wanted for output rows [0,band) and [oH-band,oH) only
    int oH, orow;   // height of the output buffers and this window's first row inside them
                    // (a raster with more tiles than one launch handles is done as row windows)
    int bH;         // > 0: batched call, the raster is a stack of samples of bH rows each (proj is [B][C][bH][W])
    int stream_hint;   // bin_points: 1 = the point stream is loaded with an L2 evict-first policy (off: profiles/r01_v11_hint_sweep.txt)
};

struct Ctl {                 // lives right after lm_bev_stats in the workspace; zeroed per call
    unsigned int reserved0;
    unsigned int tile_counter;  // reduce_tiles scheduler
    unsigned int sweep_fail;    // LM_ALGO_SWEEP: != 0 -> the two-pass kernels behind the sweep do the raster
    unsigned int next_batch;    // sweep: batch counter of the producers
    unsigned int bin_done;      // bin_points CTAs that have retired (the last one builds the tile tables)
    unsigned int pad[3];
};

struct Ws {                  // device pointers into the caller's workspace
    lm_bev_stats *stats;
    Ctl *ctl;
    uint32_t *tile_nchunks;  // [T]
    uint32_t *tile_first;    // [T]
    uint32_t *tile_cursor;   // [T]
    uint4 *tile_sched;       // [T] reduce_tiles schedule, heaviest tiles first: {tile, pieces, first index entry, 0}
    uint32_t *cta_chunks;    // [bin CTAs] chunks each bin CTA used of its region
    uint2 *chunk_meta;       // [P] {tile, count}
    uint32_t *chunk_index;   // [P] per-tile chunk lists: id | (count-1) << 23
    uint32_t *pool;          // [P][CHUNK_RECS]
    uint32_t *acc;           // direct path: [6][H][W]
    uint32_t pool_chunks;    // P
    uint32_t region;         // chunks per bin CTA: CTA b owns chunk ids [b * region, (b + 1) * region), local id 0 = none
    uint32_t bin_grid;       // bin CTAs of this launch
    uint32_t scan_in_bin = 0;   // 1: the last bin_points CTA builds the tile tables (no scan_tiles launch: small, launch-bound
                                // calls); 0: scan_tiles_kernel does (large calls: the serial tail behind a 0.35 ms kernel and the
                                // fences of 592 CTAs cost more than the launch, measured +4 us on config 2)
    const unsigned int *gate = nullptr;   // non-NULL: the kernels of the two-pass path return at once unless *gate != 0
                                // (they sit behind a sweep, lm_sweep.cuh, and only run when it gave up)
};
__device__ __forceinline__ bool gated_off(const Ws &ws) { return ws.gate != nullptr && *ws.gate == 0u; }

struct Outs {
    uint8_t *image;
    uint16_t *count16;
    float *proj;
    uint32_t *acc;
    int acc_band;
};

thread_local char g_err[512] = "";
// Tuning knobs of the call in progress (a plan's, include/lm_bev.h lm_bev_tuning; all-zero = defaults for the
// plain entry points).  Nothing in the library reads the process environment.
const lm_bev_tuning k_default_tuning = {};
thread_local const lm_bev_tuning *g_tune = &k_default_tuning;
struct TuneScope {
    const lm_bev_tuning *prev;
    explicit TuneScope(const lm_bev_tuning *t) : prev(g_tune) { g_tune = t ? t : &k_default_tuning; }
    ~TuneScope() { g_tune = prev; }
};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char *what) { return lm_cuda_fail(e, what); }

// ------------------------------------------------------------------------------------------
// per-point quantisation (the spec; oracle/bev_oracle.py::quantise_points restates it)
// ------------------------------------------------------------------------------------------
// a / c for a loop-invariant divisor, bit-identical to __fdiv_rn(a, c).
// Markstein's theorem: with rc = RN(1/c), q0 = RN(a*rc) is a faithful quotient, r = a - q0*c is
// exact in one FMA, and RN(q0 + r*rc) is the correctly rounded a/c -- provided nothing over- or
// underflows, which the exponent guard ensures (callers take the __fdiv_rn path otherwise).
// tests/test_gpu_parity.py::test_exact_division_exhaustive compares all 2^32 dividends.
__device__ __forceinline__ float div_const(float a, float c, float rc) {
    const float q0 = __fmul_rn(a, rc);
    const float r = __fmaf_rn(-q0, c, a);
    return __fmaf_rn(r, rc, q0);
}
// the same three operations on two independent lanes at once (Blackwell FMUL2 / FFMA2: packed
// binary32, each lane rounds exactly like the scalar instruction)
__device__ __forceinline__ float2 div_const2(float2 a, float2 c, float2 rc) {
    const float2 q0 = __fmul2_rn(a, rc);
    const float2 r = __ffma2_rn(make_float2(-q0.x, -q0.y), c, a);
    return __ffma2_rn(r, rc, q0);
}

// the dividends the fast path accepts: finite magnitudes in [2^-40, 2^64)
__device__ __forceinline__ bool div_fast_ok(float a) { return fabsf(a) >= 0x1p-40f && fabsf(a) < 0x1p64f; }

// quotients -> integer keys (shared tail of the fast and the IEEE-division paths).
// The conversions use the saturating F2I modes, which give the same integers as the spec's
// floorf / rintf / clamp-then-truncate on every input (huge values saturate and fall outside the
// window or into the clamp; NaN converts to 0 and is rejected / clamped exactly as fmaxf would).
// the part of the geometry that differs between the samples of a batched call (lm_bev_rasterize_batch)
struct Geo {
    float off0, off1, zmin;
    int row0, col0, H;
};
__device__ __forceinline__ Geo geo_of(const KParams &k) { return Geo{k.off0, k.off1, k.zmin, k.row0, k.col0, k.H}; }

__device__ __forceinline__ bool keys_from_quotients(float qx, float qy, float qz, float inten, const KParams &k,
                                                    const Geo &g, int &lrow, int &lcol, uint32_t &iq, uint32_t &zq) {
    const uint32_t ur = (uint32_t)__float2int_rd(qx) - (uint32_t)g.row0;      // floor, then window shift
    const uint32_t uc = (uint32_t)__float2int_rd(qy) - (uint32_t)g.col0;
    const float s = __fadd_rn(qx, qy);                                        // NaN iff a quotient is NaN
    const bool valid = ur < (uint32_t)g.H && uc < (uint32_t)k.W && s == s;    // never clamped, NaN dropped
    lrow = (int)ur;
    lcol = (int)uc;
    // inverse of coor_img2pc.py:150: round-half-even, clamp to the u8 range, NaN -> 0
    zq = (uint32_t)min(max(__float2int_rn(qz), 0), 255);
    // clip of reference baseline/datasets/laserlane_proposals.py:626-628 (bounds are integers, so
    // clamping after the truncation is the same), then the u8 mapping
    const int iv = min(max(__float2int_rz(inten), k.imin), k.imax);
    const uint32_t n = (uint32_t)(iv - k.imin) * 65280u;          // ((I-imin)*255) << 8, < 2^32
    iq = __umulhi(n, k.imagic) >> k.ishift;                        // == (I-imin)*255 / (imax-imin)
    return valid;
}

// the spec, literally: IEEE division (inverse of reference baseline/utils/coor_img2pc.py:136-139,150)
__device__ __forceinline__ bool quantise_ieee(const float4 p, const KParams &k, const Geo &g, int &lrow, int &lcol,
                                              uint32_t &iq, uint32_t &zq) {
    return keys_from_quotients(__fdiv_rn(__fsub_rn(p.x, g.off0), k.reso0), __fdiv_rn(__fsub_rn(p.y, g.off1), k.reso1),
                               __fdiv_rn(__fsub_rn(p.z, g.zmin), k.zreso), p.w, k, g, lrow, lcol, iq, zq);
}
__device__ __forceinline__ bool quantise_ieee(const float4 p, const KParams &k, int &lrow, int &lcol,
                                              uint32_t &iq, uint32_t &zq) {
    return quantise_ieee(p, k, geo_of(k), lrow, lcol, iq, zq);
}

// same result through div_const; the caller must check that lo/hi (running min/max of the
// dividends' magnitudes) stay inside [2^-40, 2^64) and redo the point with quantise_ieee otherwise
__device__ __forceinline__ bool quantise_fast(const float4 p, const KParams &k, int &lrow, int &lcol,
                                              uint32_t &iq, uint32_t &zq, float &lo, float &hi) {
    const float dx = __fsub_rn(p.x, k.off0), dy = __fsub_rn(p.y, k.off1), dz = __fsub_rn(p.z, k.zmin);
    lo = fminf(lo, fminf(fabsf(dx), fminf(fabsf(dy), fabsf(dz))));
    hi = fmaxf(hi, fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz))));
    return keys_from_quotients(div_const(dx, k.reso0, k.rreso0), div_const(dy, k.reso1, k.rreso1),
                               div_const(dz, k.zreso, k.rzreso), p.w, k, geo_of(k), lrow, lcol, iq, zq);
}
// NaN dividends propagate identically through both paths (fminf/fmaxf skip them), so only the
// magnitudes of the finite ones have to be in range; an all-NaN thread fails the test and takes
// the IEEE path.
__device__ __forceinline__ bool fast_range_ok(float lo, float hi) { return lo >= 0x1p-40f && hi < 0x1p64f; }

__device__ __forceinline__ bool quantise(const float4 p, const KParams &k, int &lrow, int &lcol,
                                         uint32_t &iq, uint32_t &zq) {
    if (k.fast_div) {
        float lo = 0x1p100f, hi = 0.0f;
        const bool v = quantise_fast(p, k, lrow, lcol, iq, zq, lo, hi);
        if (fast_range_ok(lo, hi)) return v;
    }
    return quantise_ieee(p, k, lrow, lcol, iq, zq);
}

__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }

// ------------------------------------------------------------------------------------------
// channel derivation shared by every finishing path
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t channel_value(int ch, uint32_t cnt, uint32_t sum_i, uint32_t sum_z,
                                                  uint32_t max_i, uint32_t min_z, uint32_t max_z) {
    switch (ch) {
        case LM_CH_MAX_I: return max_i;
        case LM_CH_MEAN_I: return cnt ? (sum_i + (cnt >> 1)) / cnt : 0u;     // < 2^32 while cnt < 2^24
        case LM_CH_MIN_Z: return cnt ? min_z : 0u;
        case LM_CH_MAX_Z: return max_z;
        case LM_CH_MEAN_Z: return cnt ? (sum_z + (cnt >> 1)) / cnt : 0u;
        case LM_CH_DENSITY: return cnt < 255u ? cnt : 255u;
    }
    return 0u;
}

// ------------------------------------------------------------------------------------------
// LM_ALGO_DIRECT: global atomics (cross-check path; also finishes merged halo bands)
// ------------------------------------------------------------------------------------------
__global__ void acc_init_kernel(uint32_t *acc, size_t cells) {
    const size_t n = cells * LM_ACC_PLANES;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc[i] = (i / cells == LM_ACC_MIN_Z) ? INVALID_U32 : 0u;
}

__global__ void __launch_bounds__(256) direct_accumulate_kernel(KParams kp, const float4 *__restrict__ pts,
                                                                long long n, uint32_t *acc, lm_bev_stats *stats) {
    const size_t cells = (size_t)kp.H * kp.W;
    unsigned long long nvalid = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int r, c;
        uint32_t iq, zq;
        if (!quantise(ld_stream(pts + i), kp, r, c, iq, zq)) continue;
        const size_t cell = (size_t)r * kp.W + c;
        const uint32_t old = atomicAdd(&acc[LM_ACC_COUNT * cells + cell], 1u);
        if (old >= (1u << 24)) atomicOr(&stats->error, (uint32_t)LM_DEV_ERR_CELL_OVERFLOW);
        atomicAdd(&acc[LM_ACC_SUM_I * cells + cell], iq);
        atomicAdd(&acc[LM_ACC_SUM_Z * cells + cell], zq);
        atomicMax(&acc[LM_ACC_MAX_I * cells + cell], iq);
        atomicMin(&acc[LM_ACC_MIN_Z * cells + cell], zq);
        atomicMax(&acc[LM_ACC_MAX_Z * cells + cell], zq);
        ++nvalid;
    }
    for (int o = 16; o; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    if ((threadIdx.x & 31) == 0 && nvalid) atomicAdd((unsigned long long *)&stats->n_valid, nvalid);
}

// acc [6][H][W] -> outputs, rows [r0,r1)
__global__ void finalize_kernel(KParams kp, const uint32_t *__restrict__ acc, int r0, int r1, Outs out) {
    const size_t cells = (size_t)kp.H * kp.W;
    const size_t lo = (size_t)r0 * kp.W, hi = (size_t)r1 * kp.W;
    for (size_t cell = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; cell < hi;
         cell += (size_t)gridDim.x * blockDim.x) {
        const uint32_t cnt = acc[LM_ACC_COUNT * cells + cell];
        const uint32_t si = acc[LM_ACC_SUM_I * cells + cell], sz = acc[LM_ACC_SUM_Z * cells + cell];
        const uint32_t mi = acc[LM_ACC_MAX_I * cells + cell], nz = acc[LM_ACC_MIN_Z * cells + cell];
        const uint32_t xz = acc[LM_ACC_MAX_Z * cells + cell];
        for (int c = 0; c < kp.nch; ++c) {
            const uint32_t v = channel_value(kp.ch[c], cnt, si, sz, mi, nz, xz);
            if (out.image) out.image[cell * kp.nch + c] = (uint8_t)v;
            if (out.proj) out.proj[(size_t)c * cells + cell] = __fdiv_rn((float)v, 255.0f);
        }
        if (out.count16) out.count16[cell] = (uint16_t)(cnt < 65535u ? cnt : 65535u);
    }
}

__global__ void acc_merge_kernel(uint32_t *dst, long long dstride, const uint32_t *__restrict__ src,
                                 long long sstride, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        dst[LM_ACC_COUNT * dstride + i] += src[LM_ACC_COUNT * sstride + i];
        dst[LM_ACC_SUM_I * dstride + i] += src[LM_ACC_SUM_I * sstride + i];
        dst[LM_ACC_SUM_Z * dstride + i] += src[LM_ACC_SUM_Z * sstride + i];
        dst[LM_ACC_MAX_I * dstride + i] = max(dst[LM_ACC_MAX_I * dstride + i], src[LM_ACC_MAX_I * sstride + i]);
        dst[LM_ACC_MIN_Z * dstride + i] = min(dst[LM_ACC_MIN_Z * dstride + i], src[LM_ACC_MIN_Z * sstride + i]);
        dst[LM_ACC_MAX_Z * dstride + i] = max(dst[LM_ACC_MAX_Z * dstride + i], src[LM_ACC_MAX_Z * sstride + i]);
    }
}

// halo band of a strip: merge the neighbour's raw planes (only those in plane_mask were sent: recv is
// [popc(mask)][rows][W], ascending plane order) into rows [r0,r1) of the local accumulators and finish the band in
// the same pass -- one launch instead of merge + finalize
__global__ void merge_finalize_kernel(KParams kp, uint32_t *acc, int r0, int r1, const uint32_t *__restrict__ recv, int plane_mask, Outs out) {
    const size_t cells = (size_t)kp.H * kp.W;
    const size_t n = (size_t)(r1 - r0) * kp.W, base = (size_t)r0 * kp.W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t cell = base + i;
        uint32_t v[LM_ACC_PLANES];
        int k = 0;
#pragma unroll
        for (int pl = 0; pl < LM_ACC_PLANES; ++pl) {
            v[pl] = acc[pl * cells + cell];
            if (plane_mask >> pl & 1) {
                const uint32_t o = recv[(size_t)k * n + i];
                ++k;
                v[pl] = pl <= LM_ACC_SUM_Z ? v[pl] + o : (pl == LM_ACC_MIN_Z ? min(v[pl], o) : max(v[pl], o));
                acc[pl * cells + cell] = v[pl];
            }
        }
        for (int c = 0; c < kp.nch; ++c) {
            const uint32_t ch = channel_value(kp.ch[c], v[LM_ACC_COUNT], v[LM_ACC_SUM_I], v[LM_ACC_SUM_Z], v[LM_ACC_MAX_I],
                                              v[LM_ACC_MIN_Z], v[LM_ACC_MAX_Z]);
            if (out.image) out.image[cell * kp.nch + c] = (uint8_t)ch;
            if (out.proj) out.proj[(size_t)c * cells + cell] = __fdiv_rn((float)ch, 255.0f);
        }
        if (out.count16) out.count16[cell] = (uint16_t)(v[LM_ACC_COUNT] < 65535u ? v[LM_ACC_COUNT] : 65535u);
    }
}

// One thread moves 16 consecutive bytes of a crop row (a crop row = tile * C bytes of one mosaic row, or zeros
// beyond a ragged edge).  VEC: source and destination are 16-byte aligned for every piece (tile * C, W * C
// multiples of 16 and aligned bases), so both sides are single 128-bit accesses; else bytes.
template <bool VEC>
__global__ void crop_tiles_kernel(const uint8_t *__restrict__ img, int H, int W, int C, int tile, int ncx,
                                  uint8_t *__restrict__ crops, size_t total16) {
    const size_t row_bytes = (size_t)tile * C;                 // bytes of one crop row
    const size_t pieces = (row_bytes + 15) / 16;               // 16-byte pieces per crop row (the last may be short)
    const size_t mosaic_row = (size_t)W * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total16; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pc = i % pieces;
        const size_t rr = i / pieces;                           // crop * tile + r
        const int r = (int)(rr % tile);
        const int crop = (int)(rr / tile);
        const int gy = (crop / ncx) * tile + r;
        const size_t b0 = pc * 16;                              // first byte of the piece inside the crop row
        const size_t gxb = (size_t)(crop % ncx) * row_bytes + b0;
        uint8_t *dst = crops + rr * row_bytes + b0;
        const size_t len = row_bytes - b0 < 16 ? row_bytes - b0 : 16;
        const bool inside = gy < H && gxb + len <= mosaic_row;
        if (VEC) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (inside) v = __ldcs(reinterpret_cast<const uint4 *>(img + (size_t)gy * mosaic_row + gxb));
            __stcs(reinterpret_cast<uint4 *>(dst), v);
        } else {
            for (size_t k = 0; k < len; ++k)
                dst[k] = (gy < H && gxb + k < mosaic_row) ? img[(size_t)gy * mosaic_row + gxb + k] : (uint8_t)0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// LM_ALGO_BINNED stage 1: bin_points
// ------------------------------------------------------------------------------------------
// tile_nchunks counts 512-record PIECES (the reduce kernel's work unit), a chunk holds HALVES of them
__device__ __forceinline__ void publish_chunk(const Ws &ws, uint32_t id, uint32_t count, uint32_t tile) {
    ws.chunk_meta[id] = make_uint2(tile, count);
    atomicAdd(&ws.tile_nchunks[tile], (count + PIECE_RECS - 1) / PIECE_RECS);
}

// Per-CTA append state, all in shared memory:
//   pos[t]      how many records this CTA has appended to tile t so far (monotonic over the kernel)
//   slot[t][2]  ring of 16-bit LOCAL chunk ids: record number q of tile t lives in chunk
//               region_base + slot[t][(q / CHUNK) & 1]
// A point's atomicAdd on pos[t] IS its reservation: it yields the chunk-block and the offset inside
// it.  Every bin CTA owns a private region of the chunk pool (ids [b * region, (b + 1) * region)), so
// the thread that draws the first record of a block allocates that block's chunk with one
// shared-memory atomic -- no global allocator, no abandoned ids -- and 8 bytes of state per tile keep
// four CTAs per SM up to ~2900 tiles.  Nothing in a batch is serial: two balanced phases
// (reserve / store) separated by one barrier each.
// A batch appends at most BIN_BATCH <= 2 * CHUNK records to a tile, i.e. it starts at most two new
// blocks, so four ring slots can never wrap inside the window that is still being read.
constexpr int CHUNK_LOG2 = LM_CHUNK_LOG2;
static_assert(BIN_PPT % 2 == 0, "z quotients are computed two points at a time");
static_assert(BIN_BATCH <= 2 * CHUNK_RECS, "a batch may start at most two chunk blocks per tile");
// a batch that fits one chunk starts at most ONE new block per tile: two ring slots are enough
constexpr int NSLOT = BIN_BATCH <= CHUNK_RECS ? 2 : 4;
constexpr uint32_t MAX_REGION = 65535;  // local chunk ids are 16-bit

// A batched call (lm_bev_rasterize_batch) rasterises up to MAX_BATCH equally-shaped samples as ONE
// stacked raster of n_samples * bH rows: sample s owns the 1024-point batches [first[s], first[s+1])
// and brings its own origin / window shift; everything after the keys is shared.
constexpr int MAX_BATCH = 32;
struct BatchTab {
    const float4 *pts[MAX_BATCH];
    uint32_t first[MAX_BATCH + 1];
    uint32_t count[MAX_BATCH];
    float off0[MAX_BATCH], off1[MAX_BATCH], zmin[MAX_BATCH];
    int row0[MAX_BATCH], col0[MAX_BATCH];
    int nb, bH;
};

// ------------------------------------------------------------------------------------------
// per-tile tables: first index of every tile's piece list + heaviest-first tile order.  Runs in the LAST
// bin_points CTA to retire (no launch of its own), or as scan_tiles_kernel when no bin kernel ran.
// scratch: NT + 1024 + 8 words of shared memory.
// ------------------------------------------------------------------------------------------
// inclusive block scan of one value per thread: warp shuffles + one pass over the warp totals (3 barriers)
template <int NT>
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t *s_warp /* [NT / 32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    __syncthreads();                       // s_warp may still be read from a previous scan
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) base += w < warp ? s_warp[w] : 0u;
    return v + base;
}

template <int NT>
__device__ __forceinline__ void scan_tiles_body(const Ws &ws, const KParams &kp, uint32_t *scratch) {
    const int T = kp.T;
    uint32_t *s_lvl = scratch;              // [1024] tiles per weight level -> end of every level's slot range
    uint32_t *s_warp = scratch + 1024;      // [NT / 32]
    uint32_t *s_flag = s_warp + NT / 32;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_flag[0] = ws.stats->error & LM_DEV_ERR_POOL;
        ws.stats->n_tiles = (uint32_t)T;
    }
    for (int i = tid; i < 1024; i += NT) s_lvl[i] = 0;
    __syncthreads();
    // Thread `tid` owns tiles tid, tid + NT, ... (coalesced table accesses; a raster of 24 300 tiles spent 48 us here
    // with one contiguous range per thread).  A tile's pieces only have to be contiguous in the index, in whatever
    // order the tiles follow each other: thread by thread, each thread's tiles in ascending order.
    if (s_flag[0]) {   // pool exhausted: publish an empty raster instead of reading half-built lists
        for (int t = tid; t < T; t += NT) ws.tile_nchunks[t] = 0;
    }
    uint32_t sum = 0;
#pragma unroll 4
    for (int t = tid; t < T; t += NT) {
        const uint32_t c = __ldcg(&ws.tile_nchunks[t]);      // other CTAs' atomics: read at the L2
        sum += c;
        atomicAdd(&s_lvl[1023u - min(c, 1023u)], 1u);      // level 0 = heaviest
    }
    uint32_t run = block_inclusive_scan<NT>(sum, s_warp) - sum;      // first piece index of this thread's tiles
    // inclusive scan of the 1024 level counts: 1024 / NT consecutive levels per thread, then across threads
    constexpr int LPT = 1024 / NT;
    uint32_t lsum = 0;
#pragma unroll
    for (int q = 0; q < LPT; ++q) { lsum += s_lvl[tid * LPT + q]; s_lvl[tid * LPT + q] = lsum; }
    const uint32_t lbase = block_inclusive_scan<NT>(lsum, s_warp) - lsum;
#pragma unroll
    for (int q = 0; q < LPT; ++q) s_lvl[tid * LPT + q] += lbase;
    __syncthreads();
    // counting sort by weight level: tile_sched lists the heaviest tiles first so that the persistent reduce CTAs
    // finish together (longest-processing-time-first); an entry is everything reduce_tiles needs to start on the tile
#pragma unroll 4
    for (int t = tid; t < T; t += NT) {
        const uint32_t c = __ldcg(&ws.tile_nchunks[t]);
        ws.tile_first[t] = run;
        const uint32_t lvl = 1023u - min(c, 1023u);
        const uint32_t pos = atomicAdd(&s_lvl[lvl], 0xFFFFFFFFu) - 1u;      // fill each level's slot range from its end
        ws.tile_sched[pos] = make_uint4((uint32_t)t, c, run, 0u);
        run += c;
    }
}

// LAS: the input is the point-data block of an uncompressed LAS file (record_length bytes per point)
// and the decode of lm_dev.cuh::las_decode_record runs on the staged bytes; no float4 copy of the
// cloud ever exists in HBM.
// CT (compact table): the per-tile append state is not indexed by the tile id (8 B x T of shared memory: one CTA
// per SM at config 4's 12 150 tiles) but lives in an open-addressing hash table of CT_SLOTS entries keyed by the tile id --
// a CTA's contiguous range of a scan-ordered cloud touches a few hundred tiles however many the raster has.  4 CTAs per SM
// for any T.  A CTA that meets more tiles than the table holds raises stats->ct_overflow and drops the record: the caller
// has the direct-indexed kernels queued behind, gated on that word, and they redo the raster.
constexpr int CT_SLOTS = 1024;
constexpr uint32_t CT_EMPTY = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t atoms_cas_u32(uint32_t a, uint32_t cmp, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(v) : "memory");
    return old;
}
// home slot of a tile id
__device__ __forceinline__ uint32_t ct_hash(uint32_t tile) {
    static_assert(CT_SLOTS == 1 << 10, "hash width");
    return (tile * 2654435761u) >> (32 - 10);
}
// slot of `tile` in the table at sm_key, probing from slot s whose key was just read as k (claims an empty slot on
// first sight); CT_EMPTY if the table is full
__device__ __noinline__ uint32_t ct_find_slow(uint32_t sm_key, uint32_t tile, uint32_t s, uint32_t k) {
    for (int probe = 0; probe < CT_SLOTS; ++probe) {
        if (k == CT_EMPTY) k = atoms_cas_u32(sm_key + 4u * s, CT_EMPTY, tile);
        if (k == tile || k == CT_EMPTY) return s;            // found, or claimed just now
        s = (s + 1u) & (CT_SLOTS - 1);
        k = lds_u32(sm_key + 4u * s);
    }
    return CT_EMPTY;
}

template <bool BATCHED, bool LAS, bool CT = false>
__device__ __forceinline__ void bin_points_body(const KParams &kp, const float4 *__restrict__ pts, long long n, const Ws &ws,
                                                const BatchTab *btp, const LasXform *xfp) {
    const BatchTab &bt = *btp;       // only dereferenced when BATCHED
    const LasXform &xf = *xfp;       // only dereferenced when LAS
    const uint32_t rec_bytes = LAS ? (uint32_t)xf.record_length : 16u;
    // one stage buffer: a full batch of records, 16-byte granular, + one spare 16 B (whole words are read)
    const uint32_t stage_bytes = LAS ? ((BIN_BATCH * rec_bytes + 15u) & ~15u) + 16u : BIN_BATCH * 16u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = CT ? CT_SLOTS : kp.T;          // entries of the append-state arrays (CT: table slots)
    // layout: stage[2][BIN_BATCH] float4 (TMA double buffer) | pos[T] u32 | slot[T][NSLOT] u16 | CT: key[CT_SLOTS] u32
    uint32_t sm_stage = smem_u32(smem_raw);
    uint32_t sm_pos = sm_stage + (uint32_t)BIN_STAGES * stage_bytes;
    uint32_t sm_slot = sm_pos + (uint32_t)T * 4u;
    const uint32_t sm_key = sm_slot + (uint32_t)T * 2u * NSLOT;
    // keep the three bases in registers: without this the compiler re-derives them (window base +
    // offsets, ~5 instructions) at every use because they are cheap to rematerialise
    asm volatile("" : "+r"(sm_stage), "+r"(sm_pos), "+r"(sm_slot));
    __shared__ uint32_t s_next;                                              // next local chunk id of this CTA's region
    __shared__ __align__(8) uint64_t s_bar[BIN_STAGES];                      // TMA completion barriers

    const int tid = threadIdx.x;
#if LM_BIN_STAGES == 2
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); s_next = 1u; }
#else
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < BIN_STAGES; ++b) mbar_init(&s_bar[b], 1);
        s_next = 1u;
    }
#endif
    const uint32_t sm_next = smem_u32(&s_next);
    const uint32_t region_base = blockIdx.x * ws.region;
    for (int t = tid; t < T; t += BIN_THREADS) sts_u32(sm_pos + 4u * t, 0u);
    if (CT) for (int t = tid; t < CT_SLOTS; t += BIN_THREADS) sts_u32(sm_key + 4u * t, CT_EMPTY);
    __syncthreads();

    const long long nb = BATCHED ? (long long)bt.first[bt.nb] : (n + BIN_BATCH - 1) / BIN_BATCH;
    const long long b0 = nb * blockIdx.x / gridDim.x, b1 = nb * (blockIdx.x + 1) / gridDim.x;
    const int my_batches = (int)(b1 - b0);
    const uint32_t tail = BATCHED ? 0u : (uint32_t)(n - (nb - 1) * BIN_BATCH);   // points in the very last batch
    // batch k of this CTA: where its records start and how many there are (samp = its sample, batched
    // calls only; batches are visited in order, so the sample cursor only ever moves forward)
    auto batch_src = [&](int k, int &samp, uint32_t &np) -> const void * {
        const long long g = b0 + k;
        if (BATCHED) {
            while (g >= (long long)bt.first[samp + 1]) ++samp;
            const uint32_t lb = (uint32_t)(g - bt.first[samp]);
            const uint32_t rem = bt.count[samp] - lb * (uint32_t)BIN_BATCH;
            np = rem < (uint32_t)BIN_BATCH ? rem : (uint32_t)BIN_BATCH;
            return bt.pts[samp] + (size_t)lb * BIN_BATCH;
        }
        np = g == nb - 1 ? tail : (uint32_t)BIN_BATCH;
        if (LAS) return reinterpret_cast<const unsigned char *>(pts) + (size_t)g * BIN_BATCH * rec_bytes;
        return pts + g * BIN_BATCH;
    };
    // bytes of a batch as the bulk copy wants them (multiples of 16; a LAS tail may read up to 15
    // bytes past its last record -- the caller's buffer is padded, lm_las.h)
    auto batch_bytes = [&](uint32_t np) -> uint32_t { return LAS ? (np * rec_bytes + 15u) & ~15u : np * 16u; };
    int samp = 0, samp_pf = 0;                                               // sample cursors: compute / prefetch (thread 0)
#if LM_BIN_STAGES == 2
    if (tid == 0 && my_batches > 0) {
        uint32_t np;
        const void *src = batch_src(0, samp_pf, np);
        if (kp.stream_hint & 1) bulk_load_stream(smem_raw, src, batch_bytes(np), &s_bar[0]);
        else bulk_load(smem_raw, src, batch_bytes(np), &s_bar[0]);
    }
#else
    // the first BIN_STAGES - 1 batches start flying before the loop
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < BIN_STAGES - 1; ++b) {
            if (b < my_batches) {
                uint32_t np;
                const void *src = batch_src(b, samp_pf, np);
                if (kp.stream_hint & 1) bulk_load_stream(smem_raw + (uint32_t)b * stage_bytes, src, batch_bytes(np), &s_bar[b]);
                else bulk_load(smem_raw + (uint32_t)b * stage_bytes, src, batch_bytes(np), &s_bar[b]);
            }
        }
    }
#endif
    Geo geo = geo_of(kp);
    int row_shift = 0;                                                       // first row of the sample in the stacked raster

    for (int k = 0; k < my_batches; ++k) {
#if LM_BIN_STAGES == 2
        const uint32_t buf = (uint32_t)k & 1u;
        // the NEXT batch starts flying now, into the buffer that was consumed one batch ago
        if (tid == 0 && k + 1 < my_batches) {
            uint32_t np;
            const void *src = batch_src(k + 1, samp_pf, np);
            if (kp.stream_hint & 1) bulk_load_stream(smem_raw + (buf ^ 1u) * stage_bytes, src, batch_bytes(np), &s_bar[buf ^ 1u]);
            else bulk_load(smem_raw + (buf ^ 1u) * stage_bytes, src, batch_bytes(np), &s_bar[buf ^ 1u]);
        }
#else
        const uint32_t buf = (uint32_t)k % (uint32_t)BIN_STAGES;
        // batch k + BIN_STAGES - 1 starts flying now, into the buffer that was consumed one batch ago
        if (tid == 0 && k + BIN_STAGES - 1 < my_batches) {
            const uint32_t nbuf = (uint32_t)(k + BIN_STAGES - 1) % (uint32_t)BIN_STAGES;
            uint32_t np;
            const void *src = batch_src(k + BIN_STAGES - 1, samp_pf, np);
            if (kp.stream_hint & 1) bulk_load_stream(smem_raw + nbuf * stage_bytes, src, batch_bytes(np), &s_bar[nbuf]);
            else bulk_load(smem_raw + nbuf * stage_bytes, src, batch_bytes(np), &s_bar[nbuf]);
        }
#endif
        uint32_t npts;
        batch_src(k, samp, npts);
        if (BATCHED) {
            geo = Geo{bt.off0[samp], bt.off1[samp], bt.zmin[samp], bt.row0[samp], bt.col0[samp], bt.bH};
            row_shift = samp * bt.bH;
        }
#if LM_BIN_STAGES == 2
        mbar_wait(&s_bar[buf], ((uint32_t)k >> 1) & 1u);     // this batch's records have landed
#else
        mbar_wait(&s_bar[buf], ((uint32_t)k / (uint32_t)BIN_STAGES) & 1u);
#endif
        float4 p[BIN_PPT];
        const uint32_t my_stage = sm_stage + buf * stage_bytes + (uint32_t)tid * rec_bytes;
#pragma unroll
        for (int j = 0; j < BIN_PPT; ++j) {
            if (LAS) p[j] = las_decode_record(my_stage + (uint32_t)j * (BIN_THREADS * rec_bytes), xf);
            else p[j] = lds_f4(my_stage + (uint32_t)j * (BIN_THREADS * 16u));
        }
        if (npts < (uint32_t)BIN_BATCH) {                        // only the very last batch of the cloud
#pragma unroll
            for (int j = 0; j < BIN_PPT; ++j)
                if ((uint32_t)(j * BIN_THREADS + tid) >= npts) p[j].x = __int_as_float(0x7fc00000);   // NaN x: dropped
        }

        uint32_t rec[BIN_PPT], tl[BIN_PPT], ps[BIN_PPT];
        {
            int r[BIN_PPT], c[BIN_PPT];
            uint32_t iq[BIN_PPT], zq[BIN_PPT];
            bool ok[BIN_PPT];
            // exact divisions, two lanes per instruction: (x, y) of one point, z of two points
            float2 qxy[BIN_PPT];
            float qz[BIN_PPT];
            float lo = 0x1p100f, hi = 0.0f;
            const float2 noff = make_float2(-geo.off0, -geo.off1), reso2 = make_float2(kp.reso0, kp.reso1),
                         rr2 = make_float2(kp.rreso0, kp.rreso1);
            const float2 nzmin2 = make_float2(-geo.zmin, -geo.zmin), zreso2 = make_float2(kp.zreso, kp.zreso),
                         rz2 = make_float2(kp.rzreso, kp.rzreso);
#pragma unroll
            for (int j = 0; j < BIN_PPT; ++j) {
                const float2 d = __fadd2_rn(make_float2(p[j].x, p[j].y), noff);          // x - off0, y - off1
                lo = fminf(lo, fminf(fabsf(d.x), fabsf(d.y)));
                hi = fmaxf(hi, fmaxf(fabsf(d.x), fabsf(d.y)));
                qxy[j] = div_const2(d, reso2, rr2);
            }
#pragma unroll
            for (int j = 0; j < BIN_PPT; j += 2) {
                const float2 d = __fadd2_rn(make_float2(p[j].z, p[j + 1].z), nzmin2);    // z - local_min_ele
                lo = fminf(lo, fminf(fabsf(d.x), fabsf(d.y)));
                hi = fmaxf(hi, fmaxf(fabsf(d.x), fabsf(d.y)));
                const float2 q = div_const2(d, zreso2, rz2);
                qz[j] = q.x;
                qz[j + 1] = q.y;
            }
            if (kp.fast_div && fast_range_ok(lo, hi)) {
#pragma unroll
                for (int j = 0; j < BIN_PPT; ++j)
                    ok[j] = keys_from_quotients(qxy[j].x, qxy[j].y, qz[j], p[j].w, kp, geo, r[j], c[j], iq[j], zq[j]);
            } else {                                            // rare: 0, denormal-scale, huge, inf or all-NaN dividends
#pragma unroll
                for (int j = 0; j < BIN_PPT; ++j) ok[j] = quantise_ieee(p[j], kp, geo, r[j], c[j], iq[j], zq[j]);
            }
            if (BATCHED) {
#pragma unroll
                for (int j = 0; j < BIN_PPT; ++j) r[j] += row_shift;
            }
#pragma unroll
            for (int j = 0; j < BIN_PPT; ++j) {
                const uint32_t t = (uint32_t)((r[j] >> kp.tile_h_log2) * kp.tiles_x + (c[j] >> TILE_W_LOG2));
                const uint32_t cell = (uint32_t)(((r[j] & ((1 << kp.tile_h_log2) - 1)) << TILE_W_LOG2) | (c[j] & (TILE_W - 1)));
                rec[j] = (cell << 16) | (iq[j] << 8) | zq[j];
                tl[j] = ok[j] ? t : INVALID_U32;
            }
        }
        // ---- phase 1: reserve.  The first record of a chunk block allocates the block's chunk.
        uint32_t sa[BIN_PPT];                                    // shared address of the point's ring slot
        uint32_t en[BIN_PPT];                                    // entry of the append state: the tile id, or its table slot
#pragma unroll
        for (int j = 0; j < BIN_PPT; ++j) en[j] = CT ? ct_hash(tl[j]) : tl[j];
        if (CT) {
            // all home-slot keys first (independent loads); a tile nearly always sits in its home slot
            uint32_t k[BIN_PPT];
#pragma unroll
            for (int j = 0; j < BIN_PPT; ++j) k[j] = lds_u32(sm_key + 4u * en[j]);
#pragma unroll
            for (int j = 0; j < BIN_PPT; ++j) {
                if (tl[j] != INVALID_U32 && k[j] != tl[j]) {
                    en[j] = ct_find_slow(sm_key, tl[j], en[j], k[j]);
                    if (en[j] == CT_EMPTY) {                     // more tiles than the table holds: the gated kernels redo the call
                        atomicExch(&ws.stats->ct_overflow, 1u);
                        tl[j] = INVALID_U32;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < BIN_PPT; ++j) {
            if (tl[j] != INVALID_U32) {
                ps[j] = atoms_add(sm_pos + 4u * en[j], 1u);
                sa[j] = sm_slot + (2u * NSLOT) * en[j] + ((ps[j] >> (CHUNK_LOG2 - 1)) & (2u * (NSLOT - 1)));   // ring slot of its block
                if ((ps[j] & (CHUNK_RECS - 1)) == 0) {
                    uint32_t id = atoms_add(sm_next, 1u);
                    if (id >= ws.region) {                       // cannot happen with lm_bev_workspace_bytes' size
                        atomicOr(&ws.stats->error, (uint32_t)LM_DEV_ERR_POOL);
                        id = 0;                                  // local chunk 0 is a scratch chunk nobody reads
                    }
                    sts_u16(sa[j], id);
                }
            }
        }
        __syncthreads();
        // ---- phase 2: store.  Lanes that hit the same tile drew consecutive positions, so they write
        //      neighbouring words; L2 merges partial sectors before they reach HBM.
        uint32_t cid[BIN_PPT];
#pragma unroll
        for (int j = 0; j < BIN_PPT; ++j)
            cid[j] = tl[j] != INVALID_U32 ? region_base + lds_u16(sa[j]) : 0u;
#pragma unroll
        for (int j = 0; j < BIN_PPT; ++j) {
            if (tl[j] != INVALID_U32) {
                const uint32_t off = ps[j] & (CHUNK_RECS - 1);
                ws.pool[cid[j] * (uint32_t)CHUNK_RECS + off] = rec[j];        // record index < 2^32 (checked on the host)
                if (off == 0 && ps[j] != 0) {                    // the previous block of this tile is complete
                    const uint32_t prev = lds_u16(sm_slot + 2u * (en[j] * NSLOT + (((ps[j] >> CHUNK_LOG2) - 1u) & (NSLOT - 1))));
                    if (prev) publish_chunk(ws, region_base + prev, CHUNK_RECS, tl[j]);
                }
            }
        }
        __syncthreads();
    }
    // ---- retire this CTA's open chunks; the final positions also give the number of points kept
    uint32_t my_valid = 0;
    for (int t = tid; t < T; t += BIN_THREADS) {
        const uint32_t q = lds_u32(sm_pos + 4u * t);
        if (q) {
            my_valid += q;
            const uint32_t last = q - 1;
            const uint32_t id = lds_u16(sm_slot + 2u * (t * NSLOT + ((last >> CHUNK_LOG2) & (NSLOT - 1))));
            if (id) publish_chunk(ws, region_base + id, (last & (CHUNK_RECS - 1)) + 1u, CT ? lds_u32(sm_key + 4u * t) : (uint32_t)t);
        }
    }
    for (int o = 16; o; o >>= 1) my_valid += __shfl_xor_sync(0xffffffffu, my_valid, o);
    if ((tid & 31) == 0 && my_valid) atomicAdd((unsigned long long *)&ws.stats->n_valid, (unsigned long long)my_valid);
    if (tid == 0) {                                              // nobody allocates after the last batch's first barrier
        const uint32_t used = min(s_next, ws.region) - 1u;
        ws.cta_chunks[blockIdx.x] = used;
        if (used) atomicAdd(&ws.stats->n_chunks, used);
    }
    // ---- the last CTA to retire has every tile's piece count in front of it: it builds the tile tables
    //      (what used to be the scan_tiles launch), in the stage buffers nobody needs any more
    if (!ws.scan_in_bin) return;
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&ws.ctl->bin_done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        static_assert(BIN_STAGES * BIN_BATCH * 16 >= (1024 + BIN_THREADS / 32 + 8) * 4, "scan scratch fits the stage buffers");
        scan_tiles_body<BIN_THREADS>(ws, kp, reinterpret_cast<uint32_t *>(smem_raw));
    }
}


__global__ void __launch_bounds__(BIN_THREADS, LM_BIN_MIN_CTAS) bin_points_kernel(KParams kp, const float4 *__restrict__ pts,
                                                                               long long n, Ws ws) {
    if (gated_off(ws)) return;
    bin_points_body<false, false>(kp, pts, n, ws, nullptr, nullptr);
}
__global__ void __launch_bounds__(BIN_THREADS, LM_BIN_MIN_CTAS) bin_points_ct_kernel(KParams kp, const float4 *__restrict__ pts,
                                                                                  long long n, Ws ws) {
    bin_points_body<false, false, true>(kp, pts, n, ws, nullptr, nullptr);
}
__global__ void ct_epilogue_kernel(lm_bev_stats *stats) {      // behind a compact-table pass that overflowed: the gated kernels count again
    if (stats->ct_overflow) { stats->n_valid = 0; stats->n_chunks = 0; }
}
__global__ void __launch_bounds__(BIN_THREADS, LM_BIN_MIN_CTAS) bin_points_batch_kernel(KParams kp, const __grid_constant__ BatchTab bt,
                                                                                     Ws ws) {
    bin_points_body<true, false>(kp, nullptr, 0, ws, &bt, nullptr);
}
__global__ void __launch_bounds__(BIN_THREADS, 3) bin_points_las_kernel(KParams kp, const __grid_constant__ LasXform xf,
                                                                       const unsigned char *__restrict__ recs, long long n, Ws ws) {
    bin_points_body<false, true>(kp, reinterpret_cast<const float4 *>(recs), n, ws, nullptr, &xf);
}

// ------------------------------------------------------------------------------------------
// LAS point records -> packed float4 (lm_las_decode): the same TMA-staged loop without the binning
// ------------------------------------------------------------------------------------------
constexpr int LAS_THREADS = 256;
constexpr int LAS_TILE = 1024;           // records per stage buffer
__host__ __device__ inline uint32_t las_stage_bytes(uint32_t rec_bytes, uint32_t records) {
    return ((records * rec_bytes + 15u) & ~15u) + 16u;
}

__global__ void __launch_bounds__(LAS_THREADS) las_decode_kernel(const __grid_constant__ LasXform xf,
                                                                 const unsigned char *__restrict__ recs, long long n,
                                                                 float4 *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int tid = threadIdx.x;
    const uint32_t rec_bytes = (uint32_t)xf.record_length;
    const uint32_t stage_bytes = las_stage_bytes(rec_bytes, LAS_TILE);
    const uint32_t sm_stage = smem_u32(smem_raw);
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    __syncthreads();
    const long long nb = (n + LAS_TILE - 1) / LAS_TILE;
    const long long b0 = nb * blockIdx.x / gridDim.x, b1 = nb * (blockIdx.x + 1) / gridDim.x;
    const int my_batches = (int)(b1 - b0);
    auto batch_points = [&](int k) -> uint32_t {
        const long long left = n - (b0 + k) * LAS_TILE;
        return left < LAS_TILE ? (uint32_t)left : (uint32_t)LAS_TILE;
    };
    auto load = [&](int k, uint32_t buf) {
        bulk_load(smem_raw + buf * stage_bytes, recs + (size_t)(b0 + k) * LAS_TILE * rec_bytes,
                  (batch_points(k) * rec_bytes + 15u) & ~15u, &s_bar[buf]);
    };
    if (tid == 0 && my_batches > 0) load(0, 0);
    for (int k = 0; k < my_batches; ++k) {
        const uint32_t buf = (uint32_t)k & 1u;
        if (tid == 0 && k + 1 < my_batches) load(k + 1, buf ^ 1u);
        const uint32_t npts = batch_points(k);
        mbar_wait(&s_bar[buf], ((uint32_t)k >> 1) & 1u);
        float4 *dst = out + (b0 + k) * LAS_TILE;
#pragma unroll
        for (int j = 0; j < LAS_TILE / LAS_THREADS; ++j) {
            const uint32_t i = (uint32_t)(j * LAS_THREADS + tid);
            if (i < npts) __stcs(dst + i, las_decode_record(sm_stage + buf * stage_bytes + i * rec_bytes, xf));
        }
        __syncthreads();            // everybody is done with this buffer before it is refilled
    }
}

// ------------------------------------------------------------------------------------------
// stage 2: per-tile chunk lists
// ------------------------------------------------------------------------------------------
// does tile t touch the halo bands (output rows [0,band) or [oH-band,oH))?
__device__ __forceinline__ bool tile_in_band(const KParams &kp, int t) {
    if (kp.band <= 0) return false;
    const int r0 = kp.orow + ((t / kp.tiles_x) << kp.tile_h_log2);
    const int r1 = min(r0 + (1 << kp.tile_h_log2), kp.orow + kp.H);
    return r0 < kp.band || r1 > kp.oH - kp.band;
}

__global__ void __launch_bounds__(1024) scan_tiles_kernel(Ws ws, KParams kp) {
    if (gated_off(ws)) return;
    __shared__ uint32_t s_scratch[1024 + 32 + 8];
    scan_tiles_body<1024>(ws, kp, s_scratch);
}

__global__ void index_chunks_kernel(Ws ws) {
    if (gated_off(ws)) return;
    if (ws.stats->error & LM_DEV_ERR_POOL) return;
    // chunk ids are (bin CTA, local id) pairs: CTA b used local ids 1 .. cta_chunks[b] of its region.  One block walks
    // one bin CTA's used ids (not the whole id space: a region is mostly unused reservation)
    for (uint32_t b = blockIdx.x; b < ws.bin_grid; b += gridDim.x) {
        const uint32_t used = ws.cta_chunks[b];
        for (uint32_t k = 1u + threadIdx.x; k <= used; k += blockDim.x) {
            const uint32_t id = b * ws.region + k;
            const uint2 m = ws.chunk_meta[id];           // every allocated chunk was published with count >= 1
            const uint32_t np = (m.y + PIECE_RECS - 1) / PIECE_RECS;               // non-empty pieces of this chunk
            const uint32_t slot = ws.tile_first[m.x] + atomicAdd(&ws.tile_cursor[m.x], np);
            for (uint32_t q = 0; q < np; ++q) {
                const uint32_t c = min(m.y - q * PIECE_RECS, (uint32_t)PIECE_RECS);
                ws.chunk_index[slot + q] = (id * HALVES + q) | ((c - 1u) << IDX_ID_BITS);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 3: reduce_tiles
// ------------------------------------------------------------------------------------------
// (sum + count/2) / count for count in [1, 4095], sum <= 255 * count, without an integer division:
// the float estimate is off by < 4.6e-5 absolute (quotient <= 255.5, relative error < 1.8e-7) while
// a non-integer true quotient is >= 1/4095 = 2.4e-4 away from the integers, so adding 2^-13
// before truncating lands on the exact floor.  lm_bev_selftest_mean checks every (count, sum).
__device__ __forceinline__ uint32_t mean_small(uint32_t sum, uint32_t cnt) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)cnt));
    return (uint32_t)__fmaf_rn((float)(sum + (cnt >> 1)), r, 0x1p-13f);
}

// four finished pixels (C bytes each, little end first) -> the C 32-bit words they occupy in the image
template <int C>
__device__ __forceinline__ void store_pixels4(uint8_t *dst, const uint32_t pk[4]) {
    uint32_t *w = reinterpret_cast<uint32_t *>(dst);
    if (C == 1) {
        w[0] = (pk[0] & 0xFFu) | ((pk[1] & 0xFFu) << 8) | ((pk[2] & 0xFFu) << 16) | (pk[3] << 24);
    } else if (C == 2) {
        w[0] = (pk[0] & 0xFFFFu) | (pk[1] << 16);
        w[1] = (pk[2] & 0xFFFFu) | (pk[3] << 16);
    } else if (C == 3) {
        w[0] = (pk[0] & 0xFFFFFFu) | (pk[1] << 24);
        w[1] = ((pk[1] >> 8) & 0xFFFFu) | (pk[2] << 16);
        w[2] = ((pk[2] >> 16) & 0xFFu) | (pk[3] << 8);
    } else if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    } else {
        w[0] = pk[0]; w[1] = pk[1]; w[2] = pk[2]; w[3] = pk[3];
    }
}

// count and one sum share a word [count:12 | sum:20] when PACKED: one atomic instead of two (the
// kernel is bound by shared-memory atomic wavefronts).  sum <= 255 * count < 2^20 while count < 4096.
// Nothing is returned by the atomics (no scoreboard wait on the hot path); a wrapped count field is
// found afterwards by conservation -- the counts of a tile's cells must add up to the records
// streamed into it, and a wrap loses 4096 -- and the tile is then redone unpacked: exact for any input.
constexpr uint32_t PK_SHIFT = 20, PK_SUM_MASK = (1u << PK_SHIFT) - 1u, PK_CNT_MAX = (1u << (32 - PK_SHIFT)) - 1u;

// max into a shared-memory word.  -DLM_RED_MAXCHECK=1 skips the atomic when the word already holds at least v (a
// cell's maximum is raised by only ~ln(n) of its n points; a stale load is harmless, a maximum only grows) -- one
// load + one predicated instruction, no branch.  MEASURED SLOWER (reduce_tiles 0.156 -> 0.176 ms on config 2,
// 1.19 -> 1.38 ms on config 4): the load costs the shared-memory pipe as much as the atomic it saves.  Off.
#ifndef LM_RED_MAXCHECK
#define LM_RED_MAXCHECK 0
#endif
__device__ __forceinline__ void smem_max(uint32_t *p, uint32_t v) {
#if LM_RED_MAXCHECK
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    uint32_t cur;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(a) : "memory");
    asm volatile("{\n.reg .pred q;\nsetp.gt.u32 q, %1, %2;\n@q red.shared.max.u32 [%0], %1;\n}" ::"r"(a), "r"(v), "r"(cur) : "memory");
#else
    atomicMax(p, v);
#endif
}

template <int MASK, bool PACKED>
__device__ __forceinline__ void accumulate_rec(uint32_t rec, uint32_t *a_cnt, uint32_t *a_sumi, uint32_t *a_sumz,
                                               uint32_t *a_maxi, uint32_t *a_minz, uint32_t *a_maxz) {
    const uint32_t cell = rec >> 16, iq = (rec >> 8) & 0xFFu, zq = rec & 0xFFu;
    constexpr bool PACK_Z = PACKED && (MASK & M_SUMZ);            // partner of the count: sum_z if tracked, else sum_i
    constexpr bool PACK_I = PACKED && !(MASK & M_SUMZ) && (MASK & M_SUMI);
    if (PACK_Z || PACK_I) {
        atomicAdd(&a_cnt[cell], (1u << PK_SHIFT) | (PACK_Z ? zq : iq));
    } else if (MASK & M_CNT) {
        atomicAdd(&a_cnt[cell], 1u);
    }
    if ((MASK & M_SUMI) && !PACK_I) atomicAdd(&a_sumi[cell], iq);
    if ((MASK & M_SUMZ) && !PACK_Z) atomicAdd(&a_sumz[cell], zq);
    if (MASK & M_MAXI) smem_max(&a_maxi[cell], iq);
    if (MASK & M_MINZ) smem_max(&a_minz[cell], 256u - zq);
    if (MASK & M_MAXZ) smem_max(&a_maxz[cell], zq);
}

// stream one tile's chunks: one chunk per warp and iteration, coalesced uint4 loads with the next
// chunk's index entry and records fetched while the current ones are reduced
template <int MASK, bool PACKED>
__device__ __forceinline__ uint32_t stream_tile(const Ws &ws, const uint32_t *my_index, uint32_t nchunks, int warp, int lane,
                                                uint32_t *a_cnt, uint32_t *a_sumi, uint32_t *a_sumz, uint32_t *a_maxi,
                                                uint32_t *a_minz, uint32_t *a_maxz) {
    uint32_t streamed = 0;                                       // records this warp reduced (warp-uniform)
    constexpr int V = PIECE_RECS / 128;                          // uint4 loads per lane per piece
    constexpr int NWARPS = RED_THREADS / 32;
    constexpr uint32_t ID_MASK = (1u << IDX_ID_BITS) - 1u;
    uint32_t c = warp;
    uint32_t ent = c < nchunks ? __ldg(my_index + c) : 0u;
    uint4 v[V];
    if (c < nchunks) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ws.pool + (size_t)(ent & ID_MASK) * PIECE_RECS);
        const uint32_t cnt = (ent >> IDX_ID_BITS) + 1u;
#pragma unroll
        for (int q = 0; q < V; ++q)
            if ((uint32_t)((q * 32 + lane) * 4) < cnt) v[q] = __ldcs(src + q * 32 + lane);
    }
    // pieces further ahead are pulled into L2 with a bulk prefetch (no registers, one lane): the
    // register prefetch above only covers one piece of work, less than an HBM round trip under load
    auto l2_prefetch = [&](uint32_t ci) {
        if (lane == 0 && ci < nchunks) {
            const uint32_t e = __ldg(my_index + ci);
            const uint32_t bytes = ((((e >> IDX_ID_BITS) + 1u) * 4u) + 15u) & ~15u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ws.pool + (size_t)(e & ID_MASK) * PIECE_RECS), "r"(bytes) : "memory");
        }
    };
    l2_prefetch(c + NWARPS);
    l2_prefetch(c + 2 * NWARPS);
    while (c < nchunks) {
        const uint32_t cnt = (ent >> IDX_ID_BITS) + 1u;
        const uint32_t cn = c + NWARPS;
        const uint32_t ent_next = cn < nchunks ? __ldg(my_index + cn) : 0u;
        l2_prefetch(c + 3 * NWARPS);
        streamed += cnt;
        if (cnt == PIECE_RECS) {                 // full piece: no per-record bounds checks
#pragma unroll
            for (int q = 0; q < V; ++q) {
                accumulate_rec<MASK, PACKED>(v[q].x, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
                accumulate_rec<MASK, PACKED>(v[q].y, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
                accumulate_rec<MASK, PACKED>(v[q].z, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
                accumulate_rec<MASK, PACKED>(v[q].w, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
            }
        } else {
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const uint32_t i0 = (uint32_t)((q * 32 + lane) * 4);
                if (i0 + 0 < cnt) accumulate_rec<MASK, PACKED>(v[q].x, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
                if (i0 + 1 < cnt) accumulate_rec<MASK, PACKED>(v[q].y, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
                if (i0 + 2 < cnt) accumulate_rec<MASK, PACKED>(v[q].z, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
                if (i0 + 3 < cnt) accumulate_rec<MASK, PACKED>(v[q].w, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
            }
        }
        c = cn;
        ent = ent_next;
        if (c < nchunks) {
            const uint4 *src = reinterpret_cast<const uint4 *>(ws.pool + (size_t)(ent & ID_MASK) * PIECE_RECS);
            const uint32_t cnt2 = (ent >> IDX_ID_BITS) + 1u;
#pragma unroll
            for (int q = 0; q < V; ++q)
                if ((uint32_t)((q * 32 + lane) * 4) < cnt2) v[q] = __ldcs(src + q * 32 + lane);
        }
    }
    return streamed;
}

template <int MASK>
__global__ void __launch_bounds__(RED_THREADS, RED_MIN_CTAS) reduce_tiles_kernel(KParams kp, Ws ws, Outs out) {
    if (gated_off(ws)) return;
    constexpr int NW = popc6(MASK);
    extern __shared__ __align__(16) uint32_t acc[];      // [NW][cells]; plane 0/1 reused as packed/count16
    __shared__ uint4 s_desc[2];                          // schedule entries of the current and the next tile (by parity)
    __shared__ uint32_t s_claim, s_claim_end;            // schedule positions this CTA has claimed and not yet fetched
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TH = 1 << kp.tile_h_log2;
    const int cells = TH << TILE_W_LOG2;
    uint32_t *a_cnt = acc + plane_of(MASK, M_CNT) * cells;
    uint32_t *a_sumi = acc + plane_of(MASK, M_SUMI) * cells;
    uint32_t *a_sumz = acc + plane_of(MASK, M_SUMZ) * cells;
    uint32_t *a_maxi = acc + plane_of(MASK, M_MAXI) * cells;
    uint32_t *a_minz = acc + plane_of(MASK, M_MINZ) * cells;   // holds max(256 - zq): 0 = empty
    uint32_t *a_maxz = acc + plane_of(MASK, M_MAXZ) * cells;
    __shared__ uint32_t s_in[2], s_cnt[2];               // records streamed into / counted in the current tile (by parity)
    __shared__ float s_div255[256];                      // u8 / 255 (one IEEE division each): the proj values
    if (out.proj) for (int i = tid; i < 256; i += RED_THREADS) s_div255[i] = __fdiv_rn((float)i, 255.0f);
    // Every candidate channel value is one byte.  They sit in two registers, A = [max_z | min_z | mean_i | max_i] and
    // B = [0 | 0 | density | mean_z], and ONE byte permute with this selector assembles a pixel's output word
    // (nibble c = which of the eight bytes channel c takes; byte 6 is zero for the channels past nch).
    __shared__ uint32_t s_sel;                           // (in shared memory: not a register held across the streaming loop)
    if (tid == 0) {
        uint32_t sel = 0;
        for (int c = 0; c < 4; ++c) {
            const int ch = c < kp.nch ? kp.ch[c] : -1;
            const uint32_t idx = ch == LM_CH_MAX_I ? 0u : ch == LM_CH_MEAN_I ? 1u : ch == LM_CH_MIN_Z ? 2u : ch == LM_CH_MAX_Z ? 3u
                               : ch == LM_CH_MEAN_Z ? 4u : ch == LM_CH_DENSITY ? 5u : 6u;
            sel |= idx << (4 * c);
        }
        s_sel = sel;
    }
    auto zero_tile = [&]() {
        uint4 *a4 = reinterpret_cast<uint4 *>(acc);
        const int n4 = NW * cells / 4;
        for (int i = tid; i < n4; i += RED_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
    };
    zero_tile();            // the finish pass of every tile leaves the planes zeroed for the next one

    // A tile's schedule entry is fetched while the CTA streams the tile before it (thread 0, asynchronously), and the
    // first pieces of the next tile are pulled into L2 before the current one is finished: a
    // fine raster has ~1 point per cell, a tile's stream is ONE piece per warp, and the chain claim -> entry -> index
    // -> records would otherwise be paid in full, serially, for each of a CTA's ~80 tiles.
    constexpr uint32_t NO_TILE = 0xFFFFFFFFu;
    // schedule entry i -> s_desc[slot], asynchronously (LDGSTS: no register waits for it); complete after sched_wait()
    auto fetch_sched = [&](uint32_t i, int slot) {
        if (i < (uint32_t)kp.T)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&s_desc[slot])), "l"(ws.tile_sched + i) : "memory");
        else
            s_desc[slot] = make_uint4(NO_TILE, 0u, 0u, 0u);
    };
    auto sched_wait = [&]() { asm volatile("cp.async.wait_all;" ::: "memory"); };
    // schedule positions are claimed in runs of `run` (one atomic round trip, paid by warp 0 only, per run): 1 unless a
    // CTA has many tiles to go through (>= 32), up to 8
    const uint32_t run = min(8u, max(1u, (uint32_t)kp.T / (gridDim.x * 16u)));
    if (tid == 0) {
        const uint32_t i0 = atomicAdd(&ws.ctl->tile_counter, run);
        fetch_sched(i0, 0);
        s_claim = i0 + 1u;
        s_claim_end = i0 + run;
        s_in[0] = s_in[1] = s_cnt[0] = s_cnt[1] = 0u;
        sched_wait();
    }
    for (int it = 0;; ++it) {
        const int par = it & 1;
        __syncthreads();                   // s_desc[par] is set; the previous tile's finish pass has zeroed every plane
        if (s_desc[par].x == NO_TILE) break;
        if (tid == 0) {                    // the next tile's entry: in flight while this tile is streamed
            uint32_t c = s_claim;
            if (c == s_claim_end) {
                c = atomicAdd(&ws.ctl->tile_counter, run);
                s_claim_end = c + run;
            }
            s_claim = c + 1u;
            fetch_sched(c, par ^ 1);
        }
        const uint32_t nchunks = s_desc[par].y;
        const uint32_t *my_index = ws.chunk_index + s_desc[par].z;

        // ---- stream the tile's chunks with integer atomics in shared memory; count+sum packed in one
        //      word when possible, redone unpacked if a cell's count field overflowed
        constexpr bool CAN_PACK = (MASK & M_CNT) && (MASK & (M_SUMZ | M_SUMI));
        constexpr bool PACK_Z = CAN_PACK && (MASK & M_SUMZ), PACK_I = CAN_PACK && !PACK_Z;     // as in accumulate_rec
        for (int round = 0; round < 2; ++round) {                // round 1 only after a wrapped packed count
        const bool packed_ok = CAN_PACK && round == 0;
        if (packed_ok) {
            const uint32_t streamed = stream_tile<MASK, CAN_PACK>(ws, my_index, nchunks, warp, lane, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
            if (lane == 0 && streamed) atomicAdd(&s_in[par], streamed);
        } else {
            stream_tile<MASK, false>(ws, my_index, nchunks, warp, lane, a_cnt, a_sumi, a_sumz, a_maxi, a_minz, a_maxz);
        }
        if (round == 0 && tid == 0) sched_wait();
        __syncthreads();
        if (round == 0) {
            if (tid == 0) s_in[par ^ 1] = s_cnt[par ^ 1] = 0u;     // (everybody is past the previous tile's comparison)
            const uint4 dn = s_desc[par ^ 1];                    // the next tile: its first pieces into L2, one per warp
            if (lane == 0 && dn.x != NO_TILE && (uint32_t)warp < dn.y) {
                const uint32_t e = __ldg(ws.chunk_index + dn.z + warp);
                const uint32_t bytes = ((((e >> IDX_ID_BITS) + 1u) * 4u) + 15u) & ~15u;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ws.pool + (size_t)(e & ((1u << IDX_ID_BITS) - 1u)) * PIECE_RECS), "r"(bytes) : "memory");
            }
        }
        uint32_t counted = 0;                                    // packed round: sum of this thread's cell counts
        // (the tile's geometry is derived here, behind the stream, so that no register carries it through the stream)
        const int t = (int)s_desc[par].x;                        // heaviest tiles first
        const int trow = t / kp.tiles_x, tcol = t - trow * kp.tiles_x;
        const int grow0 = trow << kp.tile_h_log2, gcol0 = tcol << TILE_W_LOG2;
        const int nrows = min(TH, kp.H - grow0), ncols = min(TILE_W, kp.W - gcol0);
        const int orow0 = kp.orow + grow0;                       // first row of this tile in the output buffers
        // raw planes: everywhere (band <= 0) or only for tiles touching the halo bands.  In band mode
        // only the planes the requested channels need are accumulated; the others are written as empty.
        const bool want_raw = out.acc != nullptr && (kp.band <= 0 || tile_in_band(kp, t));
        const uint32_t ch_sel = s_sel;

        // ---- finish + write-out in one pass: every thread takes 4 neighbouring cells of a row (a warp
        //      covers a whole 128-cell tile row), derives the channels, assembles the output words in
        //      registers and stores them directly; the planes are zeroed for the next tile on the way
        const size_t gcells = (size_t)kp.oH * kp.W;
        const int nch = kp.nch;
        const size_t row_bytes = (size_t)kp.W * nch;
        const bool full_w = ncols == TILE_W;
        const bool img_fast = full_w && (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(out.image) & 3) == 0;
        const bool c16_fast = full_w && (kp.W & 1) == 0 && (reinterpret_cast<uintptr_t>(out.count16) & 3) == 0;
        // proj planes: [C][oH][W], or [B][C][bH][W] for a batched call (a tile never straddles two samples)
        const size_t pstride = kp.bH ? (size_t)kp.bH * kp.W : gcells;
        const size_t pbase = kp.bH ? (size_t)(orow0 / kp.bH) * (size_t)(nch - 1) * pstride : 0;
        const bool proj_fast = full_w && (kp.W & 3) == 0 && (reinterpret_cast<uintptr_t>(out.proj) & 15) == 0;
        bool overflow = false;
        for (int g = tid; g < cells / 4; g += RED_THREADS) {
            const int lr = g >> (TILE_W_LOG2 - 2), lc0 = (g & (TILE_W / 4 - 1)) << 2;
            const int cell0 = (lr << TILE_W_LOG2) + lc0;
            const uint4 z4 = make_uint4(0, 0, 0, 0);
            uint4 vc = z4, vsi = z4, vsz = z4, vmi = z4, vnz = z4, vxz = z4;
            if (MASK & M_CNT) { vc = *reinterpret_cast<uint4 *>(a_cnt + cell0); *reinterpret_cast<uint4 *>(a_cnt + cell0) = z4; }
            // (the sum that shares the count's word in a packed round has an idle plane: still zero, not read)
            if ((MASK & M_SUMI) && !(packed_ok && PACK_I)) { vsi = *reinterpret_cast<uint4 *>(a_sumi + cell0); *reinterpret_cast<uint4 *>(a_sumi + cell0) = z4; }
            if ((MASK & M_SUMZ) && !(packed_ok && PACK_Z)) { vsz = *reinterpret_cast<uint4 *>(a_sumz + cell0); *reinterpret_cast<uint4 *>(a_sumz + cell0) = z4; }
            if (MASK & M_MAXI) { vmi = *reinterpret_cast<uint4 *>(a_maxi + cell0); *reinterpret_cast<uint4 *>(a_maxi + cell0) = z4; }
            if (MASK & M_MINZ) { vnz = *reinterpret_cast<uint4 *>(a_minz + cell0); *reinterpret_cast<uint4 *>(a_minz + cell0) = z4; }
            if (MASK & M_MAXZ) { vxz = *reinterpret_cast<uint4 *>(a_maxz + cell0); *reinterpret_cast<uint4 *>(a_maxz + cell0) = z4; }
            if (lr >= nrows) continue;                           // rows below the raster: only zeroed
            const uint32_t c4[4] = {vc.x, vc.y, vc.z, vc.w}, si4[4] = {vsi.x, vsi.y, vsi.z, vsi.w};
            const uint32_t sz4[4] = {vsz.x, vsz.y, vsz.z, vsz.w}, mi4[4] = {vmi.x, vmi.y, vmi.z, vmi.w};
            const uint32_t nz4[4] = {vnz.x, vnz.y, vnz.z, vnz.w}, xz4[4] = {vxz.x, vxz.y, vxz.z, vxz.w};
            uint32_t pk[4], k16[4];
            const size_t grow = (size_t)(orow0 + lr) * kp.W + gcol0 + lc0;      // first of the 4 cells in the outputs
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint32_t cnt = c4[e], si = si4[e], sz = sz4[e];
                if (packed_ok) {                                 // unpack [count:12 | sum:20]
                    if (MASK & M_SUMZ) sz = cnt & PK_SUM_MASK; else si = cnt & PK_SUM_MASK;
                    cnt >>= PK_SHIFT;
                    counted += cnt;
                }
                const uint32_t mi = mi4[e], nzr = nz4[e], xz = xz4[e];
                const uint32_t nz = nzr ? 256u - nzr : 0u;      // 0 for an empty cell, with or without a count plane
                overflow |= cnt >= (1u << 24);
                if (want_raw && lc0 + e < ncols) {
                    out.acc[LM_ACC_COUNT * gcells + grow + e] = cnt;
                    out.acc[LM_ACC_SUM_I * gcells + grow + e] = si;
                    out.acc[LM_ACC_SUM_Z * gcells + grow + e] = sz;
                    out.acc[LM_ACC_MAX_I * gcells + grow + e] = mi;
                    out.acc[LM_ACC_MIN_Z * gcells + grow + e] = nzr ? nz : INVALID_U32;
                    out.acc[LM_ACC_MAX_Z * gcells + grow + e] = xz;
                }
                // every candidate channel once, then a select per output byte
                uint32_t mean_i = 0, mean_z = 0;
                if (cnt) {
                    if (packed_ok) {                             // count <= 4095: exact float-reciprocal mean
                        if (MASK & M_SUMI) mean_i = mean_small(si, cnt);
                        if (MASK & M_SUMZ) mean_z = mean_small(sz, cnt);
                    } else {                                     // sums stay < 2^32 while count < 2^24
                        if (MASK & M_SUMI) mean_i = (si + (cnt >> 1)) / cnt;
                        if (MASK & M_SUMZ) mean_z = (sz + (cnt >> 1)) / cnt;
                    }
                }
                const uint32_t dens = cnt < 255u ? cnt : 255u;
                pk[e] = __byte_perm(mi | (mean_i << 8) | (nz << 16) | (xz << 24), mean_z | (dens << 8), ch_sel);
                k16[e] = cnt < 65535u ? cnt : 65535u;
            }
            if (out.image) {
                uint8_t *dst = out.image + grow * nch;
                if (img_fast) {
                    switch (nch) {
                        case 1: store_pixels4<1>(dst, pk); break;
                        case 2: store_pixels4<2>(dst, pk); break;
                        case 3: store_pixels4<3>(dst, pk); break;
                        default: store_pixels4<4>(dst, pk); break;
                    }
                } else {
                    for (int e = 0; e < 4; ++e)
                        if (lc0 + e < ncols)
                            for (int c = 0; c < nch; ++c) dst[e * nch + c] = (uint8_t)(pk[e] >> (8 * c));
                }
            }
            if (out.count16) {
                uint16_t *dst = out.count16 + grow;
                if (c16_fast && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(k16[0] | (k16[1] << 16), k16[2] | (k16[3] << 16));
                } else if (c16_fast) {
                    reinterpret_cast<uint32_t *>(dst)[0] = k16[0] | (k16[1] << 16);
                    reinterpret_cast<uint32_t *>(dst)[1] = k16[2] | (k16[3] << 16);
                } else {
                    for (int e = 0; e < 4; ++e)
                        if (lc0 + e < ncols) dst[e] = (uint16_t)k16[e];
                }
            }
            if (out.proj) {
                for (int c = 0; c < nch; ++c) {
                    float *dst = out.proj + pbase + (size_t)c * pstride + grow;
                    if (proj_fast) {
                        *reinterpret_cast<float4 *>(dst) =
                            make_float4(s_div255[(pk[0] >> (8 * c)) & 0xFFu], s_div255[(pk[1] >> (8 * c)) & 0xFFu],
                                        s_div255[(pk[2] >> (8 * c)) & 0xFFu], s_div255[(pk[3] >> (8 * c)) & 0xFFu]);
                    } else {
                        for (int e = 0; e < 4; ++e)
                            if (lc0 + e < ncols) dst[e] = s_div255[(pk[e] >> (8 * c)) & 0xFFu];
                    }
                }
            }
        }
        if (overflow) atomicOr(&ws.stats->error, (uint32_t)LM_DEV_ERR_CELL_OVERFLOW);
        if (!packed_ok) break;
        // conservation check of the packed round (block-uniform outcome)
        for (int o = 16; o; o >>= 1) counted += __shfl_xor_sync(0xffffffffu, counted, o);
        if (lane == 0 && counted) atomicAdd(&s_cnt[par], counted);
        __syncthreads();
        if (s_cnt[par] == s_in[par]) break;
        __syncthreads();                                         // every thread has compared before the next round
        }
    }
}

// ------------------------------------------------------------------------------------------
// self-test: mean_small == integer (sum + count/2) / count for EVERY count in [1,4095], sum <= 255*count
// ------------------------------------------------------------------------------------------
__global__ void selftest_mean_kernel(unsigned long long *out) {
    unsigned long long bad = 0, n = 0;
    for (uint32_t cnt = 1 + blockIdx.x; cnt <= PK_CNT_MAX; cnt += gridDim.x) {
        for (uint32_t sum = threadIdx.x; sum <= 255u * cnt; sum += blockDim.x) {
            ++n;
            if (mean_small(sum, cnt) != (sum + (cnt >> 1)) / cnt) ++bad;
        }
    }
    for (int o = 16; o; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&out[0], bad);
        atomicAdd(&out[1], n);
    }
}

// ------------------------------------------------------------------------------------------
// self-test: div_const == __fdiv_rn for EVERY binary32 dividend the fast path accepts
// ------------------------------------------------------------------------------------------
__global__ void selftest_div_kernel(float c, float rc, unsigned long long *out) {
    unsigned long long bad = 0, fast = 0;
    const unsigned long long total = 1ull << 32;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((uint32_t)i);
        if (!div_fast_ok(a)) continue;
        ++fast;
        const uint32_t want = __float_as_uint(__fdiv_rn(a, c));
        if (__float_as_uint(div_const(a, c, rc)) != want) ++bad;
        // the packed form the bin kernel uses: this dividend in lane 0, its negation in lane 1
        const float2 q2 = div_const2(make_float2(a, -a), make_float2(c, c), make_float2(rc, rc));
        if (__float_as_uint(q2.x) != want || __float_as_uint(q2.y) != (want ^ 0x80000000u)) ++bad;
    }
    for (int o = 16; o; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        fast += __shfl_xor_sync(0xffffffffu, fast, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&out[0], bad);
        atomicAdd(&out[1], fast);
    }
}

#include "lm_sweep.cuh"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int needed_mask(const lm_bev_params *p, bool count16, bool raw) {
    if (raw) return M_ALL;
    int m = 0;
    for (int c = 0; c < p->n_channels; ++c) {
        switch (p->channels[c]) {
            case LM_CH_MAX_I: m |= M_MAXI; break;
            case LM_CH_MEAN_I: m |= M_SUMI | M_CNT; break;
            case LM_CH_MIN_Z: m |= M_MINZ; break;
            case LM_CH_MAX_Z: m |= M_MAXZ; break;
            case LM_CH_MEAN_Z: m |= M_SUMZ | M_CNT; break;
            case LM_CH_DENSITY: m |= M_CNT; break;
        }
    }
    if (count16) m |= M_CNT;
    return m;
}

// instantiated accumulator sets; the smallest superset of the needed mask is used
constexpr int K_MASKS[] = {M_MAXI, M_CNT | M_MAXI, M_CNT | M_SUMZ | M_MAXI,
                           M_CNT | M_SUMI | M_MAXI | M_MINZ | M_MAXZ, M_ALL};
int pick_mask(int need, bool count16) {
    for (int m : K_MASKS)
        if ((m & need) == need && !(count16 && popc6(m) < 2)) return m;
    return M_ALL;
}
int sm_count() { return lm_sm_count(); }

// tile height: 128 rows for a single plane, else 64 rows (two reduce CTAs per SM up to 3 planes,
// so that one tile's finish/write pass overlaps the other's streaming; one CTA per SM beyond).
// A small raster (one 1152^2 crop = 81 tiles of 128 rows) takes shorter tiles so that the
// persistent reduce CTAs of every SM get a few tiles each.
int tile_h_log2_for(int mask, int height, int width) {
    const int nw = popc6(mask);
    if (const int v = g_tune->tile_h_log2) {                  // tuning knob (5..7); must keep NW planes <= 227 KB
        if (v >= 5 && v <= 7 && nw * (128 << v) * 4 <= 200 * 1024) return v;
    }
    int th = nw <= 1 ? 7 : 6;      // 128 x 64 tiles: 2 CTAs/SM up to 3 planes, 1 CTA/SM (<= 192 KB) up to 6
    const long long tiles_x = (width + TILE_W - 1) >> TILE_W_LOG2;
    const long long want = 4ll * sm_count();
    while (th > 5 && tiles_x * ((height + (1 << th) - 1) >> th) < want) --th;
    return th;
}

int validate(const lm_bev_params *p) {
    if (!p) return fail(LM_ERR_INVALID, "params is NULL");
    if (p->height <= 0 || p->width <= 0) return fail(LM_ERR_INVALID, "height/width must be positive");
    if (p->n_channels < 1 || p->n_channels > 4) return fail(LM_ERR_INVALID, "n_channels must be 1..4");
    for (int c = 0; c < p->n_channels; ++c)
        if (p->channels[c] < 0 || p->channels[c] >= LM_CH__COUNT) return fail(LM_ERR_INVALID, "unknown channel id %d", p->channels[c]);
    if (!(p->img_reso[0] > 0.f) || !(p->img_reso[1] > 0.f) || !(p->ele_reso > 0.f))
        return fail(LM_ERR_INVALID, "resolutions must be positive");
    if (p->inten_min < 0 || p->inten_max > 65535 || p->inten_min >= p->inten_max)
        return fail(LM_ERR_INVALID, "need 0 <= inten_min < inten_max <= 65535");
    const long long lim = 1ll << 24;
    auto absll = [](long long v) { return v < 0 ? -v : v; };
    if (absll(p->row0) + p->height >= lim || absll(p->col0) + p->width >= lim)
        return fail(LM_ERR_INVALID, "window exceeds the exact-float index range 2^24");
    return LM_OK;
}

// L2 policy of the point stream in bin_points (LM_BEV_STREAM_HINT=0/1 overrides; measured in profiles/)
int stream_hint_default() {
    return g_tune->stream_hint & 1;
}

KParams make_kparams(const lm_bev_params *p, int tile_h_log2) {
    KParams k;
    k.H = p->height; k.W = p->width; k.row0 = p->row0; k.col0 = p->col0;
    k.off0 = p->bev_img_offset[0]; k.off1 = p->bev_img_offset[1];
    k.reso0 = p->img_reso[0]; k.reso1 = p->img_reso[1];
    k.zmin = p->local_min_ele; k.zreso = p->ele_reso;
    k.row_lo = (float)p->row0; k.row_hi = (float)(p->row0 + p->height);
    k.col_lo = (float)p->col0; k.col_hi = (float)(p->col0 + p->width);
    k.imin = p->inten_min; k.imax = p->inten_max;
    // exact n/d for n < 2^24 (Granlund-Montgomery): l = ceil(log2 d), m = ceil(2^(24+l)/d) < 2^25,
    // n/d = floor(n*m / 2^(24+l)) = umulhi(n << 8, m) >> l
    const unsigned long long d = (unsigned long long)(p->inten_max - p->inten_min);
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    k.ishift = l;
    k.imagic = (uint32_t)(((1ull << (24 + l)) + d - 1) / d);
    k.rreso0 = 1.0f / k.reso0; k.rreso1 = 1.0f / k.reso1; k.rzreso = 1.0f / k.zreso;
    auto in_range = [](float c) { return c >= 0x1p-20f && c <= 0x1p20f; };
    k.fast_div = in_range(k.reso0) && in_range(k.reso1) && in_range(k.zreso);
    k.nch = p->n_channels;
    for (int c = 0; c < 4; ++c) k.ch[c] = c < p->n_channels ? p->channels[c] : 0;
    k.tile_h_log2 = tile_h_log2;
    k.tiles_x = (p->width + TILE_W - 1) >> TILE_W_LOG2;
    k.tiles_y = (p->height + (1 << tile_h_log2) - 1) >> tile_h_log2;
    k.T = k.tiles_x * k.tiles_y;
    k.oH = p->height;
    k.orow = 0;
    k.band = 0;
    k.bH = 0;
    k.stream_hint = stream_hint_default();
    return k;
}

// cudaFuncSetAttribute + cudaOccupancyMaxActiveBlocksPerMultiprocessor cost microseconds of HOST time per call -- what
// bounds the small configs when no graph is replayed.  Their results depend on (kernel, block size, dynamic shared
// memory, device) only, so they are memoised per process (a cache of pure function results, guarded by a mutex; the
// attribute is raised once per kernel and device to the largest size asked for so far).
struct OccEntry { const void *kernel; size_t smem; int threads, device, occ; };
int cached_occupancy(const void *kernel, int threads, size_t smem, int *occ_out) {
    static std::mutex mu;
    static std::vector<OccEntry> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    size_t attr_set = 0;
    for (const OccEntry &e : cache) {
        if (e.kernel == kernel && e.device == dev) {
            if (e.smem > attr_set) attr_set = e.smem;
            if (e.smem == smem && e.threads == threads) { *occ_out = e.occ; return LM_OK; }
        }
    }
    cudaError_t e = cudaSuccess;
    if (smem > attr_set) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "shared-memory attribute");
        // the same (maximal) carve-out for every kernel of the pipeline: an SM only switches its L1/shared split when idle
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    }
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "occupancy query");
    cache.push_back(OccEntry{kernel, smem, threads, dev, occ});
    *occ_out = occ;
    return LM_OK;
}

size_t bin_smem_bytes(int T) { return BIN_STAGES * (size_t)BIN_BATCH * sizeof(float4) + (size_t)T * (4 + 2 * NSLOT); }
// compact-table bin_points: CT_SLOTS entries of append state + their keys, whatever the raster's tile count
size_t bin_ct_smem_bytes() { return bin_smem_bytes(CT_SLOTS) + (size_t)CT_SLOTS * 4; }
constexpr int CT_MAX_TILES = 1 << 18;     // tiles of a compact-table call (only the tile tables in global memory grow with it)

// chunks per bin CTA when `grid` CTAs share `nb` batches: the chunks its points can fill, one open
// chunk per tile, and the unused local id 0
constexpr long long CHUNKS_PER_BATCH = (BIN_BATCH + CHUNK_RECS - 1) / CHUNK_RECS;
long long bin_region(long long nb, long long grid, int T) {
    const long long per = (nb + grid - 1) / grid;
    return per * CHUNKS_PER_BATCH + T + 1;
}

// upper bound of the bin_points grid for T tiles (shared memory limits the CTAs per SM); the record
// pool reserves one open chunk per (CTA, tile), so the workspace size and the launch both use it
int bin_ctas_bound_smem(size_t smem) {
    const size_t per_sm = 227 * 1024;
    size_t occ = per_sm / (smem + 1024);
    if (occ < 1) occ = 1;
    if (occ > (size_t)LM_BIN_MIN_CTAS + 1) occ = LM_BIN_MIN_CTAS + 1;
    return 148 * (int)occ;
}
int bin_ctas_bound(int T) { return bin_ctas_bound_smem(bin_smem_bytes(T)); }

int max_tiles() {
    if (const int v = g_tune->max_tiles) {                  // test knob: forces the row-window loop on small rasters
        if (v >= 1 && v <= MAX_TILES) return v;
    }
    return MAX_TILES;
}

// rows per launch so that tiles_x * ceil(rows / TH) <= max_tiles(); 0 if even one tile row is too wide
int window_rows(const lm_bev_params *p, int tile_h_log2) {
    const int tiles_x = (p->width + TILE_W - 1) >> TILE_W_LOG2;
    const int tile_rows = max_tiles() / tiles_x;
    if (tile_rows < 1) return 0;
    const long long rows = (long long)tile_rows << tile_h_log2;
    return rows >= p->height ? p->height : (int)rows;
}

struct Layout {
    size_t off_ctl, off_nchunks, off_first, off_cursor, off_order, off_cta, zero_bytes;
    size_t off_meta, off_index, off_pool, off_acc, total;
    uint32_t pool_chunks;
    int bin_ctas;
};
constexpr size_t SW_HEADS_WORDS = (size_t)SW_PRODUCERS * SW_OWNERS;
constexpr size_t SW_MAIL_WORDS = SW_HEADS_WORDS * SW_CAP * 8;
// The sweep's state survives from call to call, so it must not move when a call brings fewer points than the
// workspace was sized for (the two-pass layout depends on n_points): it occupies the last SW_REGION_BYTES of the
// workspace, [persistent block | consumer heads | mailboxes], each 256-byte aligned.
constexpr size_t SW_OFF_HEADS = 256, SW_OFF_MAIL = SW_OFF_HEADS + (SW_HEADS_WORDS * 4 + 255) / 256 * 256;
constexpr size_t SW_REGION_BYTES = SW_OFF_MAIL + (SW_MAIL_WORDS * 4 + 255) / 256 * 256 + 256;
size_t sweep_region(size_t workspace_bytes) { return (workspace_bytes - SW_REGION_BYTES + 255) / 256 * 256; }

int bin_ctas_for(long long n) {
    const long long nb = (n + BIN_BATCH - 1) / BIN_BATCH;
    return (int)(nb < 1 ? 1 : (nb < MAX_BIN_CTAS ? nb : MAX_BIN_CTAS));
}

// Binned workspace: [stats | ctl | tile tables | chunks used per bin CTA] (zeroed per call)
//                   [chunk side table | chunk index | record pool]
// ct: layout of the compact-table pass -- T only sizes the tile tables; a bin CTA holds at most CT_SLOTS open chunks
int make_layout(const lm_bev_params *p, long long n, int algo, int T, Layout *L, bool ct = false) {
    memset(L, 0, sizeof(*L));
    size_t o = 0;
    o = align_up(o + sizeof(lm_bev_stats), 64);
    L->off_ctl = o; o = align_up(o + sizeof(Ctl), 256);
    if (algo == LM_ALGO_DIRECT) {
        L->zero_bytes = o;
        L->off_acc = o;
        o += (size_t)LM_ACC_PLANES * p->height * p->width * sizeof(uint32_t);
        L->total = align_up(o, 256);
        return LM_OK;
    }
    L->bin_ctas = bin_ctas_for(n);
    L->off_nchunks = o; o = align_up(o + (size_t)T * 4, 256);
    L->off_first = o;   o = align_up(o + (size_t)T * 4, 256);
    L->off_cursor = o;  o = align_up(o + (size_t)T * 4, 256);
    L->off_order = o;   o = align_up(o + (size_t)T * sizeof(uint4), 256);
    const long long gb = ct ? bin_ctas_bound_smem(bin_ct_smem_bytes()) : bin_ctas_bound(T);
    const int T_open = ct && T > CT_SLOTS ? CT_SLOTS : T;      // open chunks a bin CTA can hold
    L->off_cta = o;     o = align_up(o + (size_t)gb * 4, 256);
    L->zero_bytes = o;
    // every bin CTA owns a region of the pool (bin_region): whatever grid <= gb the launch ends up with,
    // grid * region <= batches + gb * (T + 2).  A batched call adds one partial batch per sample.
    const unsigned long long nbb = (unsigned long long)(n / BIN_BATCH) + 1ull + MAX_BATCH;
    const unsigned long long chunks = nbb * CHUNKS_PER_BATCH + (unsigned long long)gb * ((unsigned long long)T_open + 2ull) + 1ull;
    if (chunks * HALVES >= (1ull << IDX_ID_BITS))     // piece ids share a word with the count; record indices are 32-bit
        return fail(LM_ERR_UNSUPPORTED, "record pool exceeds 2^23 chunks: shard the call (fewer points or a smaller row window)");
    L->pool_chunks = (uint32_t)chunks;
    L->off_meta = o;  o = align_up(o + (size_t)chunks * sizeof(uint2), 256);
    L->off_index = o; o = align_up(o + (size_t)chunks * HALVES * 4, 256);
    L->off_pool = o;  o = align_up(o + (size_t)chunks * CHUNK_RECS * 4, 256);
    if (algo == LM_ALGO_SWEEP) o += SW_REGION_BYTES;     // the sweep's state: the LAST bytes of the workspace (sweep_region)
    L->total = o;
    return LM_OK;
}

// the sweep handles: one launch window, image / count16 / proj outputs, channels out of {max_i, mean_z, density},
// rasters up to SW_MAX_LG * 4 * 148 columns, and needs all its CTAs resident at once
bool sweep_mask_ok(int mask) { return mask == M_MAXI || mask == (M_CNT | M_MAXI) || mask == (M_CNT | M_SUMZ | M_MAXI); }
bool sweep_eligible(const lm_bev_params *p, const lm_bev_outputs *out, int mask, int n_win, bool las) {
    if (las || n_win != 1 || out->acc_dev) return false;
    if (!sweep_mask_ok(mask)) return false;
    for (int c = 0; c < p->n_channels; ++c)
        if (p->channels[c] != LM_CH_MAX_I && p->channels[c] != LM_CH_MEAN_Z && p->channels[c] != LM_CH_DENSITY) return false;
    if ((p->width + 3) / 4 > SW_MAX_LG * SW_OWNERS) return false;
    return true;
}
uint32_t sweep_magic(const void *w, size_t total) {
    return 0x5EEB0001u ^ (uint32_t)(reinterpret_cast<uintptr_t>(w) >> 8) ^ (uint32_t)total * 2654435761u;
}
SweepWs make_sweep_ws(unsigned char *w, size_t workspace_bytes, Ctl *ctl) {
    unsigned char *r = w + sweep_region(workspace_bytes);
    SweepWs sw;
    sw.persist = reinterpret_cast<SweepPersist *>(r);
    sw.heads = reinterpret_cast<uint32_t *>(r + SW_OFF_HEADS);
    sw.mail = reinterpret_cast<uint4 *>(r + SW_OFF_MAIL);
    sw.fail = &ctl->sweep_fail;
    sw.next_batch = &ctl->next_batch;
    sw.stats = reinterpret_cast<lm_bev_stats *>(w);
    sw.magic = sweep_magic(w, workspace_bytes);
    return sw;
}
// fits = every CTA of the sweep can be resident at once (producers and consumers wait for each other);
// launch = false only asks that question
template <int MASK>
cudaError_t launch_sweep(const KParams &kp, const float4 *pts, long long n, const SweepWs &sw, const Outs &o, int sms, bool *fits,
                         bool launch, cudaStream_t st) {
    int occ = 0;
    if (cached_occupancy(reinterpret_cast<const void *>(sweep_kernel<MASK>), SW_THREADS, SW_SMEM, &occ)) return cudaErrorUnknown;
    *fits = (long long)occ * sms >= SW_GRID;
    if (!*fits || !launch) return cudaSuccess;
    sweep_kernel<MASK><<<SW_GRID, SW_THREADS, SW_SMEM, st>>>(kp, pts, n, sw, o);
    return cudaGetLastError();
}
cudaError_t launch_sweep_mask(int mask, const KParams &kp, const float4 *pts, long long n, const SweepWs &sw, const Outs &o, int sms,
                              bool *fits, bool launch, cudaStream_t st) {
    switch (mask) {
        case M_MAXI: return launch_sweep<M_MAXI>(kp, pts, n, sw, o, sms, fits, launch, st);
        case M_CNT | M_MAXI: return launch_sweep<M_CNT | M_MAXI>(kp, pts, n, sw, o, sms, fits, launch, st);
        default: return launch_sweep<M_CNT | M_SUMZ | M_MAXI>(kp, pts, n, sw, o, sms, fits, launch, st);
    }
}

// Grid and chunk region of one bin launch (the same values in every stage-split call of a raster):
// persistent, one wave of resident CTAs, each owning a contiguous range of the nb batches.
int bin_geometry(const void *kernel, size_t smem, long long nb, int T, int tiles_x, int sms, const Layout &L, Ws *ws, int *grid_out,
                 bool ct = false) {
    const int bound = ct ? bin_ctas_bound_smem(bin_ct_smem_bytes()) : bin_ctas_bound(T);
    if (ct && T > CT_SLOTS) T = CT_SLOTS;
    if (smem > 220 * 1024) return fail(LM_ERR_UNSUPPORTED, "%zu bytes of shared memory per bin CTA: too many tiles / too long records", smem);
    int occ = 1;
    if (int rc = cached_occupancy(kernel, BIN_THREADS, smem, &occ)) return rc;
    // A raster wider than four crops is a multi-road scene: the scan visits one road at a time and its
    // stray returns spread over every tile column, so every (CTA, tile) pair keeps a slowly filling open
    // chunk.  Fewer, faster-moving CTAs fill those sectors sooner: 3 per SM measured best on the
    // 11520-column strips of config 3 (profiles/r01_v11_hint_sweep.txt), the full wave on config 2.
    const int hw_occ = occ;
    if (tiles_x > 36 && occ > 3) occ = 3;
    if (const int v = g_tune->bin_ctas_per_sm) {                  // tuning knob, 1 .. what the hardware holds
        if (v >= 1) occ = v < hw_occ ? v : hw_occ;
    }
    long long grid = 0, region = 0;
    // The pool was sized for the layout's own tile count; a launch with FEWER tiles (the shorter last row window of a
    // raster) may fit more CTAs per SM and so reserve more open chunks than that: run it with fewer CTAs then.
    for (occ = occ < 1 ? 1 : occ;; --occ) {
        grid = (long long)sms * occ;
        if (grid > bound) grid = bound;
        if (grid > nb) grid = nb;
        if (grid < 1) grid = 1;
        region = bin_region(nb, grid, T);
        if (occ == 1 || (region <= (long long)MAX_REGION && (unsigned long long)grid * (unsigned long long)region <= L.pool_chunks)) break;
    }
    if (region > (long long)MAX_REGION)
        return fail(LM_ERR_UNSUPPORTED, "%lld chunks per bin CTA exceed the 16-bit local chunk ids: shard the call", region);
    if ((unsigned long long)grid * (unsigned long long)region > L.pool_chunks)
        return fail(LM_ERR_WORKSPACE, "record pool of %u chunks < %lld x %lld", L.pool_chunks, grid, region);
    ws->region = (uint32_t)region;
    ws->bin_grid = (uint32_t)grid;
    *grid_out = (int)grid;
    return LM_OK;
}

template <int MASK>
cudaError_t launch_reduce(const KParams &kp, const Ws &ws, const Outs &o, int sms, cudaStream_t st) {
    const size_t smem = (size_t)popc6(MASK) * ((size_t)TILE_W << kp.tile_h_log2) * 4;
    int occ = 1;
    if (cached_occupancy(reinterpret_cast<const void *>(reduce_tiles_kernel<MASK>), RED_THREADS, smem, &occ)) return cudaErrorUnknown;
    if (occ < 1) occ = 1;
    if (const int v = g_tune->red_ctas_per_sm) {                  // tuning knob
        if (v >= 1 && v < occ) occ = v;
    }
    const int grid = kp.T < sms * occ ? kp.T : sms * occ;
    reduce_tiles_kernel<MASK><<<grid, RED_THREADS, smem, st>>>(kp, ws, o);
    return cudaGetLastError();
}

cudaError_t launch_reduce_mask(int mask, const KParams &kp, const Ws &ws, const Outs &o, int sms, cudaStream_t st) {
    switch (mask) {
        case M_MAXI: return launch_reduce<M_MAXI>(kp, ws, o, sms, st);
        case M_CNT | M_MAXI: return launch_reduce<M_CNT | M_MAXI>(kp, ws, o, sms, st);
        case M_CNT | M_SUMZ | M_MAXI: return launch_reduce<M_CNT | M_SUMZ | M_MAXI>(kp, ws, o, sms, st);
        case M_CNT | M_SUMI | M_MAXI | M_MINZ | M_MAXZ:
            return launch_reduce<M_CNT | M_SUMI | M_MAXI | M_MINZ | M_MAXZ>(kp, ws, o, sms, st);
        default: return launch_reduce<M_ALL>(kp, ws, o, sms, st);
    }
}

// ---- compact-table pass (bin_points_ct_kernel): when does a call take it, and with which tiles
// tile height of the compact-table pass: its bin kernel does not care how many tiles there are, so rasters with four
// planes or more take 32-row tiles (two or three reduce CTAs per SM instead of one)
int ct_tile_h(int mask, const lm_bev_params *p) {
    const int nw = popc6(mask);
    if (const int v = g_tune->tile_h_log2) {
        if (v >= 5 && v <= 7 && nw * (128 << v) * 4 <= 200 * 1024) return v;
    }
    return nw >= 4 ? 5 : tile_h_log2_for(mask, p->height, p->width);
}
// T_direct: tiles per launch of the direct-indexed kernels for this raster (one row window)
bool ct_wanted(const lm_bev_params *p, int mask, int T_direct) {
    const int mode = g_tune->bin_compact_table;           // 0 auto, 1 always, -1 never
    if (mode < 0) return false;
    if (make_kparams(p, ct_tile_h(mask, p)).T > CT_MAX_TILES) return false;
    if (mode > 0) return true;
    return bin_ctas_bound(T_direct) < 148 * 3;            // direct indexing would run fewer than three bin CTAs per SM
}
Ws bind_ws(unsigned char *w, const Layout &L) {
    Ws ws;
    ws.stats = reinterpret_cast<lm_bev_stats *>(w);
    ws.ctl = reinterpret_cast<Ctl *>(w + L.off_ctl);
    ws.tile_nchunks = reinterpret_cast<uint32_t *>(w + L.off_nchunks);
    ws.tile_first = reinterpret_cast<uint32_t *>(w + L.off_first);
    ws.tile_cursor = reinterpret_cast<uint32_t *>(w + L.off_cursor);
    ws.tile_sched = reinterpret_cast<uint4 *>(w + L.off_order);
    ws.cta_chunks = reinterpret_cast<uint32_t *>(w + L.off_cta);
    ws.chunk_meta = reinterpret_cast<uint2 *>(w + L.off_meta);
    ws.chunk_index = reinterpret_cast<uint32_t *>(w + L.off_index);
    ws.pool = reinterpret_cast<uint32_t *>(w + L.off_pool);
    ws.acc = nullptr;
    ws.pool_chunks = L.pool_chunks;
    ws.gate = nullptr;
    ws.region = 1;
    ws.bin_grid = 0;
    ws.scan_in_bin = false;
    return ws;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int lm_bev_abi_version(void) { return LM_BEV_ABI_VERSION; }
const char *lm_bev_last_error(void) { return g_err; }

int lm_bev_workspace_bytes(const lm_bev_params *p, int64_t n_points, int algo, const lm_bev_outputs *out,
                           size_t *bytes) {
    int rc = validate(p);
    if (rc) return rc;
    if (!bytes || n_points < 0) return fail(LM_ERR_INVALID, "bytes is NULL or n_points < 0");
    if (algo != LM_ALGO_BINNED && algo != LM_ALGO_DIRECT && algo != LM_ALGO_SWEEP) return fail(LM_ERR_INVALID, "unknown algo %d", algo);
    // the tile height depends on the accumulator planes the outputs need.  Without an output set the bound has to
    // hold for every one: the pool reserves bin_ctas_bound(T) * (T + 2) chunks, which is NOT monotonic in the tile
    // count T (the CTAs per SM drop as T grows), so take the maximum over the three tile heights
    int th_lo = 5, th_hi = 7, mask_of_out = 0;
    if (out) {
        const bool want16 = out->count16_dev != nullptr;
        const bool banded = out->acc_dev != nullptr && out->acc_band > 0 &&
                            (out->image_dev || out->count16_dev || out->proj_dev);
        mask_of_out = pick_mask(needed_mask(p, want16, out->acc_dev != nullptr && !banded), want16);
        th_lo = th_hi = tile_h_log2_for(mask_of_out, p->height, p->width);
    }
    size_t best = 0;
    for (int th = th_lo; th <= th_hi; ++th) {
        KParams k = make_kparams(p, th);
        if (algo != LM_ALGO_DIRECT) {          // a raster with too many tiles runs as row windows: size for one window
            const int wrows = window_rows(p, th);
            if (wrows == 0) return fail(LM_ERR_UNSUPPORTED, "raster too wide: %d tiles per tile row > %d", k.tiles_x, max_tiles());
            lm_bev_params pw = *p;
            pw.height = wrows;
            k = make_kparams(&pw, th);
        }
        Layout L;
        rc = make_layout(p, n_points, algo, k.T, &L);
        if (rc) return rc;
        if (L.total > best) best = L.total;
        // the compact-table pass of the same call (ct_wanted; k.T is the tile count of one direct-indexed launch):
        // its own tile height, no row windows
        const int mode = algo == LM_ALGO_BINNED ? g_tune->bin_compact_table : -1;
        if (mode > 0 || (mode == 0 && bin_ctas_bound(k.T) < 148 * 3)) {
            int tc_lo = 5, tc_hi = 7;
            if (out) tc_lo = tc_hi = ct_tile_h(mask_of_out, p);
            for (int tc = tc_lo; tc <= tc_hi; ++tc) {
                const KParams kc = make_kparams(p, tc);
                if (kc.T <= CT_MAX_TILES && make_layout(p, n_points, algo, kc.T, &L, true) == LM_OK && L.total > best) best = L.total;
            }
        }
    }
    *bytes = best;
    return LM_OK;
}

int lm_bev_workspace_init(const lm_bev_params *p, int64_t n_points, int algo, const lm_bev_outputs *out,
                          void *workspace_dev, size_t workspace_bytes, void *stream) {
    size_t need = 0;
    int rc = lm_bev_workspace_bytes(p, n_points, algo, out, &need);
    if (rc) return rc;
    if (!workspace_dev || (reinterpret_cast<uintptr_t>(workspace_dev) & 255))
        return fail(LM_ERR_WORKSPACE, "workspace_dev is NULL or not 256-byte aligned");
    if (workspace_bytes < need) return fail(LM_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, need);
    if (algo != LM_ALGO_SWEEP) return LM_OK;
    unsigned char *w = static_cast<unsigned char *>(workspace_dev);
    const SweepWs sw = make_sweep_ws(w, workspace_bytes, reinterpret_cast<Ctl *>(w + align_up(sizeof(lm_bev_stats), 64)));
    sweep_init_kernel<<<sm_count() * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sw, SW_HEADS_WORDS, SW_MAIL_WORDS);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "workspace_init launch");
}

int lm_bev_sweep_state_offset(size_t workspace_bytes, size_t *offset) {
    if (!offset || workspace_bytes < SW_REGION_BYTES + 4096) return fail(LM_ERR_INVALID, "not an LM_ALGO_SWEEP workspace size");
    *offset = sweep_region(workspace_bytes);
    return LM_OK;
}

int lm_bev_rasterize(const lm_bev_params *p, const float *points_dev, int64_t n_points, int algo,
                     void *workspace_dev, size_t workspace_bytes, const lm_bev_outputs *out, void *stream) {
    return lm_bev_rasterize_stages(p, points_dev, n_points, algo, workspace_dev, workspace_bytes, out, stream,
                                   LM_STAGE_ALL);
}

static int rasterize_impl(const lm_bev_params *p, const float *points_dev, int64_t n_points, int algo,
                          void *workspace_dev, size_t workspace_bytes, const lm_bev_outputs *out, void *stream,
                          int stages, bool keep_stats, const LasXform *las = nullptr);

int lm_bev_rasterize_stages(const lm_bev_params *p, const float *points_dev, int64_t n_points, int algo,
                            void *workspace_dev, size_t workspace_bytes, const lm_bev_outputs *out, void *stream,
                            int stages) {
    return rasterize_impl(p, points_dev, n_points, algo, workspace_dev, workspace_bytes, out, stream, stages, false);
}

// keep_stats: lm_bev_stats keeps accumulating (a batched call that runs sample by sample)
// las: points_dev is the point-data block of a LAS file, decoded inside bin_points (lm_bev_rasterize_las)
static int rasterize_impl(const lm_bev_params *p, const float *points_dev, int64_t n_points, int algo,
                          void *workspace_dev, size_t workspace_bytes, const lm_bev_outputs *out, void *stream,
                          int stages, bool keep_stats, const LasXform *las) {
    int rc = validate(p);
    if (rc) return rc;
    if (n_points < 0 || (n_points > 0 && !points_dev)) return fail(LM_ERR_INVALID, "points_dev is NULL");
    if (reinterpret_cast<uintptr_t>(points_dev) & 15) return fail(LM_ERR_INVALID, "points_dev must be 16-byte aligned");
    if (!out || (!out->image_dev && !out->count16_dev && !out->proj_dev && !out->acc_dev))
        return fail(LM_ERR_INVALID, "no output buffer requested");
    if (!workspace_dev || (reinterpret_cast<uintptr_t>(workspace_dev) & 255))
        return fail(LM_ERR_WORKSPACE, "workspace_dev is NULL or not 256-byte aligned");
    if (algo != LM_ALGO_BINNED && algo != LM_ALGO_DIRECT && algo != LM_ALGO_SWEEP) return fail(LM_ERR_INVALID, "unknown algo %d", algo);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char *w = static_cast<unsigned char *>(workspace_dev);
    const Outs o = {out->image_dev, out->count16_dev, out->proj_dev, out->acc_dev, out->acc_band};
    const int sms = sm_count();

    if (algo == LM_ALGO_DIRECT) {
        const KParams kp = make_kparams(p, 7);
        Layout L;
        rc = make_layout(p, n_points, algo, kp.T, &L);
        if (rc) return rc;
        // with caller-provided accumulators the workspace only carries stats
        const size_t need = out->acc_dev ? L.off_acc : L.total;
        if (workspace_bytes < need) return fail(LM_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, need);
        cudaError_t e = cudaSuccess;
        lm_bev_stats *stats = reinterpret_cast<lm_bev_stats *>(w);
        uint32_t *acc = out->acc_dev ? out->acc_dev : reinterpret_cast<uint32_t *>(w + L.off_acc);
        const size_t cells = (size_t)p->height * p->width;
        if (stages & LM_STAGE_BIN) {
            e = cudaMemsetAsync(w, 0, L.zero_bytes, st);
            if (e != cudaSuccess) return cuda_fail(e, "memset");
            acc_init_kernel<<<sms * 8, 256, 0, st>>>(acc, cells);
            if (n_points > 0) direct_accumulate_kernel<<<sms * 8, 256, 0, st>>>(kp, reinterpret_cast<const float4 *>(points_dev), n_points, acc, stats);
        }
        if ((stages & LM_STAGE_REDUCE) && (o.image || o.count16 || o.proj)) {
            Outs fo = o;
            fo.acc = nullptr;
            finalize_kernel<<<sms * 8, 256, 0, st>>>(kp, acc, 0, p->height, fo);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "direct path launch");
        return LM_OK;
    }

    // ---- binned path (row windows when the raster has more tiles than one launch handles)
    const bool want16 = out->count16_dev != nullptr;
    // raw accumulators on halo bands only (acc_band > 0): the kernel keeps just the planes the requested
    // channels need and also emits them raw for the band tiles; acc_band <= 0 accumulates all six planes
    const bool banded = out->acc_dev != nullptr && out->acc_band > 0 && (o.image || o.count16 || o.proj);
    const int mask = pick_mask(needed_mask(p, want16, out->acc_dev != nullptr && !banded), want16);
    const int th = tile_h_log2_for(mask, p->height, p->width);
    const int wrows = window_rows(p, th);
    if (wrows == 0) return fail(LM_ERR_UNSUPPORTED, "raster too wide: more than %d tiles per tile row", max_tiles());
    const int n_win = (p->height + wrows - 1) / wrows;
    if (n_win > 1 && stages != LM_STAGE_ALL)
        return fail(LM_ERR_UNSUPPORTED, "stage-split calls need a raster that fits one launch (%d row windows here)", n_win);
    lm_bev_params pw = *p;
    pw.height = wrows;
    Layout L;
    rc = make_layout(p, n_points, algo, make_kparams(&pw, th).T, &L);
    if (rc) return rc;
    if (workspace_bytes < L.total) return fail(LM_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, L.total);
    Ws ws = bind_ws(w, L);

    // ---- compact-table pass: the whole raster in one launch set, four bin CTAs per SM whatever the tile count.  The
    //      direct-indexed kernels below stay queued behind it, gated on stats->ct_overflow (a bin CTA met more tiles
    //      than its table holds: a cloud in no spatial order), and redo the raster in that case.
    if (algo == LM_ALGO_BINNED && stages == LM_STAGE_ALL && !keep_stats && !las && n_points > 0 &&
        ct_wanted(p, mask, make_kparams(&pw, th).T)) {
        KParams kp = make_kparams(p, ct_tile_h(mask, p));
        kp.band = banded ? out->acc_band : 0;
        Layout Lc;
        rc = make_layout(p, n_points, algo, kp.T, &Lc, true);
        if (rc) return rc;
        if (workspace_bytes < Lc.total) return fail(LM_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, Lc.total);
        Ws wc = bind_ws(w, Lc);
        wc.scan_in_bin = n_points <= SCAN_IN_BIN_MAX_POINTS;
        const long long nb = (n_points + BIN_BATCH - 1) / BIN_BATCH;
        int grid = 0;
        rc = bin_geometry((const void *)bin_points_ct_kernel, bin_ct_smem_bytes(), nb, kp.T, kp.tiles_x, sms, Lc, &wc, &grid, true);
        if (rc) return rc;
        cudaError_t e = cudaMemsetAsync(w, 0, Lc.zero_bytes, st);
        if (e != cudaSuccess) return cuda_fail(e, "memset");
        bin_points_ct_kernel<<<grid, BIN_THREADS, bin_ct_smem_bytes(), st>>>(kp, reinterpret_cast<const float4 *>(points_dev), n_points, wc);
        if (!wc.scan_in_bin) scan_tiles_kernel<<<1, 1024, 0, st>>>(wc, kp);
        index_chunks_kernel<<<sms * 4, 256, 0, st>>>(wc);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "compact-table bin/index launch");
        e = launch_reduce_mask(mask, kp, wc, o, sms, st);
        if (e != cudaSuccess) return cuda_fail(e, "compact-table reduce launch");
        ct_epilogue_kernel<<<1, 1, 0, st>>>(wc.stats);
        ws.gate = &ws.stats->ct_overflow;
        keep_stats = true;                     // the gated kernels' memsets leave the stats block (and the gate in it) alone
    }
    bool sweep = algo == LM_ALGO_SWEEP && sweep_eligible(p, out, mask, n_win, las != nullptr);
    SweepWs sw;
    if (sweep) {
        sw = make_sweep_ws(w, workspace_bytes, ws.ctl);
        cudaError_t e = launch_sweep_mask(mask, make_kparams(&pw, th), nullptr, 0, sw, o, sms, &sweep, false, st);
        if (e != cudaSuccess) return cuda_fail(e, "sweep occupancy");
    }
    // behind a sweep the two-pass kernels return at once unless ctl->sweep_fail says that it gave up (or was skipped)
    if (sweep) ws.gate = &ws.ctl->sweep_fail;

    for (int win = 0; win < n_win; ++win) {
        // the window is an integer sub-window of the same global grid: bit-identical to the one-piece raster
        pw = *p;
        pw.row0 = p->row0 + win * wrows;
        pw.height = (win + 1) * wrows <= p->height ? wrows : p->height - win * wrows;
        KParams kp = make_kparams(&pw, th);
        kp.oH = p->height;
        kp.orow = win * wrows;
        kp.band = banded ? out->acc_band : 0;
        cudaError_t e = cudaSuccess;
        const long long nb = (n_points + BIN_BATCH - 1) / BIN_BATCH;
        const size_t smem = las ? BIN_STAGES * (size_t)las_stage_bytes((uint32_t)las->record_length, BIN_BATCH) + (size_t)kp.T * (4 + 2 * NSLOT)
                                : bin_smem_bytes(kp.T);
        int grid = 0;
        ws.region = 1;
        ws.bin_grid = 0;
        ws.scan_in_bin = n_points > 0 && n_points <= SCAN_IN_BIN_MAX_POINTS;
        if (n_points > 0) {
            rc = bin_geometry(las ? (const void *)bin_points_las_kernel : (const void *)bin_points_kernel, smem, nb, kp.T, kp.tiles_x, sms, L, &ws, &grid);
            if (rc) return rc;
        }
        if (sweep && (stages & LM_STAGE_SWEEP)) {
            e = cudaMemsetAsync(w, 0, L.zero_bytes, st);
            if (e != cudaSuccess) return cuda_fail(e, "memset");
            bool fits = true;
            e = launch_sweep_mask(mask, kp, reinterpret_cast<const float4 *>(points_dev), n_points, sw, o, sms, &fits, true, st);
            if (e != cudaSuccess) return cuda_fail(e, "sweep launch");
            sweep_epilogue_kernel<<<1, 1, 0, st>>>(sw);
        }
        if (stages & LM_STAGE_BIN) {
            if (!sweep) {      // (the sweep stage has zeroed the tables already, and left its verdict in them)
                // stats (first 64 bytes) accumulate over the windows; everything else restarts
                const size_t skip = win == 0 && !keep_stats ? 0 : L.off_ctl;
                e = cudaMemsetAsync(w + skip, 0, L.zero_bytes - skip, st);
                if (e != cudaSuccess) return cuda_fail(e, "memset");
            }
            if (n_points > 0 && las)
                bin_points_las_kernel<<<grid, BIN_THREADS, smem, st>>>(kp, *las, reinterpret_cast<const unsigned char *>(points_dev), n_points, ws);
            else if (n_points > 0)
                bin_points_kernel<<<grid, BIN_THREADS, smem, st>>>(kp, reinterpret_cast<const float4 *>(points_dev), n_points, ws);
        }
        if (stages & LM_STAGE_INDEX) {      // (small calls: the tile tables were built by the last bin CTA)
            if (!ws.scan_in_bin) scan_tiles_kernel<<<1, 1024, 0, st>>>(ws, kp);
            if (n_points > 0) index_chunks_kernel<<<sms * 4, 256, 0, st>>>(ws);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "bin/index launch");
        if (!(stages & LM_STAGE_REDUCE)) continue;
        e = launch_reduce_mask(mask, kp, ws, o, sms, st);
        if (e != cudaSuccess) return cuda_fail(e, "reduce_tiles launch");
    }
    return LM_OK;
}

// ---- LAS front end (include/lm_las.h)
namespace {
int las_xform(const lm_las_xform *x, LasXform *o) {
    if (!x) return fail(LM_ERR_INVALID, "lm_las_xform is NULL");
    if (x->record_length < 14 || x->record_length > 100)
        return fail(LM_ERR_INVALID, "record_length %d outside 14..100", x->record_length);
    for (int k = 0; k < 3; ++k) {
        o->scale[k] = x->scale[k]; o->offset[k] = x->offset[k];
        o->read_offset[k] = x->las_read_offset[k]; o->t[k] = x->translation[k];
    }
    for (int k = 0; k < 9; ++k) o->m[k] = x->rot[k];
    o->record_length = x->record_length;
    return LM_OK;
}
}  // namespace

int lm_las_decode(const uint8_t *records_dev, int64_t n_points, const lm_las_xform *x, float *points_dev, void *stream) {
    LasXform xf;
    int rc = las_xform(x, &xf);
    if (rc) return rc;
    if (n_points < 0 || (n_points > 0 && (!records_dev || !points_dev))) return fail(LM_ERR_INVALID, "records_dev/points_dev is NULL");
    if ((reinterpret_cast<uintptr_t>(records_dev) & 15) || (reinterpret_cast<uintptr_t>(points_dev) & 15))
        return fail(LM_ERR_INVALID, "records_dev and points_dev must be 16-byte aligned");
    if (n_points == 0) return LM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = 2 * (size_t)las_stage_bytes((uint32_t)xf.record_length, LAS_TILE);
    cudaError_t e = cudaFuncSetAttribute(las_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "las_decode smem attribute");
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, las_decode_kernel, LAS_THREADS, smem);
    if (e != cudaSuccess) return cuda_fail(e, "las_decode occupancy");
    long long grid = (long long)sm_count() * (occ < 1 ? 1 : occ);
    const long long nb = (n_points + LAS_TILE - 1) / LAS_TILE;
    if (grid > nb) grid = nb;
    las_decode_kernel<<<(int)grid, LAS_THREADS, smem, st>>>(xf, records_dev, n_points, reinterpret_cast<float4 *>(points_dev));
    e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "las_decode launch");
}

int lm_bev_rasterize_las(const lm_bev_params *p, const uint8_t *records_dev, int64_t n_points, const lm_las_xform *x,
                         void *workspace_dev, size_t workspace_bytes, const lm_bev_outputs *out, void *stream) {
    LasXform xf;
    int rc = las_xform(x, &xf);
    if (rc) return rc;
    return rasterize_impl(p, reinterpret_cast<const float *>(records_dev), n_points, LM_ALGO_BINNED, workspace_dev,
                          workspace_bytes, out, stream, LM_STAGE_ALL, false, &xf);
}

// ---- batched call: up to MAX_BATCH equally-shaped samples per launch set, stacked along the rows
namespace {
// samples per launch set: bounded by the batch table, by the tiles one launch handles and (for the
// stacked u32 cell indices of the outputs) by nothing else; 0 = cannot batch (sample too big)
int batch_group(const lm_bev_params *p, int th, int n_samples) {
    const long long tiles = (long long)((p->width + TILE_W - 1) >> TILE_W_LOG2) * ((p->height + (1 << th) - 1) >> th);
    long long g = max_tiles() / tiles;
    if (g > MAX_BATCH) g = MAX_BATCH;
    if (g > n_samples) g = n_samples;
    if ((long long)p->height * g >= (1ll << 24)) g = ((1ll << 24) - 1) / p->height;
    return (int)g;
}
int batch_tile_h(const lm_bev_params *p, int mask, int n_samples) {
    // the tile height is picked for the stacked raster; a tile must not straddle two samples
    const long long stacked = (long long)p->height * (n_samples < MAX_BATCH ? n_samples : MAX_BATCH);
    int th = tile_h_log2_for(mask, (int)(stacked < (1 << 24) ? stacked : (1 << 24) - 1), p->width);
    while (th > 5 && (p->height & ((1 << th) - 1))) --th;
    return th;
}
}  // namespace

int lm_bev_workspace_bytes_batch(const lm_bev_params *p, int32_t n_samples, int64_t n_points_total,
                                 const lm_bev_outputs *out, size_t *bytes) {
    int rc = validate(p);
    if (rc) return rc;
    if (!bytes || n_points_total < 0 || n_samples < 1) return fail(LM_ERR_INVALID, "bytes is NULL, n_points_total < 0 or n_samples < 1");
    if (!out) return fail(LM_ERR_INVALID, "out is NULL (the batched call sizes its tiles for the output set)");
    const bool want16 = out->count16_dev != nullptr;
    const int mask = pick_mask(needed_mask(p, want16, false), want16);
    const int th = batch_tile_h(p, mask, n_samples);
    if (p->height & ((1 << th) - 1)) {        // samples are rasterised one by one through lm_bev_rasterize
        return lm_bev_workspace_bytes(p, n_points_total, LM_ALGO_BINNED, out, bytes);
    }
    const int g = batch_group(p, th, n_samples);
    if (g < 1) return lm_bev_workspace_bytes(p, n_points_total, LM_ALGO_BINNED, out, bytes);
    // the call runs launch sets of g samples and one last set of n_samples % g: the smaller set has fewer tiles, may
    // therefore run more bin CTAs per SM and reserve a LARGER pool -- size for both
    size_t best = 0;
    const int sets[2] = {g, n_samples % g};
    for (int k = 0; k < 2; ++k) {
        if (sets[k] < 1) continue;
        lm_bev_params ps = *p;
        ps.height = p->height * sets[k];
        Layout L;
        rc = make_layout(&ps, n_points_total, LM_ALGO_BINNED, make_kparams(&ps, th).T, &L);
        if (rc) return rc;
        if (L.total > best) best = L.total;
    }
    *bytes = best;
    return LM_OK;
}

int lm_bev_rasterize_batch(const lm_bev_params *p, int32_t n_samples, const lm_bev_sample_geom *geoms,
                           const float *const *points_dev, const int64_t *n_points, void *workspace_dev,
                           size_t workspace_bytes, const lm_bev_outputs *out, void *stream) {
    int rc = validate(p);
    if (rc) return rc;
    if (n_samples < 1 || !geoms || !points_dev || !n_points) return fail(LM_ERR_INVALID, "n_samples < 1 or a NULL table");
    if (!out || (!out->image_dev && !out->count16_dev && !out->proj_dev))
        return fail(LM_ERR_INVALID, "no output buffer requested (image, count16 or proj)");
    if (out->acc_dev) return fail(LM_ERR_UNSUPPORTED, "raw accumulators are not available from the batched call");
    if (!workspace_dev || (reinterpret_cast<uintptr_t>(workspace_dev) & 255))
        return fail(LM_ERR_WORKSPACE, "workspace_dev is NULL or not 256-byte aligned");
    int64_t total = 0;
    for (int s = 0; s < n_samples; ++s) {
        if (n_points[s] < 0 || (n_points[s] > 0 && !points_dev[s])) return fail(LM_ERR_INVALID, "sample %d: points_dev is NULL", s);
        if (reinterpret_cast<uintptr_t>(points_dev[s]) & 15) return fail(LM_ERR_INVALID, "sample %d: points_dev must be 16-byte aligned", s);
        if (n_points[s] >= (1ll << 32) - BIN_BATCH) return fail(LM_ERR_UNSUPPORTED, "sample %d: more than 2^32 points", s);
        const long long lim = 1ll << 24;
        const long long r0 = geoms[s].row0 < 0 ? -(long long)geoms[s].row0 : geoms[s].row0;
        const long long c0 = geoms[s].col0 < 0 ? -(long long)geoms[s].col0 : geoms[s].col0;
        if (r0 + p->height >= lim || c0 + p->width >= lim) return fail(LM_ERR_INVALID, "sample %d: window exceeds the exact-float index range 2^24", s);
        total += n_points[s];
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char *w = static_cast<unsigned char *>(workspace_dev);
    const int sms = sm_count();
    const bool want16 = out->count16_dev != nullptr;
    const int mask = pick_mask(needed_mask(p, want16, false), want16);
    const int th = batch_tile_h(p, mask, n_samples);
    const int group = (p->height & ((1 << th) - 1)) ? 0 : batch_group(p, th, n_samples);
    const size_t cells = (size_t)p->height * p->width;
    if (group < 1) {
        // shapes the stacked raster cannot take (height not a multiple of the tile height, or one
        // sample already needs row windows): one ordinary call per sample, same results
        for (int s = 0; s < n_samples; ++s) {
            lm_bev_params ps = *p;
            ps.bev_img_offset[0] = geoms[s].bev_img_offset[0];
            ps.bev_img_offset[1] = geoms[s].bev_img_offset[1];
            ps.local_min_ele = geoms[s].local_min_ele;
            ps.row0 = geoms[s].row0;
            ps.col0 = geoms[s].col0;
            lm_bev_outputs os = *out;
            if (os.image_dev) os.image_dev += (size_t)s * cells * p->n_channels;
            if (os.count16_dev) os.count16_dev += (size_t)s * cells;
            if (os.proj_dev) os.proj_dev += (size_t)s * cells * p->n_channels;
            rc = rasterize_impl(&ps, points_dev[s], n_points[s], LM_ALGO_BINNED, workspace_dev, workspace_bytes, &os, stream,
                                LM_STAGE_ALL, s > 0);
            if (rc) return rc;
        }
        return LM_OK;
    }
    for (int s0 = 0; s0 < n_samples; s0 += group) {
        const int nb = n_samples - s0 < group ? n_samples - s0 : group;
        lm_bev_params ps = *p;
        ps.height = p->height * nb;
        ps.row0 = 0;
        ps.col0 = 0;
        KParams kp = make_kparams(&ps, th);
        kp.bH = p->height;
        BatchTab bt;
        memset(&bt, 0, sizeof(bt));
        bt.nb = nb;
        bt.bH = p->height;
        int64_t gpts = 0;
        uint32_t batches = 0;
        for (int s = 0; s < nb; ++s) {
            bt.pts[s] = reinterpret_cast<const float4 *>(points_dev[s0 + s]);
            bt.first[s] = batches;
            bt.count[s] = (uint32_t)n_points[s0 + s];
            batches += (uint32_t)((n_points[s0 + s] + BIN_BATCH - 1) / BIN_BATCH);
            bt.off0[s] = geoms[s0 + s].bev_img_offset[0];
            bt.off1[s] = geoms[s0 + s].bev_img_offset[1];
            bt.zmin[s] = geoms[s0 + s].local_min_ele;
            bt.row0[s] = geoms[s0 + s].row0;
            bt.col0[s] = geoms[s0 + s].col0;
            gpts += n_points[s0 + s];
        }
        for (int s = nb; s <= MAX_BATCH; ++s) bt.first[s] = batches;
        Layout L;
        rc = make_layout(&ps, total, LM_ALGO_BINNED, kp.T, &L);      // sized like lm_bev_workspace_bytes_batch
        if (rc) return rc;
        if (workspace_bytes < L.total) return fail(LM_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, L.total);
        Ws ws = bind_ws(w, L);
        const size_t skip = s0 == 0 ? 0 : L.off_ctl;             // stats accumulate over the launch sets
        cudaError_t e = cudaMemsetAsync(w + skip, 0, L.zero_bytes - skip, st);
        if (e != cudaSuccess) return cuda_fail(e, "memset");
        ws.region = 1;
        ws.bin_grid = 0;
        ws.scan_in_bin = batches > 0 && gpts <= SCAN_IN_BIN_MAX_POINTS;
        if (batches > 0) {
            const size_t smem = bin_smem_bytes(kp.T);
            int grid = 0;
            rc = bin_geometry((const void *)bin_points_batch_kernel, smem, (long long)batches, kp.T, kp.tiles_x, sms, L, &ws, &grid);
            if (rc) return rc;
            bin_points_batch_kernel<<<grid, BIN_THREADS, smem, st>>>(kp, bt, ws);
        }
        if (!ws.scan_in_bin) scan_tiles_kernel<<<1, 1024, 0, st>>>(ws, kp);
        if (batches > 0) index_chunks_kernel<<<sms * 4, 256, 0, st>>>(ws);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "bin/index launch");
        Outs o = {out->image_dev, out->count16_dev, out->proj_dev, nullptr, 0};
        if (o.image) o.image += (size_t)s0 * cells * p->n_channels;
        if (o.count16) o.count16 += (size_t)s0 * cells;
        if (o.proj) o.proj += (size_t)s0 * cells * p->n_channels;
        e = launch_reduce_mask(mask, kp, ws, o, sms, st);
        if (e != cudaSuccess) return cuda_fail(e, "reduce_tiles launch");
    }
    return LM_OK;
}

// ---- plan: everything that is fixed for a stream of equally-shaped calls (include/lm_bev.h)
struct lm_bev_plan {
    lm_bev_params params;
    int64_t max_points;
    int algo;
    lm_bev_outputs out_set;          // only which pointers are non-NULL (and acc_band) matter
    lm_bev_tuning tuning;
    size_t workspace_bytes;
    // CUDA graph of the last call's launch sequence, replayed while the arguments repeat
    cudaStream_t cap_stream;
    cudaGraphExec_t exec;
    const float *k_points;
    int64_t k_n;
    void *k_ws;
    size_t k_ws_bytes;
    lm_bev_outputs k_out;
};

int lm_bev_plan_create(const lm_bev_params *p, int64_t max_points, int algo, const lm_bev_outputs *out_set,
                       const lm_bev_tuning *tuning, lm_bev_plan **plan) {
    if (!plan) return fail(LM_ERR_INVALID, "plan is NULL");
    *plan = nullptr;
    if (!out_set) return fail(LM_ERR_INVALID, "out_set is NULL (which outputs the calls will ask for fixes the tile shape)");
    TuneScope ts(tuning);
    size_t bytes = 0;
    int rc = lm_bev_workspace_bytes(p, max_points, algo, out_set, &bytes);
    if (rc) return rc;
    lm_bev_plan *pl = new (std::nothrow) lm_bev_plan();
    if (!pl) return fail(LM_ERR_INVALID, "out of host memory");
    pl->params = *p;
    pl->max_points = max_points;
    pl->algo = algo;
    pl->out_set = *out_set;
    pl->tuning = tuning ? *tuning : k_default_tuning;
    pl->workspace_bytes = bytes;
    pl->cap_stream = nullptr;
    pl->exec = nullptr;
    pl->k_points = nullptr;
    pl->k_n = -1;
    pl->k_ws = nullptr;
    pl->k_ws_bytes = 0;
    memset(&pl->k_out, 0, sizeof(pl->k_out));
    *plan = pl;
    return LM_OK;
}

int lm_bev_plan_workspace_bytes(const lm_bev_plan *plan, size_t *bytes) {
    if (!plan || !bytes) return fail(LM_ERR_INVALID, "plan/bytes is NULL");
    *bytes = plan->workspace_bytes;
    return LM_OK;
}

int lm_bev_plan_init_workspace(lm_bev_plan *plan, void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!plan) return fail(LM_ERR_INVALID, "plan is NULL");
    TuneScope ts(&plan->tuning);
    return lm_bev_workspace_init(&plan->params, plan->max_points, plan->algo, &plan->out_set, workspace_dev, workspace_bytes, stream);
}

int lm_bev_plan_rasterize(lm_bev_plan *plan, const float *points_dev, int64_t n_points, void *workspace_dev,
                          size_t workspace_bytes, const lm_bev_outputs *out, void *stream) {
    if (!plan || !out) return fail(LM_ERR_INVALID, "plan/out is NULL");
    if (n_points > plan->max_points) return fail(LM_ERR_INVALID, "n_points exceeds the plan's max_points");
    TuneScope ts(&plan->tuning);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!plan->tuning.use_graph)
        return rasterize_impl(&plan->params, points_dev, n_points, plan->algo, workspace_dev, workspace_bytes, out, stream,
                              LM_STAGE_ALL, false);
    // same arguments as the captured call: one graph launch replaces the memset + kernel launches
    const bool same = plan->exec && plan->k_points == points_dev && plan->k_n == n_points && plan->k_ws == workspace_dev &&
                      plan->k_ws_bytes == workspace_bytes && memcmp(&plan->k_out, out, sizeof(*out)) == 0;
    if (!same) {
        cudaError_t e = cudaSuccess;
        if (!plan->cap_stream && (e = cudaStreamCreateWithFlags(&plan->cap_stream, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(e, "plan stream");
        if (plan->exec) { cudaGraphExecDestroy(plan->exec); plan->exec = nullptr; }
        if ((e = cudaStreamBeginCapture(plan->cap_stream, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return cuda_fail(e, "begin capture");
        const int rc = rasterize_impl(&plan->params, points_dev, n_points, plan->algo, workspace_dev, workspace_bytes, out,
                                      plan->cap_stream, LM_STAGE_ALL, false);
        cudaGraph_t g = nullptr;
        e = cudaStreamEndCapture(plan->cap_stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return cuda_fail(e, "end capture");
        e = cudaGraphInstantiate(&plan->exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { plan->exec = nullptr; return cuda_fail(e, "graph instantiate"); }
        plan->k_points = points_dev; plan->k_n = n_points; plan->k_ws = workspace_dev; plan->k_ws_bytes = workspace_bytes;
        plan->k_out = *out;
    }
    cudaError_t e = cudaGraphLaunch(plan->exec, st);
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "graph launch");
}

int lm_bev_plan_destroy(lm_bev_plan *plan) {
    if (!plan) return LM_OK;
    if (plan->exec) cudaGraphExecDestroy(plan->exec);
    if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
    delete plan;
    return LM_OK;
}

int lm_bev_acc_merge(uint32_t *dst, int64_t dstride, const uint32_t *src, int64_t sstride, int32_t rows,
                     int32_t width, void *stream) {
    if (!dst || !src || rows < 0 || width <= 0) return fail(LM_ERR_INVALID, "bad acc_merge arguments");
    if (rows == 0) return LM_OK;
    const size_t n = (size_t)rows * width;
    acc_merge_kernel<<<sm_count() * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dst, dstride, src, sstride, n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "acc_merge launch");
}

int lm_bev_finalize(const lm_bev_params *p, const uint32_t *acc_dev, int32_t row_begin, int32_t row_end,
                    const lm_bev_outputs *out, void *stream) {
    int rc = validate(p);
    if (rc) return rc;
    if (!acc_dev || !out) return fail(LM_ERR_INVALID, "acc_dev/out is NULL");
    if (row_begin < 0 || row_end > p->height || row_begin > row_end) return fail(LM_ERR_INVALID, "bad row range");
    if (row_begin == row_end) return LM_OK;
    const KParams kp = make_kparams(p, 7);
    const Outs o = {out->image_dev, out->count16_dev, out->proj_dev, nullptr, 0};
    finalize_kernel<<<sm_count() * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(kp, acc_dev, row_begin, row_end, o);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "finalize launch");
}

int lm_bev_merge_finalize(const lm_bev_params *p, uint32_t *acc_dev, int32_t row_begin, int32_t row_end,
                          const uint32_t *recv_dev, int32_t plane_mask, const lm_bev_outputs *out, void *stream) {
    int rc = validate(p);
    if (rc) return rc;
    if (!acc_dev || !out || (plane_mask && !recv_dev)) return fail(LM_ERR_INVALID, "acc_dev/recv_dev/out is NULL");
    if (plane_mask < 0 || plane_mask >= (1 << LM_ACC_PLANES)) return fail(LM_ERR_INVALID, "plane_mask outside [0, 63]");
    if (row_begin < 0 || row_end > p->height || row_begin > row_end) return fail(LM_ERR_INVALID, "bad row range");
    if (row_begin == row_end) return LM_OK;
    const KParams kp = make_kparams(p, 7);
    const Outs o = {out->image_dev, out->count16_dev, out->proj_dev, nullptr, 0};
    merge_finalize_kernel<<<sm_count() * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(kp, acc_dev, row_begin, row_end, recv_dev, plane_mask, o);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "merge_finalize launch");
}

int lm_bev_crop_tiles(const uint8_t *image_dev, int32_t height, int32_t width, int32_t c, int32_t tile,
                      uint8_t *crops_dev, void *stream) {
    if (!image_dev || !crops_dev || height <= 0 || width <= 0 || c < 1 || c > 4 || tile <= 0)
        return fail(LM_ERR_INVALID, "bad crop_tiles arguments");
    const int ncy = (height + tile - 1) / tile, ncx = (width + tile - 1) / tile;
    const size_t row_bytes = (size_t)tile * c, pieces = (row_bytes + 15) / 16;
    const size_t total16 = (size_t)ncy * ncx * tile * pieces;
    // 128-bit path: every piece of both buffers is 16-byte aligned and whole (1152 px crops of 1..4 channels are)
    const bool vec = row_bytes % 16 == 0 && ((size_t)width * c) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(image_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(crops_dev) & 15) == 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (vec) crop_tiles_kernel<true><<<sm_count() * 8, 256, 0, st>>>(image_dev, height, width, c, tile, ncx, crops_dev, total16);
    else crop_tiles_kernel<false><<<sm_count() * 8, 256, 0, st>>>(image_dev, height, width, c, tile, ncx, crops_dev, total16);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "crop_tiles launch");
}

int lm_bev_selftest_div(float divisor, unsigned long long *out2_dev, void *stream) {
    if (!out2_dev || !(divisor >= 0x1p-20f && divisor <= 0x1p20f))
        return fail(LM_ERR_INVALID, "selftest_div: NULL output or divisor outside the fast-division range");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out2_dev, 0, 16, st);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    selftest_div_kernel<<<sm_count() * 16, 256, 0, st>>>(divisor, 1.0f / divisor, out2_dev);
    e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "selftest_div launch");
}

int lm_bev_selftest_mean(unsigned long long *out2_dev, void *stream) {
    if (!out2_dev) return fail(LM_ERR_INVALID, "selftest_mean: NULL output");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out2_dev, 0, 16, st);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    selftest_mean_kernel<<<sm_count() * 8, 256, 0, st>>>(out2_dev);
    e = cudaGetLastError();
    return e == cudaSuccess ? LM_OK : cuda_fail(e, "selftest_mean launch");
}

}  // extern "C"
