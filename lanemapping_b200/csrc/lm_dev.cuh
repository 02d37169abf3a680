// lm_dev.cuh -- device helpers shared by the kernels of liblm_bev.so (sm_100a): TMA 1-D bulk copies
// completing on an mbarrier, explicit shared-window accesses, and the LAS point-record decode.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// ------------------------------------------------------------------------------------------
// TMA 1-D bulk copy (global -> shared) completing on an mbarrier: one thread moves a whole batch
// of packed point records; no per-thread address math, no registers held while in flight
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same copy with an L2 evict-first policy: the point stream is read exactly once, marking it so
// keeps the partially written record sectors of bin_points resident in L2 until they are full
__device__ __forceinline__ void bulk_load_stream(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "LM_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra LM_DONE;\n"
                 "bra LM_WAIT;\n"
                 "LM_DONE:\n"
                 "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Shared-memory accesses of the hot loop go through explicit 32-bit shared addresses: the base is
// computed once and stays in a register (the generic form re-derives the shared window per access).
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}


// ------------------------------------------------------------------------------------------
// LAS point data records (every point format starts with X, Y, Z int32 + intensity u16, ASPRS LAS
// 1.0-1.4) -> raster-local float4.  Restated on the CPU by oracle/las_oracle.py::decode_records.
//   world = X * scale + offset            what laspy exposes as las.x / las.y / las.z (float64), the
//                                         values read_las stacks (reference
//                                         baseline/datasets/laserlane_proposals.py:618-621)
//   d     = (world - las_read_offset) - t  inverse of reference baseline/utils/coor_img2pc.py:172,175-177
//   p     = M d,  M = R(q)^T               inverse of the quaternion rotation (coor_img2pc.py:163-171)
//   out   = (float)p, (float)intensity     one rounding to binary32 each
// every step is one IEEE binary64 operation in this order (no FMA), so numpy gives the same bits.
// ------------------------------------------------------------------------------------------
struct LasXform {
    double scale[3], offset[3], read_offset[3], t[3], m[9];
    int record_length;
};

// the record starts at shared byte address `a` (any alignment); the stage has >= 6 bytes of slack
// after its last record because whole 32-bit words are read
__device__ __forceinline__ float4 las_decode_record(uint32_t a, const LasXform &x) {
    const uint32_t w = a & ~3u, sh = (a & 3u) * 8u;
    const uint32_t w0 = lds_u32(w), w1 = lds_u32(w + 4u), w2 = lds_u32(w + 8u), w3 = lds_u32(w + 12u), w4 = lds_u32(w + 16u);
    const int X = (int)__funnelshift_r(w0, w1, sh), Y = (int)__funnelshift_r(w1, w2, sh), Z = (int)__funnelshift_r(w2, w3, sh);
    const uint32_t inten = __funnelshift_r(w3, w4, sh) & 0xFFFFu;
    const double wx = __dadd_rn(__dmul_rn((double)X, x.scale[0]), x.offset[0]);
    const double wy = __dadd_rn(__dmul_rn((double)Y, x.scale[1]), x.offset[1]);
    const double wz = __dadd_rn(__dmul_rn((double)Z, x.scale[2]), x.offset[2]);
    const double d0 = __dsub_rn(__dsub_rn(wx, x.read_offset[0]), x.t[0]);
    const double d1 = __dsub_rn(__dsub_rn(wy, x.read_offset[1]), x.t[1]);
    const double d2 = __dsub_rn(__dsub_rn(wz, x.read_offset[2]), x.t[2]);
    float4 o;
    o.x = (float)__dadd_rn(__dadd_rn(__dmul_rn(x.m[0], d0), __dmul_rn(x.m[1], d1)), __dmul_rn(x.m[2], d2));
    o.y = (float)__dadd_rn(__dadd_rn(__dmul_rn(x.m[3], d0), __dmul_rn(x.m[4], d1)), __dmul_rn(x.m[5], d2));
    o.z = (float)__dadd_rn(__dadd_rn(__dmul_rn(x.m[6], d0), __dmul_rn(x.m[7], d1)), __dmul_rn(x.m[8], d2));
    o.w = (float)inten;
    return o;
}

}  // namespace
