// lm_sweep.cuh -- LM_ALGO_SWEEP's single-pass path: one persistent kernel, no record pool in HBM.
// Included by lm_bev.cu inside its anonymous namespace (uses KParams, Outs, the per-point quantisation,
// mean_small, store_pixels4 and the lm_dev.cuh helpers).
//
// Why: the two-pass pipeline writes a 4 B record per point to HBM and reads it back (24.4 B/pt of traffic
// for 16.4 B/pt of algorithmic bytes).  tools/l2ring.cu measured that a small, continuously recycled
// record ring stays resident in the 126 MB L2 while the 16 B/pt point stream flows through it (ring
// <= 20 MB, point loads marked evict-first: DRAM write-back of the ring ~0, re-reads all hit).  So:
//
//   PRODUCER CTAs (2 per SM) claim 1024-point batches in stream order, stage them with TMA bulk copies (3 stages),
//     compute the cell keys exactly as bin_points does, and route every point as ONE 32-bit record to the
//     CTA that OWNS its cell: per-owner rings in shared memory, flushed as 32-byte granules (8 records)
//     into a per-(producer, owner) mailbox ring in global memory that never leaves L2.
//   CONSUMER CTAs (1 per SM, 148 owners) each own an interleaved set of 4-column groups (rotated every 8
//     rows so that every owner sees the whole cross-track density profile) and keep a sliding window of
//     SW_R rows of their cells' accumulators in shared memory.  They poll their mailboxes, reduce the
//     records with shared-memory atomics, and emit finished rows (u8 HWC image / u16 count / f32 proj).
//   Measured (B200, config 2): exact, DRAM traffic 0.99 x the algorithmic bytes, 2.1 ms -- slower than the two-pass path
//   (0.525 ms): the consumers' 8 warps per SM are latency-bound.  Opt-in, experimental.
//
// Exact by construction, for ANY input order:
//   * a granule's words carry a phase bit (ring-wrap parity): a consumer takes a granule only when all 8
//     words show the expected phase, so data and validity travel in the same 32-bit words (no fences);
//   * before a producer sends records with rows below its previous promise it sends a MARKER m ("every
//     record that follows has row >= m"); it checks that its records stay below m + SW_SMAX (else the sweep
//     fails over, see below).  A consumer's frontier F is the minimum over its mailboxes' latest markers, its
//     window base is F - SW_M, so a record can never fall below the window, and a mailbox whose marker is too
//     far ahead of the window is simply not consumed until the window has moved (the producer with the
//     lowest marker is never held, so the system always progresses);
//   * rows below the base are final when they are emitted: every later record is >= some marker >= F.
// The sweep gives up (device flag -> the two-pass kernels, which are launched behind it and return at once
// otherwise, redo the whole raster) when a marker falls below a window that has already moved on (the cloud
// is not ordered along the rows), when a batch spans more than SW_SMAX rows, or when a packed 12-bit cell
// count wraps (conservation check).  A failed sweep disables itself for the next calls on that workspace.
#pragma once

constexpr int SW_OWNERS = 148;                     // consumer CTAs = cell owners (one per B200 SM)
constexpr int SW_PRODUCERS = 2 * SW_OWNERS;        // producer CTAs
constexpr int SW_CTAS_PER_SM = 3;                  // 1 consumer + 2 producers: 3 x 70 KB of shared memory
constexpr int SW_STAGES = 3;                       // TMA stage buffers of a producer
constexpr int SW_GRID = SW_OWNERS + SW_PRODUCERS;
constexpr int SW_THREADS = 256;
constexpr int SW_PPT = 4;
constexpr int SW_BATCH = SW_THREADS * SW_PPT;      // points per producer batch
constexpr int SW_RS = 32;                          // slots of a per-owner ring in producer shared memory
constexpr int SW_CAP_LOG2 = 3;
constexpr int SW_CAP = 1 << SW_CAP_LOG2;           // granules per (producer, owner) mailbox ring
constexpr int SW_R_LOG2 = 10;
constexpr int SW_R = 1 << SW_R_LOG2;               // rows of the consumer's sliding window
constexpr int SW_M = 16;                           // rows kept open below the frontier
constexpr int SW_ADV = 96;                         // a producer renews its marker when its batches moved this far
constexpr int SW_SMAX = 512;                       // records stay below marker + SW_SMAX
constexpr int SW_TAG_BITS = 11;                    // a record carries row mod 2^11: with base <= row < marker + SW_SMAX <= base + 2^11
constexpr int SW_TAG_SPAN = 1 << SW_TAG_BITS;      //   (the marker gate) the consumer recovers row - base exactly
constexpr int SW_STRIDE = 41;                      // ownership rotation per row block (coprime with 148)
constexpr int SW_RB_LOG2 = 3;                      // rows per ownership block
constexpr int SW_MAX_LG = 2;                       // 4-column groups per owner and row (W <= 4 * 148 * 2)
constexpr int SW_CPO = 4 * SW_MAX_LG;              // columns per owner and row
constexpr int SW_MBPT = (SW_PRODUCERS + SW_THREADS - 1) / SW_THREADS;   // mailboxes per consumer thread
constexpr uint32_t SW_PHASE = 1u << 31, SW_VALID = 1u << 30, SW_MARKER = 1u << 29, SW_DONE = 1u << 28;
constexpr int SW_INF_ROW = 0x7fffffff;
constexpr int SW_COOLDOWN = 16;                    // calls a failed sweep stays off on its workspace
constexpr int SW_FRONTIER_EVERY = 1;               // consumer loops between two frontier updates
constexpr int SW_PAD_ROUND = 64;                   // retry rounds (~20 us) after which a waiting producer pads its rings out
static_assert(SW_RS % 8 == 0 && SW_RS >= 16, "ring = whole granules");
static_assert(SW_SMAX + SW_M + 8 <= SW_TAG_SPAN && SW_R <= SW_TAG_SPAN, "the gate must leave room for the lowest marker");
static_assert(SW_MAX_LG * 4 <= 8, "3 bits of owner-local column");

struct SweepPersist {            // survives between calls (lm_bev_workspace_init writes it)
    uint32_t magic;
    uint32_t cooldown;           // > 0: the sweep is skipped (a recent call failed over)
    uint32_t n_failed, n_ok;
    unsigned long long dbg[12];  // -DLM_SWEEP_DEBUG builds: event counters of the last call (tools/debug_sweep.py)
};
#ifdef LM_SWEEP_DEBUG
#define SW_DBG(i, v) atomicAdd(&sw.persist->dbg[i], (unsigned long long)(v))
#else
#define SW_DBG(i, v) ((void)0)
#endif
struct SweepWs {
    SweepPersist *persist;
    uint32_t *heads;             // [producer][owner] granules consumed so far (absolute, persists between calls)
    uint4 *mail;                 // [producer][owner][SW_CAP] granules of 8 words (two uint4)
    unsigned int *fail;          // per-call flag in Ctl: the two-pass kernels behind the sweep run iff != 0
    unsigned int *next_batch;    // per-call batch counter in Ctl
    lm_bev_stats *stats;
    uint32_t magic;
};

constexpr size_t SW_PROD_SMEM = SW_STAGES * (size_t)SW_BATCH * 16 + (size_t)SW_OWNERS * SW_RS * 4 + 3 * (size_t)SW_OWNERS * 4;
constexpr size_t SW_CONS_SMEM = 2 * (size_t)SW_R * SW_CPO * 4;
constexpr size_t SW_SMEM = SW_PROD_SMEM > SW_CONS_SMEM ? SW_PROD_SMEM : SW_CONS_SMEM;

// Mailbox words, heads and the fail flag are read at the L2 (ld.global.cg: never from a stale L1 line) with WEAK
// loads: the protocol needs no ordering between them (phase bits / monotonic counters), and strong (volatile)
// accesses of one thread are performed one after the other -- four L2 round trips in series per consumer loop
// (measured: 7.4 us per loop with ld.volatile).
__device__ __forceinline__ uint32_t ld_vol_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_vol_u4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void reds_add(uint32_t a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void reds_max(uint32_t a, uint32_t v) {
    asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// predicated forms: ONE instruction under a predicate, never a branch around it
__device__ __forceinline__ void reds_add_if(bool p, uint32_t a, uint32_t v) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q red.shared.add.u32 [%1], %2;\n}" ::"r"((uint32_t)p), "r"(a), "r"(v));
}
__device__ __forceinline__ void reds_max_if(bool p, uint32_t a, uint32_t v) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q red.shared.max.u32 [%1], %2;\n}" ::"r"((uint32_t)p), "r"(a), "r"(v));
}

__device__ __forceinline__ unsigned long long sweep_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Watchdog: nothing in a sweep waits longer than it takes to run; a wait of seconds means a protocol error or a
// CTA that never became resident.  fail = 2 makes every loop of every CTA leave, the two-pass kernels redo the
// raster, and the epilogue invalidates the mailboxes (the sweep stays off until lm_bev_workspace_init).
constexpr unsigned long long SW_PANIC_NS = 4000000000ull;
__device__ __forceinline__ bool sweep_panic(const SweepWs &sw, unsigned long long t_start) {
    if (ld_vol_u32((const uint32_t *)sw.fail) == 2u) return true;
    if (sweep_now_ns() - t_start > SW_PANIC_NS) { atomicExch(sw.fail, 2u); return true; }
    return false;
}

// owner of the 4-column group g in row r, and the group's index among the owner's groups of that row
__device__ __forceinline__ void sweep_owner(int r, int c, uint32_t &owner, uint32_t &lcol) {
    const uint32_t g = (uint32_t)c >> 2;
    const uint32_t lg = g / (uint32_t)SW_OWNERS;
    owner = (g + (uint32_t)SW_STRIDE * ((uint32_t)r >> SW_RB_LOG2)) % (uint32_t)SW_OWNERS;
    lcol = (lg << 2) | ((uint32_t)c & 3u);
}
// first group (lg = 0) that `owner` holds in row r: the inverse of sweep_owner
__device__ __forceinline__ uint32_t sweep_group0(int r, uint32_t owner) {
    const uint32_t s = ((uint32_t)SW_STRIDE * ((uint32_t)r >> SW_RB_LOG2)) % (uint32_t)SW_OWNERS;
    return (owner + (uint32_t)SW_OWNERS - s) % (uint32_t)SW_OWNERS;
}

// ------------------------------------------------------------------------------------------
// producer
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sweep_producer(const KParams &kp, const float4 *__restrict__ pts, long long n, const SweepWs &sw,
                                               const uint32_t pid, unsigned char *smem_raw, uint64_t *s_bar, int *s_misc) {
    const int tid = threadIdx.x;
    const uint32_t sm_stage = smem_u32(smem_raw);
    const uint32_t sm_ring = sm_stage + (uint32_t)SW_STAGES * SW_BATCH * 16u;
    const uint32_t sm_pos = sm_ring + (uint32_t)SW_OWNERS * SW_RS * 4u;     // records appended per owner (absolute)
    const uint32_t sm_flushed = sm_pos + (uint32_t)SW_OWNERS * 4u;          // records flushed per owner (multiple of 8)
    const uint32_t sm_headc = sm_flushed + (uint32_t)SW_OWNERS * 4u;        // cached consumer head (granules)
    // s_misc: [2] batch min row, [3] batch max row, [4] abort, [5 .. 5+SW_STAGES) claimed batches (ring)
    const long long nb = (n + SW_BATCH - 1) / SW_BATCH;
    uint32_t *my_heads = sw.heads + (size_t)pid * SW_OWNERS;
    uint4 *my_mail = sw.mail + (size_t)pid * SW_OWNERS * SW_CAP * 2;

    if (tid < SW_OWNERS) {
        const uint32_t h = ld_vol_u32(my_heads + tid);      // everything sent by earlier calls was consumed
        sts_u32(sm_pos + 4u * tid, h << 3);
        sts_u32(sm_flushed + 4u * tid, h << 3);
        sts_u32(sm_headc + 4u * tid, h);
    }
    auto claim_and_load = [&](int slot, bool abort) {           // thread 0: next batch of the stream -> stage `slot`
        const long long b = abort ? nb : (long long)atomicAdd(sw.next_batch, 1u);
        s_misc[5 + slot] = b < nb ? (int)b : -1;
        if (b < nb) {
            const long long left = n - b * SW_BATCH;
            bulk_load_stream(smem_raw + (uint32_t)slot * (SW_BATCH * 16u), pts + b * SW_BATCH,
                             (uint32_t)(left < SW_BATCH ? left : SW_BATCH) * 16u, &s_bar[slot]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < SW_STAGES; ++b) mbar_init(&s_bar[b], 1);
        s_misc[2] = SW_INF_ROW;
        s_misc[3] = -1;
        s_misc[4] = 0;
#pragma unroll
        for (int b = 0; b < SW_STAGES - 1; ++b) claim_and_load(b, false);     // SW_STAGES - 1 batches fly ahead
    }
    __syncthreads();

    // whole granules (8 records) of owner `o`'s ring go to its mailbox.  Never waits: returns false when the
    // consumer is SW_CAP granules behind (the caller retries; a producer that has to wait first pushes out every
    // partial granule it holds, so that none of its markers can keep another owner's window from moving)
    auto flush_ring = [&](uint32_t o) -> bool {
        uint32_t f = lds_u32(sm_flushed + 4u * o);
        const uint32_t p = lds_u32(sm_pos + 4u * o);
        uint32_t t = p & ~7u;
        if (t - f > (uint32_t)SW_RS) t = f + SW_RS;        // positions beyond f + RS are not written yet
        if (t == f) return true;
        uint32_t hc = lds_u32(sm_headc + 4u * o);
        bool room = true;
        while (f != t) {
            const uint32_t gi = f >> 3;
            if (gi - hc >= (uint32_t)SW_CAP) {
                hc = ld_vol_u32(my_heads + o);
                if (gi - hc >= (uint32_t)SW_CAP) { room = false; break; }
            }
            const uint32_t a = sm_ring + (o * SW_RS + (f & (SW_RS - 1))) * 4u;
            const uint4 lo = lds_u4(a), hi = lds_u4(a + 16u);
            uint4 *dst = my_mail + ((size_t)o * SW_CAP + (gi & (SW_CAP - 1))) * 2;
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 1), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
            f += 8u;
        }
        sts_u32(sm_flushed + 4u * o, f);
        sts_u32(sm_headc + 4u * o, hc);
        return room;
    };
    auto phase_of = [](uint32_t ps) -> uint32_t { return ((ps >> (3 + SW_CAP_LOG2)) & 1u) << 31; };
    // fill owner `o`'s ring up to a whole granule with pad words (called by thread o between barriers, when
    // nobody appends): what is in the ring can then leave with the next flush
    auto pad_ring = [&](uint32_t o) {
        uint32_t ps = lds_u32(sm_pos + 4u * o);
        const uint32_t f = lds_u32(sm_flushed + 4u * o);
        if ((ps & 7u) == 0u || ((ps + 7u) & ~7u) - f > (uint32_t)SW_RS) return;   // nothing partial, or no room (it is full of records)
        while (ps & 7u) {
            sts_u32(sm_ring + (o * SW_RS + (ps & (SW_RS - 1))) * 4u, phase_of(ps));      // neither valid nor marker
            ++ps;
        }
        sts_u32(sm_pos + 4u * o, ps);
    };
    const unsigned long long t_start = sweep_now_ns();

    int m_last = 0;
    bool have_marker = false, panic = false;
    [[maybe_unused]] unsigned long long d_rounds = 0, d_markers = 0, d_batches = 0, d_wait = 0, d_tma = 0;
    for (int k = 0;; ++k) {
        const uint32_t buf = (uint32_t)k % (uint32_t)SW_STAGES;
        const int cur = s_misc[5 + buf];
        if (cur < 0) break;                                  // block-uniform
        if (tid == 0)   // claim one more batch and start its copy into the buffer consumed one batch ago
            claim_and_load((int)((uint32_t)(k + SW_STAGES - 1) % (uint32_t)SW_STAGES),
                           s_misc[4] != 0 || ld_vol_u32((const uint32_t *)sw.fail) != 0u);
        const long long left = n - (long long)cur * SW_BATCH;
        const uint32_t npts = left < SW_BATCH ? (uint32_t)left : (uint32_t)SW_BATCH;
#ifdef LM_SWEEP_DEBUG
        const unsigned long long t_w0 = sweep_now_ns();
#endif
        mbar_wait(&s_bar[buf], ((uint32_t)k / (uint32_t)SW_STAGES) & 1u);
#ifdef LM_SWEEP_DEBUG
        d_tma += sweep_now_ns() - t_w0;
        ++d_batches;
#endif

        float4 p[SW_PPT];
        const uint32_t my_stage = sm_stage + buf * (SW_BATCH * 16u) + (uint32_t)tid * 16u;
#pragma unroll
        for (int j = 0; j < SW_PPT; ++j) p[j] = lds_f4(my_stage + (uint32_t)j * (SW_THREADS * 16u));
        if (npts < (uint32_t)SW_BATCH) {
#pragma unroll
            for (int j = 0; j < SW_PPT; ++j)
                if ((uint32_t)(j * SW_THREADS + tid) >= npts) p[j].x = __int_as_float(0x7fc00000);   // NaN x: dropped
        }
        // ---- keys: the same exact arithmetic as bin_points (packed FP32x2 divisions)
        int r[SW_PPT], c[SW_PPT];
        uint32_t iq[SW_PPT], zq[SW_PPT];
        bool ok[SW_PPT];
        {
            const Geo geo = geo_of(kp);
            float2 qxy[SW_PPT];
            float qz[SW_PPT];
            float lo = 0x1p100f, hi = 0.0f;
            const float2 noff = make_float2(-geo.off0, -geo.off1), reso2 = make_float2(kp.reso0, kp.reso1),
                         rr2 = make_float2(kp.rreso0, kp.rreso1);
            const float2 nzmin2 = make_float2(-geo.zmin, -geo.zmin), zreso2 = make_float2(kp.zreso, kp.zreso),
                         rz2 = make_float2(kp.rzreso, kp.rzreso);
#pragma unroll
            for (int j = 0; j < SW_PPT; ++j) {
                const float2 d = __fadd2_rn(make_float2(p[j].x, p[j].y), noff);
                lo = fminf(lo, fminf(fabsf(d.x), fabsf(d.y)));
                hi = fmaxf(hi, fmaxf(fabsf(d.x), fabsf(d.y)));
                qxy[j] = div_const2(d, reso2, rr2);
            }
#pragma unroll
            for (int j = 0; j < SW_PPT; j += 2) {
                const float2 d = __fadd2_rn(make_float2(p[j].z, p[j + 1].z), nzmin2);
                lo = fminf(lo, fminf(fabsf(d.x), fabsf(d.y)));
                hi = fmaxf(hi, fmaxf(fabsf(d.x), fabsf(d.y)));
                const float2 q = div_const2(d, zreso2, rz2);
                qz[j] = q.x;
                qz[j + 1] = q.y;
            }
            if (kp.fast_div && fast_range_ok(lo, hi)) {
#pragma unroll
                for (int j = 0; j < SW_PPT; ++j)
                    ok[j] = keys_from_quotients(qxy[j].x, qxy[j].y, qz[j], p[j].w, kp, geo, r[j], c[j], iq[j], zq[j]);
            } else {
#pragma unroll
                for (int j = 0; j < SW_PPT; ++j) ok[j] = quantise_ieee(p[j], kp, geo, r[j], c[j], iq[j], zq[j]);
            }
        }
        // ---- the batch's row range (decides the marker; checked against SW_SMAX)
        {
            int mn = SW_INF_ROW, mx = -1;
#pragma unroll
            for (int j = 0; j < SW_PPT; ++j)
                if (ok[j]) { mn = min(mn, r[j]); mx = max(mx, r[j]); }
            mn = __reduce_min_sync(0xffffffffu, mn);
            mx = __reduce_max_sync(0xffffffffu, mx);
            if ((tid & 31) == 0 && mx >= 0) { atomicMin(&s_misc[2], mn); atomicMax(&s_misc[3], mx); }
        }
        __syncthreads();                                               // barrier 1
        const int bmn = s_misc[2], bmx = s_misc[3];
        const bool any = bmx >= 0;
        // a new promise (the first one, a lower one, or one raised by SW_ADV rows) sits in front of this batch's
        // records in every ring (one more barrier): both of its bounds, m <= row < m + SW_SMAX, then hold for
        // everything the consumer reads behind it
        const bool send = any && (!have_marker || bmn < m_last || bmn >= m_last + SW_ADV);
        const int m_ref = send ? bmn : m_last;
        bool skip = false;
        if (any && bmx - m_ref >= SW_SMAX) {                           // the batch is not a narrow band of rows: give up
            skip = true;
            if (tid == 0) { atomicExch(sw.fail, 1u); s_misc[4] = 1; }
        }
        if (send && !skip) { m_last = bmn; have_marker = true; ++d_markers; }

        // ---- items of this thread: up to 4 records + (threads < 148) one marker, appended to the owners' rings
        uint32_t word[SW_PPT + 1], own[SW_PPT + 1], ps[SW_PPT + 1];
        uint32_t pend = 0;
        if (send && !skip && tid < SW_OWNERS) {
            own[SW_PPT] = (uint32_t)tid;
            word[SW_PPT] = SW_MARKER | ((uint32_t)bmn & 0xFFFFFFu);
            ps[SW_PPT] = atoms_add(sm_pos + 4u * tid, 1u);
            pend |= 1u << SW_PPT;
        }
        if (send && !skip) __syncthreads();                           // block-uniform
#pragma unroll
        for (int j = 0; j < SW_PPT; ++j) {
            if (ok[j] && !skip) {
                uint32_t lcol;
                sweep_owner(r[j], c[j], own[j], lcol);
                word[j] = SW_VALID | (lcol << 27) | (((uint32_t)r[j] & (SW_TAG_SPAN - 1)) << 16) | (iq[j] << 8) | zq[j];
                ps[j] = atoms_add(sm_pos + 4u * own[j], 1u);
                pend |= 1u << j;
            }
        }
        for (int round = 0;; ++round) {
#pragma unroll
            for (int j = 0; j <= SW_PPT; ++j) {
                if (pend & (1u << j)) {
                    const uint32_t fl = lds_u32(sm_flushed + 4u * own[j]);
                    if (ps[j] - fl < (uint32_t)SW_RS) {
                        sts_u32(sm_ring + (own[j] * SW_RS + (ps[j] & (SW_RS - 1))) * 4u, word[j] | phase_of(ps[j]));
                        pend &= ~(1u << j);
                    }
                }
            }
            __syncthreads();                                           // barrier 2: ring writes visible
            bool more = pend != 0u;
            if (tid < SW_OWNERS) {
                if (round == SW_PAD_ROUND && pend == 0u) pad_ring((uint32_t)tid);   // a long wait: let every marker this CTA holds go
                more |= !flush_ring((uint32_t)tid);
            }
            if (tid == SW_THREADS - 1) { s_misc[2] = SW_INF_ROW; s_misc[3] = -1; }
            if (!__syncthreads_or(more)) break;                        // barrier 3 (more rounds: a full ring or a full mailbox)
            ++d_rounds;
#ifdef LM_SWEEP_DEBUG
            const unsigned long long t_s0 = sweep_now_ns();
#endif
            {
                __nanosleep(round < 4 ? 300 : 1000);
                if ((round & 255) == 0 && __syncthreads_or(sweep_panic(sw, t_start))) { panic = true; break; }
            }
#ifdef LM_SWEEP_DEBUG
            d_wait += sweep_now_ns() - t_s0;
#endif
        }
        if (panic) break;
    }
    if (tid == 0) { SW_DBG(0, d_batches); SW_DBG(1, d_rounds); SW_DBG(2, d_markers); SW_DBG(3, d_wait); SW_DBG(4, d_tma);
                    SW_DBG(5, sweep_now_ns() - t_start); }
    // ---- end of stream: DONE marker, pad every ring to a whole granule, flush
    if (tid < SW_OWNERS) {
        const uint32_t o = (uint32_t)tid;
        uint32_t ps = lds_u32(sm_pos + 4u * o);                       // nobody else appends any more
        sts_u32(sm_ring + (o * SW_RS + (ps & (SW_RS - 1))) * 4u, SW_MARKER | SW_DONE | 0xFFFFFFu | phase_of(ps));
        ++ps;
        while (ps & 7u) {
            sts_u32(sm_ring + (o * SW_RS + (ps & (SW_RS - 1))) * 4u, phase_of(ps));     // pad: neither valid nor marker
            ++ps;
        }
        sts_u32(sm_pos + 4u * o, ps);
        for (int round = 0; !flush_ring(o); ++round) {
            __nanosleep(200);
            if ((round & 255) == 255 && sweep_panic(sw, t_start)) break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// consumer
// ------------------------------------------------------------------------------------------
template <int MASK>
__device__ __forceinline__ void sweep_consumer(const KParams &kp, const SweepWs &sw, const Outs &out, const uint32_t oid,
                                               unsigned char *smem_raw, int *s_misc, float *s_div255) {
    constexpr bool HAS_CNT = (MASK & M_CNT) != 0;
    constexpr bool HAS_SUM = (MASK & M_SUMZ) != 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sm_w0 = smem_u32(smem_raw);                         // [SW_R][SW_CPO] packed count:12 | sum_z:20
    const uint32_t sm_w1 = sm_w0 + (uint32_t)SW_R * SW_CPO * 4u;       // [SW_R][SW_CPO] max intensity
    // s_misc: [0..7] per-warp frontier minima, [8] mailboxes done, [10] failed, [12],[13] markers seen (by loop parity)
    for (uint32_t i = (uint32_t)tid * 16u; i < 2u * SW_R * SW_CPO * 4u; i += SW_THREADS * 16u) sts_u4(sm_w0 + i, make_uint4(0, 0, 0, 0));
    if (out.proj) for (int i = tid; i < 256; i += SW_THREADS) s_div255[i] = __fdiv_rn((float)i, 255.0f);
    if (tid == 0) { s_misc[8] = 0; s_misc[10] = 0; s_misc[12] = 0; s_misc[13] = 0; }

    uint32_t head[SW_MBPT], hpub[SW_MBPT], sub[SW_MBPT];
    int mark[SW_MBPT];
    bool done[SW_MBPT];
#pragma unroll
    for (int i = 0; i < SW_MBPT; ++i) {
        const uint32_t p = (uint32_t)tid + (uint32_t)i * SW_THREADS;
        done[i] = p >= (uint32_t)SW_PRODUCERS;
        head[i] = done[i] ? 0u : ld_vol_u32(sw.heads + (size_t)p * SW_OWNERS + oid);
        hpub[i] = head[i];
        sub[i] = 0;
        mark[i] = done[i] ? SW_INF_ROW : 0;                            // no marker yet: the frontier stays at row 0
    }
    __syncthreads();

    const int H = kp.H, W = kp.W, nch = kp.nch, ch0 = kp.ch[0], ch1 = kp.ch[1], ch2 = kp.ch[2], ch3 = kp.ch[3];
    const int groups = (W + 3) >> 2;
    const size_t gcells = (size_t)kp.oH * W;
    const size_t row_bytes = (size_t)W * nch;
    const bool img_fast = (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(out.image) & 3) == 0;
    const bool c16_fast = (W & 1) == 0 && (reinterpret_cast<uintptr_t>(out.count16) & 3) == 0;
    const bool proj_fast = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(out.proj) & 15) == 0;
    unsigned long long n_valid = 0, n_counted = 0;
    int base = 0;
    int n_done_local = 0;
#pragma unroll
    for (int i = 0; i < SW_MBPT; ++i) n_done_local += done[i] ? 1 : 0;
    int done_reported = 0;

    // finished rows [r0, r1): derive the channels, write them, zero the window slots
    auto emit_rows = [&](int r0, int r1) {
        const int items = (r1 - r0) * SW_MAX_LG;
        for (int it = tid; it < items; it += SW_THREADS) {
            const int row = r0 + it / SW_MAX_LG, lg = it % SW_MAX_LG;
            const int g = (int)sweep_group0(row, oid) + lg * SW_OWNERS;
            if (g >= groups) continue;                                 // this owner holds fewer groups in this row block
            const uint32_t slot = (((uint32_t)row & (SW_R - 1)) * SW_CPO + (uint32_t)lg * 4u) * 4u;
            const uint4 z4 = make_uint4(0, 0, 0, 0);
            uint4 v0 = z4, v1 = lds_u4(sm_w1 + slot);
            sts_u4(sm_w1 + slot, z4);
            if (HAS_CNT) { v0 = lds_u4(sm_w0 + slot); sts_u4(sm_w0 + slot, z4); }
            const uint32_t a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w};
            const int c0 = g << 2, ncols = min(4, W - c0);
            uint32_t pk[4], k16[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t cnt = a0[e] >> PK_SHIFT, sz = a0[e] & PK_SUM_MASK, mi = a1[e];
                n_counted += cnt;
                const uint32_t mean_z = (HAS_SUM && cnt) ? mean_small(sz, cnt) : 0u;
                const uint32_t dens = cnt < 255u ? cnt : 255u;
                auto pick = [&](int ch) -> uint32_t { return ch == LM_CH_MAX_I ? mi : ch == LM_CH_MEAN_Z ? mean_z : ch == LM_CH_DENSITY ? dens : 0u; };
                uint32_t v = pick(ch0);
                if (nch > 1) v |= pick(ch1) << 8;
                if (nch > 2) v |= pick(ch2) << 16;
                if (nch > 3) v |= pick(ch3) << 24;
                pk[e] = v;
                k16[e] = cnt;                                         // <= 4095
            }
            const size_t cell = (size_t)(kp.orow + row) * W + c0;
            if (out.image) {
                uint8_t *dst = out.image + cell * nch;
                if (img_fast && ncols == 4) {
                    switch (nch) {
                        case 1: store_pixels4<1>(dst, pk); break;
                        case 2: store_pixels4<2>(dst, pk); break;
                        case 3: store_pixels4<3>(dst, pk); break;
                        default: store_pixels4<4>(dst, pk); break;
                    }
                } else {
                    for (int e = 0; e < ncols; ++e)
                        for (int cc = 0; cc < nch; ++cc) dst[e * nch + cc] = (uint8_t)(pk[e] >> (8 * cc));
                }
            }
            if (out.count16) {
                uint16_t *dst = out.count16 + cell;
                if (c16_fast && ncols == 4) {
                    reinterpret_cast<uint32_t *>(dst)[0] = k16[0] | (k16[1] << 16);
                    reinterpret_cast<uint32_t *>(dst)[1] = k16[2] | (k16[3] << 16);
                } else {
                    for (int e = 0; e < ncols; ++e) dst[e] = (uint16_t)k16[e];
                }
            }
            if (out.proj) {
                for (int cc = 0; cc < nch; ++cc) {
                    float *dst = out.proj + (size_t)cc * gcells + cell;
                    if (proj_fast && ncols == 4) {
                        *reinterpret_cast<float4 *>(dst) =
                            make_float4(s_div255[(pk[0] >> (8 * cc)) & 0xFFu], s_div255[(pk[1] >> (8 * cc)) & 0xFFu],
                                        s_div255[(pk[2] >> (8 * cc)) & 0xFFu], s_div255[(pk[3] >> (8 * cc)) & 0xFFu]);
                    } else {
                        for (int e = 0; e < ncols; ++e) dst[e] = s_div255[(pk[e] >> (8 * cc)) & 0xFFu];
                    }
                }
            }
        }
    };

    const unsigned long long t_start = sweep_now_ns();
    [[maybe_unused]] unsigned long long d_polls = 0, d_hits = 0, d_gated = 0, d_emit = 0;
    // the granule at every mailbox's head is fetched one loop ahead (registers): the L2 round trip of mailbox i
    // runs under the reduction of the other mailboxes' granules and the frontier bookkeeping
    uint4 glo[SW_MBPT], ghi[SW_MBPT];
#pragma unroll
    for (int i = 0; i < SW_MBPT; ++i) {
        glo[i] = ghi[i] = make_uint4(0, 0, 0, 0);
        if (!done[i]) {
            const uint4 *src = sw.mail + (((size_t)tid + (size_t)i * SW_THREADS) * SW_OWNERS + oid) * SW_CAP * 2 + (size_t)(head[i] & (SW_CAP - 1)) * 2;
            glo[i] = ld_vol_u4(src);
            ghi[i] = ld_vol_u4(src + 1);
        }
    }
    for (int it = 0;; ++it) {
        bool saw = false;
#pragma unroll
        for (int i = 0; i < SW_MBPT; ++i) {
            if (done[i]) continue;
            const uint32_t p = (uint32_t)tid + (uint32_t)i * SW_THREADS;
            const uint4 *mb = sw.mail + ((size_t)p * SW_OWNERS + oid) * SW_CAP * 2;
            const uint4 lo = glo[i], hi = ghi[i];
            const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            const uint32_t exp = ((head[i] >> SW_CAP_LOG2) & 1u) << 31;
            uint32_t diff = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) diff |= (w[q] ^ exp);
            ++d_polls;
            // Words are taken in order; a record that does not fit the window under the promise in force stops the
            // granule there (resumed at `sub` in a later loop).  Markers in front of it are applied, so the mailbox
            // with the lowest promise always moves on.  Straight-line code: every lane of the warp has its own
            // granule with its own mix of records, markers and pads, so everything is selects and predicated
            // atomics (a branchy version ran 60 instructions per word: ncu, profiles/r02_sweep_v2).
            const bool ready = !(diff & SW_PHASE);                    // all 8 words have arrived
            {
                bool stop = !ready;
                uint32_t nsub = sub[i], taken = 0;
                int cm = mark[i];
                bool bad = false, fin = false;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t x = w[q];
                    const bool act = !stop && (uint32_t)q >= sub[i];
                    const bool is_mk = act && (x & (SW_VALID | SW_MARKER)) == SW_MARKER;
                    const bool is_fin = is_mk && (x & SW_DONE) != 0u;
                    const int v = (int)(x & 0xFFFFFFu);
                    bad |= is_mk && !is_fin && v < base;               // rows that were already emitted: the cloud is not row-ordered
                    cm = is_fin ? SW_INF_ROW : (is_mk ? v : cm);
                    fin |= is_fin;
                    saw |= is_mk;
                    const bool rec = act && (x & SW_VALID) != 0u;
                    // row - base from the 11-bit tag: exact while the promise in force keeps the row below
                    // base + 2^11 (first test); the record waits for the frontier while it is beyond the window
                    const uint32_t ahead = ((x >> 16) - (uint32_t)base) & (SW_TAG_SPAN - 1);
                    const bool held = rec && (cm + SW_SMAX > base + SW_TAG_SPAN || ahead >= (uint32_t)SW_R);
                    nsub = held ? (uint32_t)q : nsub;
                    stop |= held;
                    const bool take = rec && !held;
                    taken += take ? 1u : 0u;
                    const uint32_t slot = ((((x >> 16) & (SW_R - 1)) * SW_CPO) + ((x >> 27) & 7u)) * 4u;   // always inside the window
                    const uint32_t iqv = (x >> 8) & 0xFFu;
                    if (HAS_CNT) reds_add_if(take, sm_w0 + slot, (1u << PK_SHIFT) | (HAS_SUM ? (x & 0xFFu) : 0u));
                    reds_max_if(take, sm_w1 + slot, iqv);
                }
                n_valid += taken;
                mark[i] = cm;
                if (ready) {
                    ++d_hits;
                    if (bad) { s_misc[10] = 1; atomicExch(sw.fail, 1u); }      // producers stop claiming batches
                    if (fin) { done[i] = true; ++n_done_local; }
                    if (stop) { sub[i] = nsub; ++d_gated; }
                    else { sub[i] = 0; ++head[i]; }
                }
            }
            if (!done[i]) {                                           // next loop's granule of this mailbox
                const uint4 *src = mb + (size_t)(head[i] & (SW_CAP - 1)) * 2;
                glo[i] = ld_vol_u4(src);
                ghi[i] = ld_vol_u4(src + 1);
            }
            if (head[i] - hpub[i] >= 2u || (done[i] && head[i] != hpub[i])) {
                st_vol_u32(sw.heads + (size_t)p * SW_OWNERS + oid, head[i]);
                hpub[i] = head[i];
            }
        }
        // ---- every SW_FRONTIER_EVERY loops: frontier = min over the mailboxes' latest markers; emit the rows below it
        if (saw) s_misc[12] = 1;
        if (n_done_local != done_reported) { atomicAdd(&s_misc[8], n_done_local - done_reported); done_reported = n_done_local; }
        if ((it % SW_FRONTIER_EVERY) != SW_FRONTIER_EVERY - 1) continue;
        int mn = SW_INF_ROW;
#pragma unroll
        for (int i = 0; i < SW_MBPT; ++i) mn = min(mn, mark[i]);
        mn = __reduce_min_sync(0xffffffffu, mn);
        if (lane == 0) s_misc[warp] = mn;
        __syncthreads();
        const bool any_marker = s_misc[12] != 0;
        const bool all_done = s_misc[8] >= SW_THREADS * SW_MBPT;
        int F = SW_INF_ROW;
#pragma unroll
        for (int q = 0; q < SW_THREADS / 32; ++q) F = min(F, s_misc[q]);
        __syncthreads();
        if (tid == 0) s_misc[12] = 0;            // (a flag raised before this store by a thread already in the next loop is
                                                 //  only lost until the next marker: emission is delayed, never wrong)
        if (any_marker || all_done) {
            int nbase = all_done ? H : min(F - SW_M, H);
            if (nbase > base) {
#ifdef LM_SWEEP_DEBUG
                const unsigned long long t_e0 = sweep_now_ns();
#endif
                emit_rows(base, nbase);
                base = nbase;
                __syncthreads();
#ifdef LM_SWEEP_DEBUG
                d_emit += sweep_now_ns() - t_e0;
#endif
            }
        }
        if (all_done) break;
        if ((it & 1023) == 1023 && __syncthreads_or(sweep_panic(sw, t_start))) break;
    }
#ifdef LM_SWEEP_DEBUG
    for (int o = 16; o; o >>= 1) {
        d_polls += __shfl_xor_sync(0xffffffffu, d_polls, o);
        d_hits += __shfl_xor_sync(0xffffffffu, d_hits, o);
        d_gated += __shfl_xor_sync(0xffffffffu, d_gated, o);
    }
    if (lane == 0) { SW_DBG(6, d_polls); SW_DBG(7, d_hits); SW_DBG(8, d_gated); }
    if (tid == 0) { SW_DBG(9, d_emit); SW_DBG(10, sweep_now_ns() - t_start); }
#endif
    // ---- hand the mailboxes over to the next call, report
#pragma unroll
    for (int i = 0; i < SW_MBPT; ++i) {
        const uint32_t p = (uint32_t)tid + (uint32_t)i * SW_THREADS;
        if (p < (uint32_t)SW_PRODUCERS && head[i] != hpub[i]) st_vol_u32(sw.heads + (size_t)p * SW_OWNERS + oid, head[i]);
    }
    for (int o = 16; o; o >>= 1) {
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
        n_counted += __shfl_xor_sync(0xffffffffu, n_counted, o);
    }
    __shared__ unsigned long long s_tot[2];
    if (tid == 0) { s_tot[0] = 0; s_tot[1] = 0; }
    __syncthreads();
    if (lane == 0) { atomicAdd(&s_tot[0], n_valid); atomicAdd(&s_tot[1], n_counted); }
    __syncthreads();
    if (tid == 0) {
        if (s_tot[0]) atomicAdd((unsigned long long *)&sw.stats->n_valid, s_tot[0]);
        // a wrapped 12-bit count loses 4096: the cells' counts no longer add up to the records reduced
        if (s_misc[10] || (HAS_CNT && s_tot[0] != s_tot[1])) atomicExch(sw.fail, 1u);
    }
}

template <int MASK>
__global__ void __launch_bounds__(SW_THREADS, SW_CTAS_PER_SM) sweep_kernel(KParams kp, const float4 *__restrict__ pts, long long n, SweepWs sw, Outs out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[SW_STAGES];
    __shared__ int s_misc[16];
    __shared__ float s_div255[256];
    // the same decision in every CTA: nothing in the persistent block changes while this kernel runs
    if (sw.persist->magic != sw.magic || sw.persist->cooldown != 0u) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(sw.fail, 1u);
        return;
    }
    if (blockIdx.x < (unsigned)SW_OWNERS) sweep_consumer<MASK>(kp, sw, out, blockIdx.x, smem_raw, s_misc, s_div255);
    else sweep_producer(kp, pts, n, sw, blockIdx.x - SW_OWNERS, smem_raw, s_bar, s_misc);
}

// first kernel behind a sweep: bookkeeping of the persistent block (runs after the sweep has finished)
__global__ void sweep_epilogue_kernel(SweepWs sw) {
    if (sw.persist->magic != sw.magic) return;
    if (*sw.fail == 2u) sw.persist->magic = 0;                        // watchdog: the mailboxes are in an unknown state
    if (*sw.fail) {
        sw.stats->n_valid = 0;                                         // the two-pass kernels count again
        if (sw.persist->cooldown) --sw.persist->cooldown;              // the sweep was skipped
        else { sw.persist->cooldown = SW_COOLDOWN; ++sw.persist->n_failed; }
    } else {
        ++sw.persist->n_ok;
    }
}

__global__ void sweep_init_kernel(SweepWs sw, size_t heads_words, size_t mail_words) {
    const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < heads_words; i += st) sw.heads[i] = 0u;
    uint32_t *m = reinterpret_cast<uint32_t *>(sw.mail);
    for (size_t i = i0; i < mail_words; i += st) m[i] = SW_PHASE;      // wrap 0 expects phase 0: nothing has arrived
    if (i0 == 0) { sw.persist->magic = sw.magic; sw.persist->cooldown = 0; sw.persist->n_failed = 0; sw.persist->n_ok = 0; }
}
