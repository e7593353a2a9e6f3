// Fused tile-stream linearisation kernel (sm_100a): the default hot path of libpcr_b200.so.
// Included by pcr_linearize.cu (needs LinParams, BlockShared, MatchRec, accumulate_match,
// finish_iteration, load_pose).
//
// One warp = one row of 32 consecutive scan points (cell-ordered on upload):
//   coalesced SoA loads -> float32 SE(3) transform -> cells meeting the box [min - rho, max + rho]
//   of the row -> per (y,z) row of that box ONE bulk copy global -> shared (cp.async.bulk,
//   completion on an mbarrier: SASS UBLKCP + SYNCS) of its contiguous range of pair records ->
//   every lane evaluates every staged candidate (broadcast shared-memory loads, packed f32x2
//   distances, no divergence; pcr_tile.cuh) -> settle test -> larger box for the lanes still open
//   -> matched record gathered once -> residual + Jacobian terms in registers -> float64 reduction
//   -> last block assembles the record and does the GN step.
// No per-point list structure, no parked positions, no second kernel: compulsory traffic is the
// scan, the staged target ranges (L2-resident between neighbouring rows) and one payload gather.
#pragma once

namespace pcr {

constexpr float kTileMinRadius = 0.03125f;      // smallest halo radius (grid units) a row starts with

struct TileParams {
    TileGrid G;
    const float4* pay;        // per indexed point: PLANE / VPLANE normal (1 float4), NDT (2 float4); null for ICP
    float* hint;              // [n_pad / 32] halo radius (grid units) every warp row needed last time (first guess of this one)
    const uint32_t* perm;     // position -> position in the source index (only for match_out)
    int* match_out;           // optional [n_pad]: matched source position per scan slot (-1 none)
    int cap;                  // points per warp stage buffer
    int bulk_min;             // ranges of at least this many pair records are bulk copies (TMA); shorter ones are copied by their lane
    float rmax;               // halo radius at which every lane is settled by construction
    int core_e;               // lanes farther than this many cells from the leader wait for their own pass
    int warp_bytes;           // shared-memory stride between the stage buffers of two warps
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared through the TMA engine, bytes a multiple of 16, both sides 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// per-lane float32 sums of a few rows -> float64 totals, lane t keeps term t (fixed order: deterministic)
template <int NACC>
__device__ __forceinline__ void tile_flush(float* acc, double& acc64, float* scratch, int lane) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NACC; ++i) { scratch[i * 33 + lane] = acc[i]; acc[i] = 0.f; }
    __syncwarp();
    if (lane < NACC) {
        double s = 0.0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) s += (double)scratch[lane * 33 + j];
        acc64 += s;
    }
    __syncwarp();
}

// min / max over the L consecutive lanes of a group (every lane of the warp takes part: xor butterflies
// inside the groups; the whole warp at once is one redux instruction)
template <int L>
__device__ __forceinline__ int group_min(int v) {
    if (L == 32) return __reduce_min_sync(0xffffffffu, v);
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int L>
__device__ __forceinline__ int group_max(int v) {
    if (L == 32) return __reduce_max_sync(0xffffffffu, v);
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// search half of one warp row: the matched point (x, y, z, position bits; position 0xffffffff = no
// correspondence) of every lane is parked in shared memory for the accumulate half.
//
// The warp works as NG independent GROUPS of L = 32 / NG consecutive lanes: every group stages its
// OWN box (the cells around its L scan points) into its own slice of the stage buffer, and its lanes
// scan only that slice.  Scanning costs (lanes x staged candidates); L consecutive scan points span
// ~L / ppc cells, so smaller groups stage far fewer candidates per lane, while all groups still run
// the same instructions (broadcast loads inside a group, NG distinct addresses per request).
template <int NG>
__device__ __forceinline__ void tile_search_row(const LinParams& P, const TileParams& TP, const Pose32* spose, long long row, int lane,
                                                float4* spts, uint64_t* bar, uint32_t& phase, float4* smatch) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int L = 32 / NG;
    const TileGrid& G = TP.G;
    const int gl = lane & (L - 1), gbase = lane & ~(L - 1);     // lane within its group, first lane of the group
    const unsigned gmask = (NG == 1 ? FULL : ((1u << L) - 1u)) << gbase;
    const uint32_t capp = ((uint32_t)TP.cap >> 1) / NG;          // this group's slice of the stage buffer, in pair records
    float4* gpts = spts + 2 * (size_t)capp * (lane / L);
    const long long i = row * 32 + lane;
    const float px = __ldg(P.sx + i), py = __ldg(P.sy + i), pz = __ldg(P.sz + i);
    const float nanf_ = __int_as_float(0x7fc00000);
    TileQuery q;
    q.qx = q.qy = q.qz = nanf_;                                  // a lane without a query compares false against every candidate
    q.gx = q.gy = q.gz = 0.f; q.ix = q.iy = q.iz = 0;
    TileBest b;
    b.d2 = P.max_d2; b.pos = kTileNone; b.x = b.y = b.z = 0.f;
    bool valid = false;
    if (px == px) {                                              // NaN = padding
        float qx, qy, qz;
        transform32(*spose, px, py, pz, qx, qy, qz);
        valid = tile_make_query(G, qx, qy, qz, q) && !tile_query_far_outside(G, q, P.max_d2);
    }
    const float hint_in = TP.hint[row];
    float rl = fminf(fmaxf(hint_in, kTileMinRadius), TP.rmax);   // this lane's halo radius (grid units)
    unsigned todo = __ballot_sync(FULL, valid);
    while (todo) {
        // ---- one pass (per group): leader = first open lane; everybody near it joins, the halo is the largest any of them asks for ----
        const unsigned gtodo = todo & gmask;
        const int leader = gtodo ? __ffs(gtodo) - 1 : lane;
        const int lx = __shfl_sync(FULL, q.ix, leader), ly = __shfl_sync(FULL, q.iy, leader), lz = __shfl_sync(FULL, q.iz, leader);
        const bool elig = ((todo >> lane) & 1u) && abs(q.ix - lx) <= TP.core_e && abs(q.iy - ly) <= TP.core_e && abs(q.iz - lz) <= TP.core_e;
        const float rho = __int_as_float(group_max<L>(elig ? __float_as_int(rl) : 0));   // radii are positive: bit order = value order
        TileBox U;                                               // cells meeting [min - rho, max + rho] of the joined lanes (not clipped)
        U.x0 = group_min<L>(elig ? tile_cell_floor(q.gx - rho) : INT_MAX); U.x1 = group_max<L>(elig ? tile_cell_floor(q.gx + rho) : INT_MIN);
        U.y0 = group_min<L>(elig ? tile_cell_floor(q.gy - rho) : INT_MAX); U.y1 = group_max<L>(elig ? tile_cell_floor(q.gy + rho) : INT_MIN);
        U.z0 = group_min<L>(elig ? tile_cell_floor(q.gz - rho) : INT_MAX); U.z1 = group_max<L>(elig ? tile_cell_floor(q.gz + rho) : INT_MIN);
        TileBox R;                                               // the part of it inside the grid
        R.x0 = max(U.x0, 0); R.x1 = min(U.x1, G.nx - 1);
        R.y0 = max(U.y0, 0); R.y1 = min(U.y1, G.ny - 1);
        R.z0 = max(U.z0, 0); R.z1 = min(U.z1, G.nz - 1);
        const bool box = gtodo != 0u && R.x0 <= R.x1 && R.y0 <= R.y1 && R.z0 <= R.z1;
        const int rnx = R.x1 - R.x0 + 1, rny = box ? R.y1 - R.y0 + 1 : 1, rnz = R.z1 - R.z0 + 1;
        const int nrows = box ? rny * rnz : 0;                   // (y,z) rows of this group's box
        int ra = 0;
        while (__any_sync(FULL, ra < nrows)) {
            // ---- table pass: lane l of the group looks up the pair range of row ra + l of the group's box ----
            const int r = ra + gl;
            const bool rowv = r < nrows;
            uint32_t ps = 0u, lenp = 0u;
            if (rowv) {
                const int jz = R.z0 + r / rny, jy = R.y0 + r % rny;
                const size_t base = ((size_t)jz * G.ny + jy) * G.nx + R.x0;
                const uint32_t gs = __ldg(G.cs + base), ge = __ldg(G.cs + base + rnx);
                if (ge > gs) { ps = gs >> 1; lenp = ((ge + 1u) >> 1) - ps; }   // whole pair records: a neighbour point at either end is harmless
            }
            uint32_t incl = lenp;
#pragma unroll
            for (int o = 1; o < L; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, o, L);
                if (gl >= o) incl += t;
            }
            const int nfit = __popc(__ballot_sync(FULL, rowv && incl <= capp) & gmask);   // a prefix of the group: incl is monotone
            // mode of this group for this piece: 0 nothing left, 1 staged, 2 its next row alone exceeds the slice: scanned in global memory
            const int mode = ra >= nrows ? 0 : (nfit > 0 ? 1 : 2);
            // (every warp-wide shuffle is executed by ALL lanes, whatever their group's mode: a full-mask shuffle
            //  inside a per-group condition would wait for lanes that never arrive)
            const uint32_t total_any = __shfl_sync(FULL, incl, gbase + (nfit > 0 ? nfit - 1 : 0));
            const uint32_t total = mode == 1 ? total_any : 0u;
            const uint32_t all = __reduce_add_sync(FULL, gl == 0 ? total : 0u);
            if (all > 0u) {
                // ---- stage: every non-empty row range -> the group's slice, then every lane scans EVERY candidate its
                //      group staged.  A range of at least bulk_min pair records is ONE bulk copy (TMA engine, completion
                //      on the mbarrier); shorter ones are copied by their lane with 16-byte loads: the TMA engine needs
                //      ~50 ns per copy whatever its size (measured, profiles/r2_notes.md), which bounds a kernel that
                //      issues a dozen 100-byte copies per row ----
                const bool mine = mode == 1 && gl < nfit && lenp > 0u;
                const bool bulk = mine && lenp >= (uint32_t)TP.bulk_min;
                const uint32_t bulk_pairs = __reduce_add_sync(FULL, bulk ? lenp : 0u);
                if (bulk_pairs) {
                    if (lane == 0) mbar_expect_tx(bar, bulk_pairs * 32u);
                    __syncwarp();
                    if (bulk) bulk_g2s(gpts + 2 * (incl - lenp), G.pairs + 2 * (size_t)ps, lenp * 32u, bar);
                }
                if (mine && !bulk) {
                    const float4* src = G.pairs + 2 * (size_t)ps;
                    float4* dst = gpts + 2 * (incl - lenp);
                    for (uint32_t k = 0; k < 2u * lenp; ++k) dst[k] = __ldg(src + k);
                }
                if (bulk_pairs) {
                    mbar_wait(bar, phase);
                    phase ^= 1u;
                }
                __syncwarp();                                    // lane-copied ranges are visible to the whole warp
                float best = b.d2;
                const uint32_t trip = tile_scan_pairs(gpts, total, q, best);
                if (trip != kTileNone) tile_take(gpts, trip, total, q, best, b);
                __syncwarp();                                    // everybody is done reading before the next piece overwrites
            }
            if (__any_sync(FULL, mode == 2)) {
                const uint32_t ps0 = __shfl_sync(FULL, ps, gbase), np_any = __shfl_sync(FULL, lenp, gbase);
                const uint32_t np0 = mode == 2 ? np_any : 0u;
                const float4* gp = G.pairs + 2 * (size_t)ps0;
                float best = b.d2;
                const uint32_t trip = tile_scan_pairs(gp, np0, q, best);
                if (trip != kTileNone) tile_take(gp, trip, np0, q, best, b);
            }
            ra += mode == 1 ? nfit : (mode == 2 ? 1 : 0);
        }
        // ---- settle: final iff nothing outside the staged box can be closer; else ask for the ball of the
        //      candidate (settles for certain next time) or, with no candidate yet, for twice the radius ----
        bool settled = false;
        if (elig) {
            settled = rho >= TP.rmax || tile_settled(G, q, b, U);
            if (!settled) rl = fminf(b.pos != kTileNone ? tile_radius_for(G, b) : 2.0f * rho, TP.rmax);
        }
        todo &= ~__ballot_sync(FULL, settled);
    }
    if (smatch) smatch[lane] = make_float4(b.x, b.y, b.z, __uint_as_float(b.pos));
    if (TP.match_out) TP.match_out[i] = b.pos != kTileNone ? (int)TP.perm[b.pos] : -1;
    // first guess of the next linearisation: a little more than the farthest correspondence of this row
    float need = b.pos != kTileNone ? sqrtf(b.d2) * G.inv_c : 0.0f;
    need = __int_as_float(__reduce_max_sync(FULL, __float_as_int(need)));
    need = need * 1.25f + 0.02f;
    if (lane == 0 && need != hint_in) TP.hint[row] = need;
}

// accumulate half: matched record -> residual + Jacobian terms of this lane's correspondence
template <int METHOD>
__device__ __forceinline__ void tile_accumulate_row(const LinParams& P, const TileParams& TP, const Pose32* spose, long long row, int lane,
                                                    const float4* smatch, float* acc) {
    const float4 m = smatch[lane];
    const uint32_t pos = __float_as_uint(m.w);
    if (pos == kTileNone) return;
    const long long i = row * 32 + lane;
    MatchRec r;
    r.a = make_float4(m.x, m.y, m.z, 0.f);
    if (METHOD == PCR_METHOD_PLANE || METHOD == PCR_METHOD_VPLANE) {
        r.b = __ldg(TP.pay + pos);
    } else if (METHOD == PCR_METHOD_NDT) {
        const float4 p0 = __ldg(TP.pay + 2 * (size_t)pos), p1 = __ldg(TP.pay + 2 * (size_t)pos + 1);
        r.a.w = p0.x;
        r.b = make_float4(p0.y, p0.z, p0.w, p1.x);
        r.c = make_float4(p1.y, 0.f, 0.f, 0.f);
    }
    const float px = __ldg(P.sx + i), py = __ldg(P.sy + i), pz = __ldg(P.sz + i);
    const Pose32 pose = *spose;
    accumulate_match<METHOD>(pose, acc, r, px, py, pz);
}

// KR consecutive warp rows form one unit of work: their searches run first (the 29 accumulators are
// dead meanwhile -- the search keeps its registers), then their terms are accumulated and flushed
template <int METHOD, int MINB, int KR, int NG>
__global__ void __launch_bounds__(kLinThreads, MINB) tile_linearize_kernel(const LinParams P, const TileParams TP) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ BlockShared sh;
    __shared__ Pose32 spose;
    constexpr int NACC = NAcc<METHOD>::value;
    constexpr int NWARP = kLinThreads / 32;
    {
        Pose32 pose;
        if (!load_pose(P, sh, pose)) return;
        if (threadIdx.x == 0) spose = pose;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = tile_smem + (size_t)warp * TP.warp_bytes;
    float4* spts = reinterpret_cast<float4*>(wbase);
    uint64_t* bar = reinterpret_cast<uint64_t*>(wbase + (size_t)TP.cap * 16);
    float4* smatch = reinterpret_cast<float4*>(wbase + (size_t)TP.cap * 16 + 16);   // [KR][32]
    if (lane == 0) {
        mbar_init(bar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0u;
    double acc64 = 0.0;
    const long long rows = P.n_pad >> 5;
    const long long groups = (rows + KR - 1) / KR;
    const long long nwarps = (long long)gridDim.x * NWARP;
    for (long long grp = (long long)blockIdx.x * NWARP + warp; grp < groups; grp += nwarps) {
        const long long row0 = grp * KR;
#pragma unroll 1
        for (int u = 0; u < KR; ++u)
            if (row0 + u < rows) tile_search_row<NG>(P, TP, &spose, row0 + u, lane, spts, bar, phase, smatch + u * 32);
        __syncwarp();
        float acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
#pragma unroll 1
        for (int u = 0; u < KR; ++u)
            if (row0 + u < rows) tile_accumulate_row<METHOD>(P, TP, &spose, row0 + u, lane, smatch + u * 32, acc);
        tile_flush<NACC>(acc, acc64, reinterpret_cast<float*>(spts), lane);
    }
    if (lane < NACC) sh.red[warp][lane] = acc64;
    __syncthreads();
    block_finish<METHOD>(P, sh);
}

// Split form: correspondences only (the matched position in the SOURCE index is parked per scan
// slot, TP.match_out), followed by accumulate_kernel.  No accumulators and no method dependence: the
// kernel keeps few registers, so twice as many warps are resident to hide the latency chain of a row
// (scan load -> cell starts -> bulk copy -> scan), which is what bounds the fused form.
template <int MINB, int NG>
__global__ void __launch_bounds__(kLinThreads, MINB) tile_correspond_kernel(const LinParams P, const TileParams TP) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ BlockShared sh;
    __shared__ Pose32 spose;
    constexpr int NWARP = kLinThreads / 32;
    {
        Pose32 pose;
        if (!load_pose(P, sh, pose)) return;
        if (threadIdx.x == 0) spose = pose;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = tile_smem + (size_t)warp * TP.warp_bytes;
    float4* spts = reinterpret_cast<float4*>(wbase);
    uint64_t* bar = reinterpret_cast<uint64_t*>(wbase + (size_t)TP.cap * 16);
    if (lane == 0) {
        mbar_init(bar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0u;
    const long long rows = P.n_pad >> 5;
    const long long nwarps = (long long)gridDim.x * NWARP;
    for (long long row = (long long)blockIdx.x * NWARP + warp; row < rows; row += nwarps)
        tile_search_row<NG>(P, TP, &spose, row, lane, spts, bar, phase, nullptr);
}

}  // namespace pcr
