// The per-iteration hot path of libpcr_b200.so (sm_100a).  One Gauss-Newton linearisation =
//
//   correspond   coalesced loads of the SoA scan -> SE(3) transform (float32) -> exact nearest
//                neighbour: every query streams the list of its cell (voxel means: exact candidate
//                list; target points: margin-ordered shell list);
//                the few queries a list cannot settle are searched in the brick grid
//                -> matched position parked per scan slot (4 B/point)
//   accumulate   gather the matched record -> residual + 6-DoF Jacobian terms -> per-thread float32
//                sums -> float64 warp shuffles + shared-memory block reduction -> per-block partial
//                -> the LAST block to arrive sums the partials in a fixed order (deterministic),
//                assembles the 29-entry record and (on-device loop) performs the 6x6 solve, the stop
//                test and the SE(3) update.
//
// The two passes run as two kernels (default: the latency-bound list streaming wants many resident
// warps, the 30-accumulator reduction wants registers) or fused in one (PCR_SPLIT=0, A/B).
// Replaces calc_H_g_e2 of icp.py:24-57, plane_icp.py:30-69, voxelized_plane_icp.py:23-64,
// ndt.py:24-57 and the loop of registration.py:89-111.  No tensor cores: the path is a
// gather + 29-term reduction (arithmetic intensity ~3 flop/B).
#include <dlfcn.h>

#include <cub/cub.cuh>

#include "pcr_context.cuh"
#include "pcr_grid.cuh"
#include "pcr_linalg.cuh"
#include "pcr_terms.cuh"

namespace pcr {

struct LinParams {
    const float* sx; const float* sy; const float* sz;   // SoA scan, padded to a multiple of 32 with NaN
    long long n_pad;                                     // padded point count
    GridView grid;                                       // target points (ICP/PLANE) or voxel means (VPLANE/NDT)
    const float4* pn;                                    // PLANE: (point, normal) records in grid order, 32 B each
    const float4* vrec;                                  // VPLANE: 2 float4 / voxel, NDT: 3 float4 / voxel
    CandLists lists;                                     // VPLANE/NDT: per-cell candidate lists (null = absent)
    int use_lists;
    ShellLists shell;                                    // ICP/PLANE: per-cell shell lists (null = absent)
    int use_shell;
    int ball_first;                                      // list misses: one pass over the ball of max_dist instead of ring growth
    int grab_rows;                                       // rows of 32 scan slots a warp fetches at a time (correspondence pass)
    int* prev;                                           // per scan slot: position matched by the previous linearisation (-1 none)
    float max_d2;
    double T_param[16];
    int use_param_T;          // 1: transform comes from T_param, 0: from st->T (device loop)
    int device_loop;          // 1: last block performs the Gauss-Newton step on st
    int max_iter;
    double tol;
    LoopState* st;
    double* partials;         // [gridDim.x][PCR_NEQ_PAD]
    double* out_mapped;       // pinned host memory (may be NULL): rec[32], T[16], iter, done
};

// accumulators per thread
template <int METHOD> struct NAcc { static constexpr int value = PCR_NEQ; };
template <> struct NAcc<PCR_METHOD_ICP> { static constexpr int value = 17; };

struct BlockShared {
    double T[16];
    double red[kLinThreads / 32][PCR_NEQ_PAD];
    double sum[PCR_NEQ_PAD];
    int flag;
};

// the matched record (fetched first, for several slots at once, so that the gathers overlap) ...
struct MatchRec { float4 a, b, c; };

template <int METHOD>
__device__ __forceinline__ void fetch_match(const LinParams& P, int pos, MatchRec& r) {
    if (pos < 0) return;
    if (METHOD == PCR_METHOD_ICP) {
        r.a = __ldg(P.grid.pts + pos);
    } else if (METHOD == PCR_METHOD_PLANE) {
        r.a = __ldg(P.pn + 2 * (size_t)pos);
        r.b = __ldg(P.pn + 2 * (size_t)pos + 1);
    } else if (METHOD == PCR_METHOD_VPLANE) {
        r.a = __ldg(P.vrec + 2 * (size_t)pos);
        r.b = __ldg(P.vrec + 2 * (size_t)pos + 1);
    } else {
        r.a = __ldg(P.vrec + 3 * (size_t)pos);
        r.b = __ldg(P.vrec + 3 * (size_t)pos + 1);
        r.c = __ldg(P.vrec + 3 * (size_t)pos + 2);
    }
}

// ... and this correspondence's terms
template <int METHOD, typename A>
__device__ __forceinline__ void accumulate_match(const Pose32& pose, A* acc, const MatchRec& r, float px, float py, float pz) {
    float qx, qy, qz;
    transform32(pose, px, py, pz, qx, qy, qz);
    if (METHOD == PCR_METHOD_ICP) {
        accum_icp(acc, pose, px, py, pz, qx - r.a.x, qy - r.a.y, qz - r.a.z);
    } else if (METHOD == PCR_METHOD_PLANE || METHOD == PCR_METHOD_VPLANE) {
        accum_plane(acc, pose, px, py, pz, qx - r.a.x, qy - r.a.y, qz - r.a.z, r.b.x, r.b.y, r.b.z);
    } else {
        const float w6[6] = {r.a.w, r.b.x, r.b.y, r.b.z, r.b.w, r.c.x};
        accum_ndt(acc, pose, px, py, pz, qx - r.a.x, qy - r.a.y, qz - r.a.z, w6);
    }
}

// Last block only, one thread: assemble the record, publish it, optionally do the GN step.
// (Out of line, and fed a small by-value parameter block: taking the address of the kernel's
// parameter struct would copy all of it to local memory at the start of every thread.)
struct FinishParams {
    LoopState* st;
    double* out_mapped;
    double tol;
    int device_loop;
    int max_iter;
};

template <int METHOD>
__device__ __noinline__ void finish_iteration(const FinishParams P, BlockShared& sh) {
    LoopState* st = P.st;
    double rec[PCR_NEQ_PAD];
    if (METHOD == PCR_METHOD_ICP) {
        assemble_icp(sh.sum, sh.T, rec);
    } else {
        for (int i = 0; i < PCR_NEQ; ++i) rec[i] = sh.sum[i];
    }
    for (int i = 0; i < PCR_NEQ; ++i) st->rec[i] = rec[i];
    st->ticket = 0u;
    st->next_row = 0;                                     // every block is past its correspondence pass: re-arm the row counter
    int iter = st->iter;
    int done = 0;
    if (P.device_loop) {
        if (iter < kMaxTrace) st->e2_trace[iter] = rec[27];
        iter += 1;
        double T[16];
        for (int i = 0; i < 16; ++i) T[i] = sh.T[i];
        double dx[6], dxn = 0.0;
        const int rc = gauss_newton_step(rec, P.tol, T, dx, &dxn);
        if (rc == 0) {
            for (int i = 0; i < 16; ++i) st->T[i] = T[i];
            if (iter >= P.max_iter) done = 3;            // iteration budget exhausted
        } else {
            done = rc;                                    // 1 converged, 2 singular
        }
        for (int i = 0; i < 6; ++i) st->dx[i] = dx[i];
        st->dx_norm = dxn;
        st->iter = iter;
        st->done = done;
    }
    if (P.out_mapped) {
        for (int i = 0; i < PCR_NEQ; ++i) P.out_mapped[i] = rec[i];
        for (int i = 0; i < 16; ++i) P.out_mapped[32 + i] = P.device_loop ? st->T[i] : sh.T[i];
        P.out_mapped[48] = (double)iter;
        P.out_mapped[49] = (double)done;
        __threadfence_system();
    }
}

// float32 per-thread sums -> float64 warp shuffles -> shared memory -> per-block partial ->
// the last block to arrive reduces all partials in a fixed order and finishes the iteration.
// second half of the reduction: sh.red[warp][term] holds the float64 sums of every warp (the
// caller has synchronised the block) -> per-block partial -> last block finishes the iteration
template <int METHOD>
__device__ __forceinline__ void block_finish(const LinParams& P, BlockShared& sh) {
    constexpr int NRED = NAcc<METHOD>::value;
    constexpr int NWARP = kLinThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < NRED) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) v += sh.red[w][threadIdx.x];
        P.partials[(size_t)blockIdx.x * PCR_NEQ_PAD + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&P.st->ticket, 1u);
        sh.flag = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!sh.flag) return;
    __threadfence();
    for (int i = warp; i < NRED; i += NWARP) {
        double v = 0.0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) v += __ldcg(P.partials + (size_t)b * PCR_NEQ_PAD + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh.sum[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        FinishParams F;
        F.st = P.st; F.out_mapped = P.out_mapped; F.tol = P.tol; F.device_loop = P.device_loop; F.max_iter = P.max_iter;
        finish_iteration<METHOD>(F, sh);
    }
}

template <int METHOD, typename A>
__device__ __forceinline__ void reduce_and_finish(const LinParams& P, BlockShared& sh, const A* acc) {
    constexpr int NRED = NAcc<METHOD>::value;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NRED; ++i) {
        double v = (double)acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh.red[warp][i] = v;
    }
    __syncthreads();
    block_finish<METHOD>(P, sh);
}

__device__ __forceinline__ bool load_pose(const LinParams& P, BlockShared& sh, Pose32& pose) {
    LoopState* st = P.st;
    if (threadIdx.x == 0) sh.flag = P.use_param_T ? 0 : *((volatile int*)&st->done);   // device loop finished?
    if (threadIdx.x < 16) sh.T[threadIdx.x] = P.use_param_T ? P.T_param[threadIdx.x] : ((volatile double*)st->T)[threadIdx.x];
    __syncthreads();
    if (sh.flag) return false;
    pose32_from_T(sh.T, pose);
    return true;
}

// The general brick-grid search as an out-of-line call: it only serves stragglers, and inlined it
// would set the register budget (and so the occupancy) of the list-streaming loop around it.
// far: the query's list found nothing near it -- one pruned pass over the ball of max_dist instead of
// growing ring by ring (see grid_search).
__device__ __noinline__ int general_nn(const GridView G, float qx, float qy, float qz, float max_d2, bool far) {   // by value: the kernel
    // parameter block stays in the constant bank instead of being copied to local memory for its address
    float d2;
    return grid_nn(G, qx, qy, qz, max_d2, d2, far);
}

// Continuation of an exhausted shell list (out of line, like general_nn).
__device__ __noinline__ int shell_continue_nn(const GridView G, const ShellLists S, float qx, float qy, float qz, float d2, int pos) {
    shell_continue(G, S, qx, qy, qz, d2, pos);
    return pos;
}

// ---- pass 1: correspondences ---------------------------------------------------------------------
// Every query streams the list of its cell -- lanes of one cell read the same addresses for nearly
// the same number of steps.  The few queries a list cannot settle (no list for the cell, or the
// best still beyond the listed margin) run the general brick-grid search in place.
// one scan slot:
template <int METHOD>
__device__ __forceinline__ void correspond_point(const LinParams& P, const Pose32& pose, long long i, bool lists, float px, float py, float pz) {
    constexpr bool kVoxel = METHOD == PCR_METHOD_VPLANE || METHOD == PCR_METHOD_NDT;
    int pos = -1;
    if (px == px) {                                              // NaN = padding: no match
        float qx, qy, qz, d2;
        transform32(pose, px, py, pz, qx, qy, qz);
        bool settled = false;
        if (lists) {
            if (P.use_shell) {
                // margin-ordered shell list of the query's cell (target points, or kept voxel means)
                const int st = shell_scan(P.grid, P.shell, qx, qy, qz, P.max_d2, d2, pos);
                settled = st != 0;
                // list exhausted before the best was proven nearest: the search resumes outside
                // the block of cells the list covered, pruned by the list's best
                if (st == 2) pos = shell_continue_nn(P.grid, P.shell, qx, qy, qz, d2, pos);
            } else if (kVoxel) {
                settled = list_nn(P.grid, P.lists, qx, qy, qz, P.max_d2, d2, pos);   // exact candidate list (A/B: PCR_VOXEL_SHELL=0)
            }
        }
        if (!settled) pos = general_nn(P.grid, qx, qy, qz, P.max_d2, lists && P.ball_first);
    }
    P.prev[i] = pos;
}

template <int METHOD>
__device__ __forceinline__ void correspond_slot(const LinParams& P, const Pose32& pose, long long i, bool lists) {
    correspond_point<METHOD>(P, pose, i, lists, __ldg(P.sx + i), __ldg(P.sy + i), __ldg(P.sz + i));
}

// DYNAMIC: warps fetch rows of 32 consecutive scan slots from a device-wide counter (P.grab_rows
// rows per fetch) -- the cost of a row varies with the local geometry, and a static partition left
// the slowest SM 30 % behind the average.  The parked result of a slot does not depend on who
// computed it and the accumulate pass keeps its fixed order, so results stay deterministic.  Only
// possible when the accumulate pass is a separate kernel (it reads slots parked by other blocks).
// (Measured and removed: a block-level queue that collected the list misses and searched them at
// the end of the block -- with dynamic rows it only serialises the stragglers into a tail; and two
// list streams per lane -- no gain, the L1 tag stage is already 80 % busy, profiles/r1_notes.md.)
template <int METHOD, bool DYNAMIC>
__device__ __forceinline__ void correspond_pass(const LinParams& P, const Pose32& pose) {
    constexpr bool kVoxel = METHOD == PCR_METHOD_VPLANE || METHOD == PCR_METHOD_NDT;
    const int lane = threadIdx.x & 31;
    const bool lists = P.use_shell != 0 || (kVoxel && P.use_lists != 0);
    if (DYNAMIC) {
        const int rows_total = (int)(P.n_pad >> 5);
        for (;;) {
            int r0 = 0;
            if (lane == 0) r0 = atomicAdd(&P.st->next_row, P.grab_rows);
            r0 = __shfl_sync(0xffffffffu, r0, 0);
            if (r0 >= rows_total) break;
            const int r1 = r0 + P.grab_rows < rows_total ? r0 + P.grab_rows : rows_total;
            // software pipeline over the rows of one fetch: the next row's coordinates are requested before this row's
            // list is streamed (the chain scan load -> brick record -> list start -> entries is latency bound)
            long long i = ((long long)r0 << 5) + lane;
            float px = __ldg(P.sx + i), py = __ldg(P.sy + i), pz = __ldg(P.sz + i);
            for (int r = r0; r < r1; ++r) {
                const long long in = i + 32;
                float nx = 0.f, ny = 0.f, nz = 0.f;
                if (r + 1 < r1) { nx = __ldg(P.sx + in); ny = __ldg(P.sy + in); nz = __ldg(P.sz + in); }
                correspond_point<METHOD>(P, pose, i, lists, px, py, pz);
                i = in; px = nx; py = ny; pz = nz;
            }
        }
    } else {
        // static partition of the fused form: the SAME quads of slots the thread accumulates afterwards
        // (accumulate_pass reads the positions parked here without any grid-wide synchronisation)
        const long long stride = (long long)gridDim.x * kLinThreads;
        for (long long t = blockIdx.x * (long long)kLinThreads + threadIdx.x; t < (P.n_pad >> 2); t += stride)
            for (int u = 0; u < 4; ++u) correspond_slot<METHOD>(P, pose, 4 * t + u, lists);
    }
}

// ---- pass 2: residual + Jacobian terms of the parked correspondences, reduction, GN step ----------
// Every thread owns FOUR consecutive scan slots per trip: one 16-byte load each for the parked
// positions and the three scan coordinates (instead of sixteen scalar loads: the pass used to be
// bound by the load/store queue, ncu lg_throttle 2.7 per issue), then the four matched records are
// requested together (the gathers are latency bound) before the first is consumed.  Terms are added
// in slot order: the summation order is fixed.
template <int METHOD>
__device__ __forceinline__ void accumulate_pass(const LinParams& P, BlockShared& sh, const Pose32& pose) {
    constexpr int NACC = NAcc<METHOD>::value;
    const long long quads = P.n_pad >> 2;                        // n_pad is a multiple of 32
    const long long stride = (long long)gridDim.x * kLinThreads;
    // float64 accumulators: per-point terms are float32 products (as the reference's float32 geometry), every
    // SUM is float64 -- the reference sums in float64 after the gather (ndt.py:39-56), and records no longer
    // depend on how the scan is split over threads or GPUs beyond 1e-13.  (Measured and dropped: summing the
    // four slots of a trip in float32 first -- no faster, and align() then differs by 4e-8 between scan orders.)
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
    const int4* prev4 = reinterpret_cast<const int4*>(P.prev);
    const float4* sx4 = reinterpret_cast<const float4*>(P.sx);
    const float4* sy4 = reinterpret_cast<const float4*>(P.sy);
    const float4* sz4 = reinterpret_cast<const float4*>(P.sz);
    for (long long t = blockIdx.x * (long long)kLinThreads + threadIdx.x; t < quads; t += stride) {
        const int4 pq = prev4[t];
        const int pos[4] = {pq.x, pq.y, pq.z, pq.w};
        const float4 X = __ldg(sx4 + t), Y = __ldg(sy4 + t), Z = __ldg(sz4 + t);
        const float px[4] = {X.x, X.y, X.z, X.w}, py[4] = {Y.x, Y.y, Y.z, Y.w}, pz[4] = {Z.x, Z.y, Z.z, Z.w};
        constexpr int B = METHOD == PCR_METHOD_NDT ? 2 : 4;      // records in flight (an NDT record is 12 registers)
#pragma unroll
        for (int h = 0; h < 4; h += B) {
            MatchRec rec[B];
#pragma unroll
            for (int u = 0; u < B; ++u) fetch_match<METHOD>(P, pos[h + u], rec[u]);
#pragma unroll
            for (int u = 0; u < B; ++u)
                if (pos[h + u] >= 0) accumulate_match<METHOD>(pose, acc, rec[u], px[h + u], py[h + u], pz[h + u]);
        }
    }
    reduce_and_finish<METHOD>(P, sh, acc);
}

// split form (default): two kernels, each with the occupancy it wants
template <int METHOD, int MINB>
__global__ void __launch_bounds__(kLinThreads, MINB) correspond_kernel(const LinParams P) {
    __shared__ BlockShared sh;
    Pose32 pose;
    if (!load_pose(P, sh, pose)) return;                         // loop already finished: nothing to do
    correspond_pass<METHOD, true>(P, pose);
}

template <int METHOD, int MINB>
__global__ void __launch_bounds__(kLinThreads, MINB) accumulate_kernel(const LinParams P) {
    __shared__ BlockShared sh;
    Pose32 pose;
    if (!load_pose(P, sh, pose)) return;
    accumulate_pass<METHOD>(P, sh, pose);
}

// fused form (PCR_SPLIT=0): both passes in one kernel; thread t accumulates the slots it searched
template <int METHOD, int MINB>
__global__ void __launch_bounds__(kLinThreads, MINB) linearize_fused_kernel(const LinParams P) {
    __shared__ BlockShared sh;
    Pose32 pose;
    if (!load_pose(P, sh, pose)) return;
    correspond_pass<METHOD, false>(P, pose);
    accumulate_pass<METHOD>(P, sh, pose);
}

}  // namespace pcr
#include "pcr_tile_kernel.cuh"
namespace pcr {

// Per-correspondence Gauss-Newton rows of the point-to-plane variants: the 28 numbers whose SUM over the
// scan is the normal-equation record -- upper triangle of J^T J (21), J r (6), r^2 (1), with
// J = [n^T, (p x R^T n)^T], r = n.(R p + t - q) -- i.e. column i of the reference's
// caratheodory.create_gn_set(J, r) (caratheodory.py:118-138).  Zeros for slots without correspondence.
template <int METHOD>
__global__ void gn_rows_kernel(const LinParams P, double* __restrict__ rows, long long n) {
    __shared__ BlockShared sh;
    Pose32 pose;
    if (!load_pose(P, sh, pose)) return;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* out = rows + 28 * i;
    const int pos = P.prev[i];
    if (pos < 0) {
        for (int k = 0; k < 28; ++k) out[k] = 0.0;
        return;
    }
    MatchRec r;
    fetch_match<METHOD>(P, pos, r);
    const float px = P.sx[i], py = P.sy[i], pz = P.sz[i];
    float acc[PCR_NEQ];
    for (int k = 0; k < PCR_NEQ; ++k) acc[k] = 0.f;
    accumulate_match<METHOD>(pose, acc, r, px, py, pz);
    for (int k = 0; k < 28; ++k) out[k] = (double)acc[k];
}

__global__ void matches_kernel(const int* __restrict__ prev, const float4* __restrict__ pts, long long n, long long* __restrict__ idx) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pos = prev[i];
    idx[i] = pos >= 0 ? (long long)__float_as_uint(pts[pos].w) : -1ll;
}

// Gauss-Newton step as its own tiny kernel (multi-GPU path: runs after the all-reduce).
__global__ void gn_step_kernel(LoopState* st, double tol, int max_iter) {
    if (threadIdx.x != 0 || st->done) return;
    int iter = st->iter;
    if (iter < kMaxTrace) st->e2_trace[iter] = st->rec[27];
    iter += 1;
    double T[16], rec[PCR_NEQ_PAD], dx[6], dxn = 0.0;
    for (int i = 0; i < 16; ++i) T[i] = st->T[i];
    for (int i = 0; i < PCR_NEQ; ++i) rec[i] = st->rec[i];
    int done = 0;
    const int rc = gauss_newton_step(rec, tol, T, dx, &dxn);
    if (rc == 0) {
        for (int i = 0; i < 16; ++i) st->T[i] = T[i];
        if (iter >= max_iter) done = 3;
    } else {
        done = rc;
    }
    for (int i = 0; i < 6; ++i) st->dx[i] = dx[i];
    st->dx_norm = dxn;
    st->iter = iter;
    st->done = done;
}

__global__ void fill_float_kernel(float* p, long long n, float v) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct T16 { double v[16]; };
__global__ void loop_init_kernel(LoopState* st, T16 T0) {
    if (threadIdx.x < 16) st->T[threadIdx.x] = T0.v[threadIdx.x];
    if (threadIdx.x == 0) { st->iter = 0; st->done = 0; st->ticket = 0u; st->next_row = 0; st->dx_norm = 0.0; }
}

// ---------------------------------------------------------------------------------------
// scan upload: AoS float3 -> (optionally Morton-sorted) SoA, NaN padded to a multiple of 4
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// mm = {min x, y, z, max x, y, z} as order-preserving ints (scan_bbox_kernel); read on the device
// so that the upload needs no host round trip between the bounding box and the sort keys.
__device__ __forceinline__ float ord2f_dev(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void morton_key_kernel(const float* __restrict__ xyz, long long n, const int* __restrict__ mm,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ox = 0.f, oy = 0.f, oz = 0.f, scale = 0.f;
    if (mm[0] != INT_MAX) {                                   // at least one finite point
        ox = ord2f_dev(mm[0]); oy = ord2f_dev(mm[1]); oz = ord2f_dev(mm[2]);
        const float ext = fmaxf(fmaxf(ord2f_dev(mm[3]) - ox, ord2f_dev(mm[4]) - oy), ord2f_dev(mm[5]) - oz);
        scale = ext > 0.f ? 1023.999f / ext : 0.f;
    }
    float fx = (xyz[3 * i] - ox) * scale, fy = (xyz[3 * i + 1] - oy) * scale, fz = (xyz[3 * i + 2] - oz) * scale;
    uint32_t ix = (uint32_t)fminf(fmaxf(fx == fx ? fx : 0.f, 0.f), 1023.f);
    uint32_t iy = (uint32_t)fminf(fmaxf(fy == fy ? fy : 0.f, 0.f), 1023.f);
    uint32_t iz = (uint32_t)fminf(fmaxf(fz == fz ? fz : 0.f, 0.f), 1023.f);
    keys[i] = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
    vals[i] = (uint32_t)i;
}

// Sort key of a scan point = the cell of the correspondence grid its image under the given pose
// falls into (brick-major, as the grid itself is ordered): the 32 slots of a warp row then sit in
// one or two cells and stream the SAME list, instead of 4-8 cells with a Morton order in the
// scan's own frame.  A rigid motion moves the points of one cell together, so the order stays
// coherent over the Gauss-Newton iterations.
template <typename KEY>
__global__ void cell_key_kernel(const float* __restrict__ xyz, long long n, GridView G, Pose32 pose,
                                KEY* __restrict__ keys, uint32_t* __restrict__ vals) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float qx, qy, qz;
    transform32(pose, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], qx, qy, qz);
    const float big = 1.0e9f;
    float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
    if (!(gx == gx)) gx = 0.f;
    if (!(gy == gy)) gy = 0.f;
    if (!(gz == gz)) gz = 0.f;
    const int cx = cell_of(fminf(fmaxf(gx, -big), big), G.cnx);
    const int cy = cell_of(fminf(fmaxf(gy, -big), big), G.cny);
    const int cz = cell_of(fminf(fmaxf(gz, -big), big), G.cnz);
    const unsigned long long brick = ((unsigned long long)(cz >> 2) * G.bny + (cy >> 2)) * G.bnx + (cx >> 2);
    keys[i] = (KEY)(brick * 64ull + (unsigned long long)brick_bit(cx, cy, cz));
    vals[i] = (uint32_t)i;
}

__global__ void scan_to_soa_kernel(const float* __restrict__ xyz, const uint32_t* __restrict__ order, long long n, long long n_pad,
                                   float* __restrict__ sx, float* __restrict__ sy, float* __restrict__ sz) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i < n) {
        const size_t j = order ? (size_t)order[i] : (size_t)i;
        sx[i] = xyz[3 * j]; sy[i] = xyz[3 * j + 1]; sz[i] = xyz[3 * j + 2];
    } else {
        const float nan = __int_as_float(0x7fc00000);
        sx[i] = nan; sy[i] = nan; sz[i] = nan;
    }
}

__global__ void scan_bbox_kernel(const float* __restrict__ xyz, long long n, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[3 * i + a];
            if (isfinite(v)) {
                int o = __float_as_int(v);
                o = o >= 0 ? o : o ^ 0x7fffffff;
                lo[a] = min(lo[a], o); hi[a] = max(hi[a], o);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&mm[a], lo[a]); atomicMax(&mm[3 + a], hi[a]); }
    }
}

__global__ void mm_init_kernel(int* mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 6) mm[threadIdx.x] = INT_MIN;
}

int ensure_loop_buffers(pcr_ctx* ctx) {
    PCR_CUDA(ctx->partials.ensure((size_t)kMaxLinBlocks * PCR_NEQ_PAD * sizeof(double)));
    PCR_CUDA(ctx->state.ensure(sizeof(LoopState)));
    PCR_CUDA(cudaMemsetAsync(ctx->state.p, 0, sizeof(LoopState), ctx->stream));
    if (!ctx->h_state) PCR_CUDA(cudaHostAlloc((void**)&ctx->h_state, sizeof(LoopState), cudaHostAllocDefault));
    if (!ctx->h_out) {
        PCR_CUDA(cudaHostAlloc((void**)&ctx->h_out, 64 * 8 * sizeof(double), cudaHostAllocMapped));   // 8 record slots (pcr_linearize_host: one per chunk)
        memset(ctx->h_out, 0, 64 * 8 * sizeof(double));
        PCR_CUDA(cudaHostGetDevicePointer((void**)&ctx->d_out_mapped, ctx->h_out, 0));
    }
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    return PCR_OK;
}

// ---------------------------------------------------------------------------------------
// NCCL, resolved at run time so that single-GPU use has no dependency on it
// ---------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*ncclGetUniqueId_t)(nccl_uid_t*);
typedef int (*ncclCommInitRank_t)(void**, int, nccl_uid_t, int);
typedef int (*ncclCommDestroy_t)(void*);
typedef int (*ncclAllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*ncclGetErrorString_t)(int);
constexpr int kNcclFloat64 = 8;   // ncclDouble
constexpr int kNcclSum = 0;       // ncclSum

struct NcclApi {
    void* lib = nullptr;
    ncclGetUniqueId_t GetUniqueId = nullptr;
    ncclCommInitRank_t CommInitRank = nullptr;
    ncclCommDestroy_t CommDestroy = nullptr;
    ncclAllReduce_t AllReduce = nullptr;
    ncclGetErrorString_t GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int load_nccl(pcr_ctx* ctx) {
    if (g_nccl.lib) return PCR_OK;
    const char* env = getenv("PCR_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(ctx, PCR_ERR_NCCL, std::string("cannot dlopen libnccl.so.2 (set PCR_NCCL_LIB): ") + (dlerror() ? dlerror() : ""));
    g_nccl.GetUniqueId = (ncclGetUniqueId_t)dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (ncclCommInitRank_t)dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (ncclCommDestroy_t)dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (ncclAllReduce_t)dlsym(lib, "ncclAllReduce");
    g_nccl.GetErrorString = (ncclGetErrorString_t)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
        return fail(ctx, PCR_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.lib = lib;
    return PCR_OK;
}

static std::string nccl_err(int rc) {
    return g_nccl.GetErrorString ? std::string(g_nccl.GetErrorString(rc)) : ("nccl error " + std::to_string(rc));
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
template <typename K>
static int blocks_per_sm(K kernel) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kLinThreads, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
    return nb;
}

static int lin_grid_blocks(pcr_ctx* ctx, long long n_pad, int per_sm) {
    long long want = (n_pad + kLinThreads - 1) / kLinThreads;
    long long cap = (long long)ctx->sm_count * per_sm;
    if (cap > kMaxLinBlocks) cap = kMaxLinBlocks;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

static int check_method(pcr_ctx* ctx, int method) {
    switch (method) {
        case PCR_ICP:
            if (!ctx->tgt_grid.built) return fail(ctx, PCR_ERR_STATE, "ICP: target NN index not built");
            break;
        case PCR_PLANE:
            if (!ctx->tgt_grid.built) return fail(ctx, PCR_ERR_STATE, "PlaneICP: target NN index not built");
            if (!ctx->has_normals) return fail(ctx, PCR_ERR_STATE, "PlaneICP: target normals not set");
            break;
        case PCR_VPLANE:
            if (!ctx->has_voxels) return fail(ctx, PCR_ERR_STATE, "VPlaneICP: voxels not built");
            break;
        case PCR_NDT:
            if (!ctx->has_voxels || !ctx->has_icov) return fail(ctx, PCR_ERR_STATE, "NDT: voxels with inverse covariance not built");
            break;
        default:
            return fail(ctx, PCR_ERR_ARG, "unknown method id " + std::to_string(method));
    }
    if (!ctx->scan_set) return fail(ctx, PCR_ERR_STATE, "scan not set");
    return PCR_OK;
}

static void fill_params(pcr_ctx* ctx, int method, double max_dist, LinParams& P) {
    P.sx = ctx->scan_x.as<float>(); P.sy = ctx->scan_y.as<float>(); P.sz = ctx->scan_z.as<float>();
    P.n_pad = ctx->n_scan_pad;
    P.grid = (method == PCR_ICP || method == PCR_PLANE) ? ctx->tgt_grid.view : ctx->vox_grid.view;
    P.pn = ctx->tgt_pn.as<float4>();
    P.prev = ctx->scan_prev.as<int>();
    P.lists = ctx->vox_lists;
    P.use_lists = (method == PCR_VPLANE || method == PCR_NDT) && ctx->use_voxel_lists && ctx->vox_lists.bricks != nullptr;
    P.shell = (method == PCR_ICP || method == PCR_PLANE) ? ctx->tgt_shell : ctx->vox_shell;
    P.use_shell = ((method == PCR_ICP || method == PCR_PLANE) ? ctx->use_shell_lists : ctx->use_voxel_lists) && P.shell.bricks != nullptr;
    P.ball_first = ctx->ball_first;
    // rows per fetch: about 16 fetches per resident warp (one device-wide counter serves them all;
    // a fetch per row would make it the bottleneck of a 10M-point scan), at least 1
    {
        const long long rows = ctx->n_scan_pad / 32, warps = (long long)ctx->sm_count * 4 * (kLinThreads / 32);
        long long per = ctx->grab_rows > 0 ? ctx->grab_rows : rows / (warps * 16);
        P.grab_rows = (int)(per < 1 ? 1 : (per > 64 ? 64 : per));
    }
    P.vrec = method == PCR_NDT ? ctx->vox_rec_ndt.as<float4>() : ctx->vox_rec_plane.as<float4>();
    const float md = (float)max_dist;
    P.max_d2 = md * md;
    P.st = ctx->state.as<LoopState>();
    P.partials = ctx->partials.as<double>();
    P.out_mapped = ctx->d_out_mapped;
}

template <typename K>
static int blocks_for_kernel(pcr_ctx* ctx, K kernel, int& cached, long long n_pad) {
    if (cached == 0) cached = blocks_per_sm(kernel);
    return lin_grid_blocks(ctx, n_pad, cached);
}

template <int METHOD>
static int launch_accumulate(pcr_ctx* ctx, const LinParams& P) {
    int* cache = ctx->lin_blocks_per_sm[METHOD];
    // quads of slots per thread: the grid only needs a quarter of the threads
    {
        const int blocks = blocks_for_kernel(ctx, accumulate_kernel<METHOD, 2>, cache[1], (P.n_pad + 3) / 4);
        accumulate_kernel<METHOD, 2><<<blocks, kLinThreads, 0, ctx->stream>>>(P);
    }
    PCR_LAUNCH_CHECK();
    return PCR_OK;
}

template <int METHOD>
static int launch_method(pcr_ctx* ctx, const LinParams& P) {
    // ctx->min_blocks = resident blocks per SM requested for the correspondence pass (2..6)
    // resident blocks per SM of the correspondence kernel: measured best 6 (40 registers) for the
    // shell-list stream on B-01 (round 2), 4 for the voxel candidate lists; PCR_MIN_BLOCKS overrides
    int mb = ctx->min_blocks > 0 ? ctx->min_blocks : ((METHOD == PCR_METHOD_ICP || METHOD == PCR_METHOD_PLANE) ? 6 : 4);
    mb = mb < 3 ? 3 : (mb > 6 ? 6 : mb);
    int* cache = ctx->lin_blocks_per_sm[METHOD];
    if (ctx->split_passes) {
        int blocks;
        switch (mb) {
            case 3: blocks = blocks_for_kernel(ctx, correspond_kernel<METHOD, 3>, cache[3], P.n_pad); correspond_kernel<METHOD, 3><<<blocks, kLinThreads, 0, ctx->stream>>>(P); break;
            case 4: blocks = blocks_for_kernel(ctx, correspond_kernel<METHOD, 4>, cache[4], P.n_pad); correspond_kernel<METHOD, 4><<<blocks, kLinThreads, 0, ctx->stream>>>(P); break;
            case 5: blocks = blocks_for_kernel(ctx, correspond_kernel<METHOD, 5>, cache[5], P.n_pad); correspond_kernel<METHOD, 5><<<blocks, kLinThreads, 0, ctx->stream>>>(P); break;
            default: blocks = blocks_for_kernel(ctx, correspond_kernel<METHOD, 6>, cache[6], P.n_pad); correspond_kernel<METHOD, 6><<<blocks, kLinThreads, 0, ctx->stream>>>(P); break;
        }
        PCR_LAUNCH_CHECK();
        return launch_accumulate<METHOD>(ctx, P);
    }
    const int blocks = blocks_for_kernel(ctx, linearize_fused_kernel<METHOD, 3>, cache[7], P.n_pad);
    linearize_fused_kernel<METHOD, 3><<<blocks, kLinThreads, 0, ctx->stream>>>(P);
    PCR_LAUNCH_CHECK();
    return PCR_OK;
}

// ---- tile-stream path -------------------------------------------------------------------------
template <int METHOD, int MINB, int KR, int NG>
static int launch_tile_kernel(pcr_ctx* ctx, const LinParams& P, const TileParams& TP, int& cached_blocks) {
    const size_t smem = (size_t)TP.warp_bytes * (kLinThreads / 32);
    auto kernel = tile_linearize_kernel<METHOD, MINB, KR, NG>;
    if (cached_blocks == 0) {
        PCR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kLinThreads, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
        cached_blocks = nb;
    }
    const long long rows = P.n_pad / 32;
    long long want = (rows + kLinThreads / 32 - 1) / (kLinThreads / 32);
    long long cap = (long long)ctx->sm_count * cached_blocks;
    if (cap > kMaxLinBlocks) cap = kMaxLinBlocks;
    if (want < 1) want = 1;
    const int blocks = (int)(want < cap ? want : cap);
    kernel<<<blocks, kLinThreads, smem, ctx->stream>>>(P, TP);
    PCR_LAUNCH_CHECK();
    return PCR_OK;
}

template <int MINB, int NG>
static int launch_tile_correspond(pcr_ctx* ctx, const LinParams& P, const TileParams& TP, int& cached_blocks) {
    const size_t smem = (size_t)TP.warp_bytes * (kLinThreads / 32);
    auto kernel = tile_correspond_kernel<MINB, NG>;
    if (cached_blocks == 0) {
        PCR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kLinThreads, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
        cached_blocks = nb;
    }
    const long long rows = P.n_pad / 32;
    long long want = (rows + kLinThreads / 32 - 1) / (kLinThreads / 32);
    long long cap = (long long)ctx->sm_count * cached_blocks;
    if (cap > kMaxLinBlocks) cap = kMaxLinBlocks;
    if (want < 1) want = 1;
    kernel<<<(int)(want < cap ? want : cap), kLinThreads, smem, ctx->stream>>>(P, TP);
    PCR_LAUNCH_CHECK();
    return PCR_OK;
}

template <int METHOD, int KR, int NG>
static int launch_tile_minb(pcr_ctx* ctx, const LinParams& P, const TileParams& TP, int& cached_blocks) {
    return launch_tile_kernel<METHOD, 3, KR, NG>(ctx, P, TP, cached_blocks);
}

template <int METHOD>
static int launch_tile(pcr_ctx* ctx, const LinParams& P) {
    const bool vox = METHOD == PCR_METHOD_VPLANE || METHOD == PCR_METHOD_NDT;
    const TileIndex& t = vox ? ctx->tile_vox : ctx->tile_tgt;
    TileParams TP{};
    TP.G = t.view;
    TP.pay = METHOD == PCR_METHOD_ICP ? nullptr : (METHOD == PCR_METHOD_NDT ? t.pay2.as<float4>() : t.pay.as<float4>());
    TP.hint = ctx->scan_hint.as<float>();
    TP.perm = t.perm.as<uint32_t>();
    TP.match_out = ctx->record_matches ? ctx->scan_prev.as<int>() : nullptr;
    TP.cap = ctx->tile_cap;
    TP.bulk_min = ctx->tile_bulk_min;
    TP.rmax = tile_rmax(t.view, P.max_d2);
    TP.core_e = ctx->tile_core_e;
    constexpr int KRMAX = 4;                       // match slots reserved per warp: [KRMAX][32] float4
    TP.warp_bytes = (int)(((size_t)TP.cap * 16 + 16 + (size_t)KRMAX * 32 * 16 + 127) / 128 * 128);
    int* cache = ctx->lin_blocks_per_sm[METHOD];
    if (ctx->tile_split) {
        // correspondences (method independent), then the accumulate kernel of the list path on the parked positions
        TP.match_out = ctx->scan_prev.as<int>();
        TP.warp_bytes = (int)(((size_t)TP.cap * 16 + 16 + 127) / 128 * 128);
        int rc;
        const int mb = ctx->tile_min_blocks > 0 ? ctx->tile_min_blocks : 5;
        if (ctx->tile_groups == 4) rc = mb >= 6 ? launch_tile_correspond<6, 4>(ctx, P, TP, cache[11]) : (mb == 5 ? launch_tile_correspond<5, 4>(ctx, P, TP, cache[11]) : launch_tile_correspond<4, 4>(ctx, P, TP, cache[11]));
        else rc = mb >= 6 ? launch_tile_correspond<6, 1>(ctx, P, TP, cache[11]) : (mb == 5 ? launch_tile_correspond<5, 1>(ctx, P, TP, cache[11]) : launch_tile_correspond<4, 1>(ctx, P, TP, cache[11]));
        if (rc) return rc;
        return launch_accumulate<METHOD>(ctx, P);
    }
    // rows per unit of work: 4 when every warp still gets many units, 2 for small scans (less tail)
    const long long rows = P.n_pad / 32;
    const bool big = ctx->tile_rows_per_unit > 0 ? ctx->tile_rows_per_unit >= 4 : rows >= (long long)ctx->sm_count * 32 * 4 * 8;
    const int ng = ctx->tile_groups;
    if (big) {
        switch (ng) {
            case 1: return launch_tile_minb<METHOD, 4, 1>(ctx, P, TP, cache[8]);
            case 2: return launch_tile_minb<METHOD, 4, 2>(ctx, P, TP, cache[8]);
            default: return launch_tile_minb<METHOD, 4, 4>(ctx, P, TP, cache[8]);
        }
    }
    switch (ng) {
        case 1: return launch_tile_minb<METHOD, 2, 1>(ctx, P, TP, cache[9]);
        case 2: return launch_tile_minb<METHOD, 2, 2>(ctx, P, TP, cache[9]);
        default: return launch_tile_minb<METHOD, 2, 4>(ctx, P, TP, cache[9]);
    }
}

static int launch_linearize(pcr_ctx* ctx, int method, LinParams& P) {
    if (ctx->use_tile) {
        int rc = ensure_tile_index(ctx, method);
        if (rc) return rc;
        ctx->prev_which = (ctx->record_matches || ctx->tile_split) ? ((method == PCR_ICP || method == PCR_PLANE) ? 0 : 1) : -1;
        switch (method) {
            case PCR_ICP: return launch_tile<PCR_METHOD_ICP>(ctx, P);
            case PCR_PLANE: return launch_tile<PCR_METHOD_PLANE>(ctx, P);
            case PCR_VPLANE: return launch_tile<PCR_METHOD_VPLANE>(ctx, P);
            default: return launch_tile<PCR_METHOD_NDT>(ctx, P);
        }
    }
    // target-point methods stream shell lists: build them now if the caller did not (C-ABI users
    // that skipped pcr_build_correspondence_lists still get the fast path, one call late)
    if ((method == PCR_ICP || method == PCR_PLANE) && !ctx->shell_tried) {
        int rc = pcr_build_correspondence_lists(ctx);
        if (rc) return rc;
        P.shell = ctx->tgt_shell;
        P.use_shell = ctx->use_shell_lists && ctx->tgt_shell.bricks != nullptr;
    }
    // warm-start positions refer to one particular index: forget them when it changed
    const int which = (method == PCR_ICP || method == PCR_PLANE) ? 0 : 1;
    const long long epoch = which == 0 ? ctx->tgt_grid_epoch : ctx->vox_grid_epoch;
    if (ctx->prev_which != which || ctx->prev_epoch != epoch) {
        PCR_CUDA(cudaMemsetAsync(ctx->scan_prev.p, 0xFF, (size_t)ctx->n_scan_pad * 4, ctx->stream));
        ctx->prev_which = which;
        ctx->prev_epoch = epoch;
    }
    switch (method) {
        case PCR_ICP: return launch_method<PCR_METHOD_ICP>(ctx, P);
        case PCR_PLANE: return launch_method<PCR_METHOD_PLANE>(ctx, P);
        case PCR_VPLANE: return launch_method<PCR_METHOD_VPLANE>(ctx, P);
        default: return launch_method<PCR_METHOD_NDT>(ctx, P);
    }
}

static int allreduce_record(pcr_ctx* ctx) {
    double* rec = ctx->state.as<LoopState>()->rec;
    int rc = g_nccl.AllReduce(rec, rec, PCR_NEQ, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, PCR_ERR_NCCL, "ncclAllReduce: " + nccl_err(rc));
    return PCR_OK;
}

}  // namespace pcr

using namespace pcr;

extern "C" {

static int bits_for_u64(unsigned long long v) {   // number of bits needed to represent values < v
    int b = 0;
    while (b < 64 && (1ull << b) < v) ++b;
    return b < 1 ? 1 : b;
}

// Order n device points (float[3n]) by the correspondence-grid cell of their posed images (or along a Morton
// curve when no grid exists yet) and lay them out as NaN-padded SoA at sx / sy / sz (n_pad slots); everything on
// the context's stream, scratch in the context's tmp buffers.
// kept != nullptr: where the permutation of these n points is remembered (OrderCache below).  kept_valid: it holds
// the permutation of an earlier upload of the same cloud -- it is used as it is and the key + radix-sort kernels are
// skipped; otherwise the permutation computed here is stored there.
static int order_and_layout(pcr_ctx* ctx, const float* d_xyz, long long n, long long n_pad, int sort, const double* T, int method,
                            float* sx, float* sy, float* sz, uint32_t* kept = nullptr, bool kept_valid = false) {
    const uint32_t* order = nullptr;
    if (kept && kept_valid && n > 1) {
        order = kept;
    } else if (sort > 0 && n > 1) {
        // the grid the correspondences will be searched in, if it exists already
        const Grid* g = nullptr;
        Grid tile_as_grid;                          // the row grid of the tile-stream path, described as a GridView for the key kernel
        if (ctx->use_tile) {
            const TileIndex* t = nullptr;
            if ((method == PCR_VPLANE || method == PCR_NDT) && ctx->tile_vox.built && ctx->tile_vox.view.n > 0) t = &ctx->tile_vox;
            else if ((method == PCR_ICP || method == PCR_PLANE || method < 0) && ctx->tile_tgt.built) t = &ctx->tile_tgt;
            else if (method < 0 && ctx->tile_vox.built && ctx->tile_vox.view.n > 0) t = &ctx->tile_vox;
            if (t) {
                GridView& V = tile_as_grid.view;
                V.ox = t->view.ox; V.oy = t->view.oy; V.oz = t->view.oz;
                V.h = t->view.c; V.inv_h = t->view.inv_c;
                V.cnx = t->view.nx; V.cny = t->view.ny; V.cnz = t->view.nz;
                V.bnx = (V.cnx + 3) / 4; V.bny = (V.cny + 3) / 4; V.bnz = (V.cnz + 3) / 4;
                g = &tile_as_grid;
            }
        }
        if (g) {
        } else if ((method == PCR_VPLANE || method == PCR_NDT) && ctx->vox_grid.built && ctx->vox_grid.view.n_pts > 0) g = &ctx->vox_grid;
        else if ((method == PCR_ICP || method == PCR_PLANE || method < 0) && ctx->tgt_grid.built) g = &ctx->tgt_grid;
        else if (method < 0 && ctx->vox_grid.built && ctx->vox_grid.view.n_pts > 0) g = &ctx->vox_grid;
        if (g && ctx->cell_order) {
            // order by the grid cell of the posed point (see cell_key_kernel)
            const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
            Pose32 pose;
            pose32_from_T(T ? T : I, pose);
            PCR_CUDA(ctx->tmp_a.ensure((size_t)n * 8));
            PCR_CUDA(ctx->tmp_b.ensure((size_t)n * 8));
            PCR_CUDA(ctx->tmp_d.ensure((size_t)n * 4 * 2));
            uint32_t* v_in = ctx->tmp_d.as<uint32_t>(); uint32_t* v_out = v_in + n;
            const unsigned long long nkeys = (unsigned long long)g->view.bnx * g->view.bny * g->view.bnz * 64ull;
            const int end_bit = bits_for_u64(nkeys);
            size_t tmp = 0;
            if (end_bit <= 32) {                                  // the usual case: 32-bit keys sort faster
                uint32_t* k_in = ctx->tmp_a.as<uint32_t>();
                uint32_t* k_out = ctx->tmp_b.as<uint32_t>();
                cell_key_kernel<uint32_t><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_xyz, n, g->view, pose, k_in, v_in);
                PCR_LAUNCH_CHECK();
                PCR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (long long)n, 0, end_bit, ctx->stream));
                PCR_CUDA(ctx->cub_tmp.ensure(tmp));
                PCR_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, k_in, k_out, v_in, v_out, (long long)n, 0, end_bit, ctx->stream));
            } else {
                unsigned long long* k_in = ctx->tmp_a.as<unsigned long long>();
                unsigned long long* k_out = ctx->tmp_b.as<unsigned long long>();
                cell_key_kernel<unsigned long long><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_xyz, n, g->view, pose, k_in, v_in);
                PCR_LAUNCH_CHECK();
                PCR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (long long)n, 0, end_bit, ctx->stream));
                PCR_CUDA(ctx->cub_tmp.ensure(tmp));
                PCR_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, k_in, k_out, v_in, v_out, (long long)n, 0, end_bit, ctx->stream));
            }
            ctx->launches += 4;
            order = v_out;
        } else {
            // no grid yet: Morton order in the scan's own bounding box
            PCR_CUDA(ctx->tmp_e.ensure(64));
            int* d_mm = ctx->tmp_e.as<int>();
            mm_init_kernel<<<1, 32, 0, ctx->stream>>>(d_mm);
            PCR_LAUNCH_CHECK();
            int nb = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
            scan_bbox_kernel<<<nb, 256, 0, ctx->stream>>>(d_xyz, n, d_mm);
            PCR_LAUNCH_CHECK();
            PCR_CUDA(ctx->tmp_c.ensure((size_t)n * 4 * 2));
            PCR_CUDA(ctx->tmp_d.ensure((size_t)n * 4 * 2));
            uint32_t* k_in = ctx->tmp_c.as<uint32_t>(); uint32_t* k_out = k_in + n;
            uint32_t* v_in = ctx->tmp_d.as<uint32_t>(); uint32_t* v_out = v_in + n;
            morton_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_xyz, n, d_mm, k_in, v_in);
            PCR_LAUNCH_CHECK();
            size_t tmp = 0;
            PCR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (long long)n, 0, 30, ctx->stream));
            PCR_CUDA(ctx->cub_tmp.ensure(tmp));
            PCR_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, k_in, k_out, v_in, v_out, (long long)n, 0, 30, ctx->stream));
            ctx->launches += 4;
            order = v_out;
        }
        if (kept) PCR_CUDA(cudaMemcpyAsync(kept, order, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    scan_to_soa_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, ctx->stream>>>(d_xyz, order, n, n_pad,
                                                                                         sx, sy, sz);
    PCR_LAUNCH_CHECK();
    return PCR_OK;
}

// The remembered scan order (pcr_set_scan, sort == 2).  A Gauss-Newton loop on the host calls calc_H_g_e2 with one
// array every iteration: the cell order of its first upload stays coherent under the small rigid motions in between,
// and ANY order gives the same correspondences -- so the permutation (one per chunk when the upload is pipelined in
// `chunks` pieces of `chunk_len` points) is kept while (size, partition, method, target structures) stay the same.
struct OrderCache {
    uint32_t* slot = nullptr;    // device memory for the permutation(s); nullptr: nothing is kept
    bool valid = false;          // it holds the order of the previous upload of this cloud
    long long n, chunk_len, epoch; int chunks, method;
};
static int order_cache_open(pcr_ctx* ctx, long long n, long long n_slots, int chunks, long long chunk_len, int sort, int method, OrderCache& oc) {
    oc.n = n; oc.chunks = chunks; oc.chunk_len = chunk_len; oc.method = method;
    oc.epoch = ctx->tgt_grid_epoch * 1000003ll + ctx->vox_grid_epoch;
    if (!ctx->order_reuse || sort <= 0 || n <= 1) return PCR_OK;
    oc.valid = sort == 2 && ctx->scan_order.p != nullptr && ctx->order_n == n && ctx->order_chunks == chunks &&
               ctx->order_chunk_len == chunk_len && ctx->order_method == method && ctx->order_epoch == oc.epoch;
    if (!oc.valid) ctx->order_n = -1;                       // about to be overwritten: void until order_cache_close
    PCR_CUDA(ctx->scan_order.ensure((size_t)n_slots * 4));  // (same n: same size, a valid content is never reallocated away)
    oc.slot = ctx->scan_order.as<uint32_t>();
    return PCR_OK;
}
static void order_cache_close(pcr_ctx* ctx, const OrderCache& oc) {
    if (!oc.slot || oc.valid) return;
    ctx->order_n = oc.n; ctx->order_chunks = oc.chunks; ctx->order_chunk_len = oc.chunk_len;
    ctx->order_method = oc.method; ctx->order_epoch = oc.epoch;
}

static int set_scan_impl(pcr_ctx* ctx, const float* xyz, int64_t n, int sort, const double* T, int method) {
    if (!ctx) return PCR_ERR_ARG;
    if (n < 0 || (n > 0 && !xyz)) return fail(ctx, PCR_ERR_ARG, "pcr_set_scan: bad arguments");
    if (n >= (1ll << 31)) return fail(ctx, PCR_ERR_LIMIT, "pcr_set_scan: point count exceeds 2^31-1");
    PCR_CUDA(cudaSetDevice(ctx->device));
    ctx->n_scan = n;
    ctx->n_scan_pad = (n + 31) / 32 * 32;
    ctx->scan_set = true;
    ctx->scan_sorted = false;

    if (n == 0) return PCR_OK;
    const float* d_xyz;
    const bool from_host = !is_device_pointer(xyz);
    if (!from_host) {
        d_xyz = xyz;
    } else {
        PCR_CUDA(ctx->scan_raw.ensure((size_t)n * 12));
        PCR_CUDA(cudaMemcpyAsync(ctx->scan_raw.p, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        PCR_CUDA(cudaEventRecord(ctx->ev_copy, ctx->stream));
        d_xyz = ctx->scan_raw.as<float>();
    }
    const size_t pad_bytes = (size_t)ctx->n_scan_pad * 4;
    PCR_CUDA(ctx->scan_x.ensure(pad_bytes));
    PCR_CUDA(ctx->scan_y.ensure(pad_bytes));
    PCR_CUDA(ctx->scan_z.ensure(pad_bytes));
    PCR_CUDA(ctx->scan_prev.ensure(pad_bytes));
    ctx->prev_which = -1;                       // new scan: parked positions are void
    PCR_CUDA(ctx->scan_hint.ensure((size_t)ctx->n_scan_pad / 32 * 4));
    fill_float_kernel<<<(unsigned)((ctx->n_scan_pad / 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->scan_hint.as<float>(), ctx->n_scan_pad / 32, ctx->tile_first_radius);
    PCR_LAUNCH_CHECK();
    {
        OrderCache oc;
        int rc = order_cache_open(ctx, n, n, 1, n, sort, method, oc);
        if (rc) return rc;
        rc = order_and_layout(ctx, d_xyz, n, ctx->n_scan_pad, sort, T, method, ctx->scan_x.as<float>(), ctx->scan_y.as<float>(),
                              ctx->scan_z.as<float>(), oc.slot, oc.valid);
        if (rc) return rc;
        order_cache_close(ctx, oc);
    }
    ctx->scan_sorted = sort != 0 && n > 1;
    // Host source: return as soon as the caller's buffer has been read; sort and re-layout keep
    // running on the stream (every consumer is stream-ordered behind them).  Device source: the
    // caller's array must stay untouched until the re-layout has read it, so wait for the stream.
    if (from_host) PCR_CUDA(cudaEventSynchronize(ctx->ev_copy));
    else PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    return PCR_OK;
}

int pcr_set_scan(pcr_ctx* ctx, const float* xyz, int64_t n, int sort) { return set_scan_impl(ctx, xyz, n, sort, nullptr, -1); }

int pcr_set_scan_posed(pcr_ctx* ctx, const float* xyz, int64_t n, int sort, const double T[16], int method) {
    return set_scan_impl(ctx, xyz, n, sort, T, method);
}

int pcr_linearize_async(pcr_ctx* ctx, int method, const double T[16], double max_dist, int reps) {
    if (!ctx || !T) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    int rc = check_method(ctx, method);
    if (rc) return rc;
    LinParams P{};
    fill_params(ctx, method, max_dist, P);
    memcpy(P.T_param, T, sizeof(double) * 16);
    P.use_param_T = 1; P.device_loop = 0; P.max_iter = 0; P.tol = 0.0;
    for (int r = 0; r < reps; ++r) {
        rc = launch_linearize(ctx, method, P);
        if (rc) return rc;
        if (ctx->nccl_comm) { rc = allreduce_record(ctx); if (rc) return rc; }
    }
    return PCR_OK;
}

int pcr_linearize(pcr_ctx* ctx, int method, const double T[16], double max_dist, double out[PCR_RECORD_LEN]) {
    if (!ctx || !T || !out) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    PCR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = pcr_linearize_async(ctx, method, T, max_dist, 1);
    if (rc) return rc;
    PCR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (ctx->nccl_comm) {
        PCR_CUDA(cudaMemcpyAsync(ctx->h_state->rec, ctx->state.as<LoopState>()->rec, PCR_NEQ * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(out, ctx->h_state->rec, PCR_NEQ * sizeof(double));
    } else {
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(out, ctx->h_out, PCR_NEQ * sizeof(double));
    }
    PCR_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return PCR_OK;
}

// One linearisation of a scan that lives in HOST memory, as ONE call: the scan is cut into chunks; chunk k+1
// crosses PCIe on a copy stream while chunk k is ordered, searched and accumulated on the compute stream, each
// chunk leaving its own 29-double record in mapped pinned memory; the records are summed on the host (they are
// sums over scan points).  End to end this path is bound by the host->device copy (14.3 MB at C2: 0.28 ms at PCIe
// 5 speed against 0.19 ms of kernels), so hiding everything else behind the copy is what counts.
int pcr_linearize_host(pcr_ctx* ctx, int method, const double T[16], double max_dist, const float* xyz, int64_t n, int sort,
                       double out[PCR_RECORD_LEN]) {
    if (!ctx || !T || !out) return PCR_ERR_ARG;
    if (n < 0 || (n > 0 && !xyz)) return fail(ctx, PCR_ERR_ARG, "pcr_linearize_host: bad arguments");
    PCR_CUDA(cudaSetDevice(ctx->device));
    // chunks of at least 2M points: below that the dozen launches per chunk (ordering + kernels) cost more than the
    // overlap saves -- measured at C2 (1.19M points): 1599 it/s in one piece, 1410 / 1030 / 591 in 2 / 4 / 8 chunks; at C3
    // (10M points) 311 -> 376 it/s with 4 chunks (profiles/r2_notes.md)
    int K = (int)std::min<long long>(ctx->host_chunks, n / 2000000);
    // sort == 2 (the same cloud again, see OrderCache): the orders of the chunks are computed once and kept, a chunk then
    // costs three launches.  Smaller chunks were tried for this case and lost (C2 in 4 chunks of 300k points: 1387 it/s
    // against 1815 in one piece), so the threshold is the same 2M points unless PCR_E2E_MIN_CHUNK says otherwise
    if (sort == 2 && ctx->order_reuse) K = std::max(K, (int)std::min<long long>(ctx->host_chunks, n / ctx->host_chunk_min_reuse));
    // small scans, device-resident scans, multi-GPU contexts and the tile-stream path take the two-step route
    if (K <= 1 || ctx->nccl_comm || ctx->use_tile || is_device_pointer(xyz)) {
        int rc = set_scan_impl(ctx, xyz, n, sort, T, method);
        if (rc) return rc;
        return pcr_linearize(ctx, method, T, max_dist, out);
    }
    if (n >= (1ll << 31)) return fail(ctx, PCR_ERR_LIMIT, "pcr_linearize_host: point count exceeds 2^31-1");
    ctx->scan_set = true;                                    // (check_method wants a scan; the chunks below become the resident scan)
    int rc = check_method(ctx, method);
    if (rc) { ctx->scan_set = false; return rc; }
    if (!ctx->copy_stream) PCR_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < K; ++k)
        if (!ctx->ev_chunk[k]) PCR_CUDA(cudaEventCreateWithFlags(&ctx->ev_chunk[k], cudaEventDisableTiming));
    const long long chunk = ((n + K - 1) / K + 31) / 32 * 32;              // scan points per chunk (the last one may be shorter)
    const long long n_pad = chunk * (K - 1) + ((n - chunk * (K - 1)) + 31) / 32 * 32;
    ctx->n_scan = n;
    ctx->n_scan_pad = n_pad;
    ctx->scan_sorted = sort != 0;
    PCR_CUDA(ctx->scan_raw.ensure((size_t)n * 12));
    const size_t pad_bytes = (size_t)n_pad * 4;
    PCR_CUDA(ctx->scan_x.ensure(pad_bytes));
    PCR_CUDA(ctx->scan_y.ensure(pad_bytes));
    PCR_CUDA(ctx->scan_z.ensure(pad_bytes));
    PCR_CUDA(ctx->scan_prev.ensure(pad_bytes));
    PCR_CUDA(ctx->scan_hint.ensure((size_t)n_pad / 32 * 4));
    PCR_CUDA(cudaMemsetAsync(ctx->scan_prev.p, 0xFF, pad_bytes, ctx->stream));
    ctx->prev_which = (method == PCR_ICP || method == PCR_PLANE) ? 0 : 1;
    ctx->prev_epoch = ctx->prev_which == 0 ? ctx->tgt_grid_epoch : ctx->vox_grid_epoch;
    if ((method == PCR_ICP || method == PCR_PLANE) && !ctx->shell_tried) {
        rc = pcr_build_correspondence_lists(ctx);
        if (rc) return rc;
    }
    PCR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    // the copy stream must not overwrite the staging buffer while an earlier call's re-layout still reads it
    PCR_CUDA(cudaEventRecord(ctx->ev_copy, ctx->stream));
    PCR_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    for (int k = 0; k < K; ++k) {
        const long long off = chunk * k, cn = std::min<long long>(chunk, n - off);
        PCR_CUDA(cudaMemcpyAsync(ctx->scan_raw.as<float>() + 3 * off, xyz + 3 * off, (size_t)cn * 12, cudaMemcpyHostToDevice, ctx->copy_stream));
        PCR_CUDA(cudaEventRecord(ctx->ev_chunk[k], ctx->copy_stream));
    }
    LinParams P{};
    fill_params(ctx, method, max_dist, P);
    memcpy(P.T_param, T, sizeof(double) * 16);
    P.use_param_T = 1; P.device_loop = 0; P.max_iter = 0; P.tol = 0.0;
    OrderCache oc;
    rc = order_cache_open(ctx, n, n_pad, K, chunk, sort, method, oc);
    if (rc) return rc;
    for (int k = 0; k < K; ++k) {
        const long long off = chunk * k, cn = std::min<long long>(chunk, n - off), cpad = (cn + 31) / 32 * 32;
        PCR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[k], 0));
        float* sx = ctx->scan_x.as<float>() + off; float* sy = ctx->scan_y.as<float>() + off; float* sz = ctx->scan_z.as<float>() + off;
        rc = order_and_layout(ctx, ctx->scan_raw.as<float>() + 3 * off, cn, cpad, sort, T, method, sx, sy, sz,
                              oc.slot ? oc.slot + off : nullptr, oc.valid);
        if (rc) return rc;
        LinParams Pk = P;
        Pk.sx = sx; Pk.sy = sy; Pk.sz = sz;
        Pk.prev = ctx->scan_prev.as<int>() + off;
        Pk.n_pad = cpad;
        Pk.out_mapped = ctx->d_out_mapped + 64 * k;
        {
            const long long rows = cpad / 32, warps = (long long)ctx->sm_count * 4 * (kLinThreads / 32);
            long long per = ctx->grab_rows > 0 ? ctx->grab_rows : rows / (warps * 16);
            Pk.grab_rows = (int)(per < 1 ? 1 : (per > 64 ? 64 : per));
        }
        switch (method) {
            case PCR_ICP: rc = launch_method<PCR_METHOD_ICP>(ctx, Pk); break;
            case PCR_PLANE: rc = launch_method<PCR_METHOD_PLANE>(ctx, Pk); break;
            case PCR_VPLANE: rc = launch_method<PCR_METHOD_VPLANE>(ctx, Pk); break;
            default: rc = launch_method<PCR_METHOD_NDT>(ctx, Pk); break;
        }
        if (rc) return rc;
    }
    order_cache_close(ctx, oc);
    PCR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < PCR_NEQ; ++i) {
        double v = 0.0;
        for (int k = 0; k < K; ++k) v += ctx->h_out[64 * k + i];
        out[i] = v;
    }
    PCR_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    // tile-stream halo radii: not used on this path, but keep them defined for a later switch
    fill_float_kernel<<<(unsigned)((n_pad / 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->scan_hint.as<float>(), n_pad / 32, ctx->tile_first_radius);
    PCR_LAUNCH_CHECK();
    return PCR_OK;
}

int pcr_loop_begin(pcr_ctx* ctx, const double T0[16]) {
    if (!ctx || !T0) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    T16 t0;
    memcpy(t0.v, T0, sizeof(t0.v));
    ctx->h_out[49] = 0.0;
    loop_init_kernel<<<1, 32, 0, ctx->stream>>>(ctx->state.as<LoopState>(), t0);
    PCR_LAUNCH_CHECK();
    if (ctx->use_tile && ctx->scan_set && ctx->n_scan_pad > 0) {
        // a new trajectory starts from a new pose: the halo radii remembered from the last one are void
        fill_float_kernel<<<(unsigned)((ctx->n_scan_pad / 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->scan_hint.as<float>(), ctx->n_scan_pad / 32, ctx->tile_first_radius);
        PCR_LAUNCH_CHECK();
    }
    return PCR_OK;
}

int pcr_loop_step_async(pcr_ctx* ctx, int method, int max_iter, double tol, double max_dist, int reps) {
    if (!ctx) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    int rc = check_method(ctx, method);
    if (rc) return rc;
    LoopState* st = ctx->state.as<LoopState>();
    LinParams P{};
    fill_params(ctx, method, max_dist, P);
    P.use_param_T = 0; P.max_iter = max_iter; P.tol = tol;
    const bool multi = ctx->nccl_comm != nullptr;
    P.device_loop = multi ? 0 : 1;
    for (int r = 0; r < reps; ++r) {
        rc = launch_linearize(ctx, method, P);
        if (rc) return rc;
        if (multi) {
            rc = allreduce_record(ctx);
            if (rc) return rc;
            gn_step_kernel<<<1, 32, 0, ctx->stream>>>(st, tol, max_iter);
            PCR_LAUNCH_CHECK();
        }
    }
    return PCR_OK;
}

int pcr_loop_state(pcr_ctx* ctx, double T_out[16], int* iters, int* done, double* e2_trace, int trace_cap) {
    if (!ctx) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    PCR_CUDA(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(LoopState), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (T_out) memcpy(T_out, ctx->h_state->T, sizeof(double) * 16);
    if (iters) *iters = ctx->h_state->iter;
    if (done) *done = ctx->h_state->done;
    if (e2_trace && trace_cap > 0)
        memcpy(e2_trace, ctx->h_state->e2_trace, sizeof(double) * std::min(std::min(ctx->h_state->iter, trace_cap), kMaxTrace));
    return PCR_OK;
}

int pcr_align(pcr_ctx* ctx, int method, const double T0[16], int max_iter, double tol, double max_dist, double T_out[16],
              int* iters, double* e2_trace) {
    if (!ctx || !T0 || !T_out) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    int rc = check_method(ctx, method);
    if (rc) return rc;
    if (max_iter < 0) return fail(ctx, PCR_ERR_ARG, "pcr_align: max_iter < 0");
    LoopState* st = ctx->state.as<LoopState>();
    PCR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = pcr_loop_begin(ctx, T0);
    if (rc) return rc;
    const bool multi = ctx->nccl_comm != nullptr;
    int launched = 0;
    const int chunk = 4;
    int done = max_iter == 0 ? 3 : 0;
    while (!done && launched < max_iter) {
        const int todo = std::min(chunk, max_iter - launched);
        rc = pcr_loop_step_async(ctx, method, max_iter, tol, max_dist, todo);
        if (rc) return rc;
        launched += todo;
        if (multi) {
            PCR_CUDA(cudaMemcpyAsync(&ctx->h_state->done, &st->done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            PCR_CUDA(cudaStreamSynchronize(ctx->stream));
            done = ctx->h_state->done;
        } else {
            PCR_CUDA(cudaStreamSynchronize(ctx->stream));
            done = (int)ctx->h_out[49];
        }
    }
    PCR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    int n_it = 0, dn = 0;
    rc = pcr_loop_state(ctx, T_out, &n_it, &dn, e2_trace, max_iter);
    if (rc) return rc;
    PCR_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    if (iters) *iters = n_it;
    if (dn == 2) return fail(ctx, PCR_ERR_SINGULAR, "pcr_align: singular normal equations (no inliers?)");
    return PCR_OK;
}

int pcr_debug_matches(pcr_ctx* ctx, int which, int64_t* idx) {
    if (!ctx || !idx) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    const Grid& g = which == 0 ? ctx->tgt_grid : ctx->vox_grid;
    if (!g.built) return fail(ctx, PCR_ERR_STATE, "pcr_debug_matches: index not built");
    if (!ctx->scan_set || ctx->n_scan == 0) return fail(ctx, PCR_ERR_STATE, "pcr_debug_matches: scan not set");
    if (ctx->prev_which != which) return fail(ctx, PCR_ERR_STATE, "pcr_debug_matches: no linearisation against this index yet");
    DevBuf di;
    PCR_CUDA(di.ensure((size_t)ctx->n_scan * 8));
    matches_kernel<<<(unsigned)((ctx->n_scan + 255) / 256), 256, 0, ctx->stream>>>(ctx->scan_prev.as<int>(), g.view.pts, ctx->n_scan, di.as<long long>());
    PCR_LAUNCH_CHECK();
    PCR_CUDA(cudaMemcpyAsync(idx, di.p, (size_t)ctx->n_scan * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    di.release();
    return PCR_OK;
}

int pcr_export_gn_rows(pcr_ctx* ctx, int method, const double T[16], double max_dist, double* rows) {
    if (!ctx || !T || !rows) return PCR_ERR_ARG;
    if (method != PCR_PLANE && method != PCR_VPLANE)
        return fail(ctx, PCR_ERR_ARG, "pcr_export_gn_rows: scalar-residual methods only (PCR_PLANE, PCR_VPLANE)");
    PCR_CUDA(cudaSetDevice(ctx->device));
    int rc = check_method(ctx, method);
    if (rc) return rc;
    if (ctx->n_scan == 0) return PCR_OK;
    const int saved = ctx->record_matches;
    ctx->record_matches = 1;                                  // the tile-stream path parks positions only on request
    rc = pcr_linearize_async(ctx, method, T, max_dist, 1);
    ctx->record_matches = saved;
    if (rc) return rc;
    LinParams P{};
    fill_params(ctx, method, max_dist, P);
    memcpy(P.T_param, T, sizeof(double) * 16);
    P.use_param_T = 1;
    const long long n = ctx->n_scan;
    const bool to_device = is_device_pointer(rows);
    DevBuf tmp;
    double* d_rows = rows;
    if (!to_device) {
        PCR_CUDA(tmp.ensure((size_t)n * 28 * sizeof(double)));
        d_rows = tmp.as<double>();
    }
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (method == PCR_PLANE) gn_rows_kernel<PCR_METHOD_PLANE><<<blocks, 256, 0, ctx->stream>>>(P, d_rows, n);
    else gn_rows_kernel<PCR_METHOD_VPLANE><<<blocks, 256, 0, ctx->stream>>>(P, d_rows, n);
    PCR_LAUNCH_CHECK();
    if (!to_device) PCR_CUDA(cudaMemcpyAsync(rows, d_rows, (size_t)n * 28 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    tmp.release();
    return PCR_OK;
}

int pcr_last_kernel_ms(pcr_ctx* ctx, float* ms) {
    if (!ctx || !ms) return PCR_ERR_ARG;
    *ms = ctx->last_ms;
    return PCR_OK;
}

int pcr_comm_unique_id(void* id128) {
    pcr_ctx* ctx = nullptr;
    int rc = load_nccl(ctx);
    if (rc) return rc;
    nccl_uid_t id;
    rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(ctx, PCR_ERR_NCCL, "ncclGetUniqueId: " + nccl_err(rc));
    memcpy(id128, &id, 128);
    return PCR_OK;
}

int pcr_comm_init_rank(pcr_ctx* ctx, int nranks, int rank, const void* id128) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, PCR_ERR_ARG, "pcr_comm_init_rank: bad arguments");
    PCR_CUDA(cudaSetDevice(ctx->device));
    int rc = load_nccl(ctx);
    if (rc) return rc;
    pcr_comm_destroy(ctx);
    nccl_uid_t id;
    memcpy(&id, id128, 128);
    void* comm = nullptr;
    rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (rc != 0) return fail(ctx, PCR_ERR_NCCL, "ncclCommInitRank: " + nccl_err(rc));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return PCR_OK;
}

int pcr_comm_move(pcr_ctx* dst, pcr_ctx* src) {
    if (!dst || !src) return PCR_ERR_ARG;
    if (dst == src) return PCR_OK;
    if (dst->device != src->device) return fail(dst, PCR_ERR_ARG, "pcr_comm_move: contexts live on different devices");
    cudaSetDevice(src->device);
    if (src->stream) cudaStreamSynchronize(src->stream);      // no collective of the old context may still be in flight
    pcr_comm_destroy(dst);
    dst->nccl_comm = src->nccl_comm; dst->nranks = src->nranks; dst->rank = src->rank;
    src->nccl_comm = nullptr; src->nranks = 1; src->rank = 0;
    return PCR_OK;
}

int pcr_sync_producer(pcr_ctx* ctx, int has_stream, void* stream) {
    if (!ctx) return PCR_ERR_ARG;
    PCR_CUDA(cudaSetDevice(ctx->device));
    if (has_stream) {
        // __cuda_array_interface__ v3: 1 = legacy default stream, 2 = per-thread default stream, else a cudaStream_t
        cudaStream_t s = stream == (void*)1 ? cudaStreamLegacy : (stream == (void*)2 ? cudaStreamPerThread : (cudaStream_t)stream);
        PCR_CUDA(cudaStreamSynchronize(s));
    } else {
        PCR_CUDA(cudaDeviceSynchronize());
    }
    return PCR_OK;
}

int pcr_comm_destroy(pcr_ctx* ctx) {
    if (!ctx) return PCR_ERR_ARG;
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy(ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return PCR_OK;
}

}  // extern "C"
