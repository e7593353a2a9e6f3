// Host-side context shared by the translation units of libpcr_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/pcr_b200.h"
#include "pcr_common.cuh"
#include "pcr_tile.cuh"

#ifndef PCR_ACC_BATCH
#define PCR_ACC_BATCH 2
#endif
#ifndef PCR_ACC_MINB
#define PCR_ACC_MINB 2
#endif

namespace pcr {

constexpr int kMaxTrace = 256;          // per-iteration e2 trace capacity of the on-device GN loop
constexpr int kLinThreads = 256;        // threads per block of the linearise kernels
constexpr int kMaxLinBlocks = 148 * 8;  // upper bound on persistent grid size (partials buffer)
constexpr int kAccBatch = PCR_ACC_BATCH;     // scan slots whose gathers are in flight together in the accumulate pass
constexpr int kAccMinBlocks = PCR_ACC_MINB;  // resident blocks per SM requested for the accumulate kernel

// Device-resident Gauss-Newton loop state (one per context).
struct LoopState {
    double T[16];           // current transform, row major
    double rec[PCR_NEQ_PAD];// last reduced normal-equation record (see pcr_common.cuh)
    double dx[6];
    double dx_norm;
    double e2_trace[kMaxTrace];
    int iter;               // linearisations executed so far
    int done;               // 0 running, 1 converged, 2 singular H
    unsigned int ticket;    // block arrival counter of the running linearise kernel
    int next_row;           // next unassigned row of 32 scan slots of the running correspondence pass
};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t want) {
        if (want <= bytes && p) return cudaSuccess;
        release();
        if (want == 0) want = 16;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want; else p = nullptr;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Grid {
    GridView view{};
    DevBuf bricks, cell_start, pts;
    uint32_t n_cells = 0;     // occupied cells
    bool built = false;
    void release() { bricks.release(); cell_start.release(); pts.release(); built = false; n_cells = 0; }
};

// Row grid + row-major point copy the tile-stream kernel stages from (see pcr_tile.cuh).
struct TileIndex {
    TileGrid view{};
    DevBuf cs, pts, perm;     // cell starts, points as pair records, position -> position in the source index
    DevBuf pay, pay2;         // per-point payload in the same order: normals (PlaneICP / VPlaneICP); NDT inverse covariances
    long long n_cells_occupied = 0;
    long long pay_epoch = -1;
    bool built = false;
    void release() { cs.release(); pts.release(); perm.release(); pay.release(); pay2.release(); built = false; pay_epoch = -1; }
};

}  // namespace pcr

struct pcr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int sm_count = 148;

    // ---- target points (ICP / PlaneICP / KDTree facade) ----
    long long n_tgt = 0;
    pcr::DevBuf tgt_xyz;          // float[3n], caller order
    pcr::Grid tgt_grid;           // NN index over target points (payload = caller index)
    pcr::DevBuf tgt_nrm_sorted;   // float4[n], same order as tgt_grid.pts
    pcr::DevBuf tgt_pn;           // float4[2n]: (point, normal) interleaved in that order -- the record PlaneICP gathers
    pcr::DevBuf tgt_nrm_orig;     // float[3n], caller order
    bool has_normals = false;
    pcr::DevBuf shell_bricks, shell_start, shell_pts, shell_margin2;   // per-cell shell lists over the target grid
    pcr::ShellLists tgt_shell{};  // null pointers = not built
    long long n_shell_band = 0, n_shell_entries = 0;
    bool shell_tried = false;     // pcr_build_correspondence_lists ran for the current grid
    int use_shell_lists = 1;
    double shell_dmax_frac = 3.0; // requested list margin in cell edges (<= 3); reduced until the lists fit the caps below
    double shell_max_gib = 96.0;  // memory cap of the lists (a B200 has 180 GB; 100M target points at margin 1.5 need 76 GB)
    double shell_wide_gib = 8.0;  // ... and of lists with a margin above two cells (worth their 2.5x memory for small targets only)
    double shell_dmax_used = 0.0; // margin actually built (0: no lists)

    // ---- voxel statistics (VPlaneICP / NDT / VoxelGrid facade) ----
    long long n_vox = 0;          // kept voxels
    long long n_vox_all = 0;      // occupied voxels before the min_points filter
    double voxel_size = 0.0;
    pcr::DevBuf vox_mean, vox_cov, vox_norm, vox_icov, vox_count;   // double[3n], [9n], [3n], [9n], int64[n]
    pcr::Grid vox_grid;           // NN index over kept voxel means (payload = voxel ordinal)
    pcr::DevBuf vox_lbricks, vox_list_start, vox_list_idx, vox_list_pts;   // per-cell candidate lists over the voxel means
    pcr::CandLists vox_lists{};   // null pointers = not built
    pcr::DevBuf vshell_bricks, vshell_start, vshell_pts, vshell_margin2;   // margin-ordered shell lists over the kept voxel means
    pcr::ShellLists vox_shell{};  // null pointers = not built
    long long n_vshell_band = 0, n_vshell_entries = 0;
    double vshell_dmax_used = 0.0;
    long long n_band_cells = 0, n_list_entries = 0;
    int use_voxel_lists = 1;
    int list_dilate = 3, list_radius = 5;   // candidate-list band dilation / build neighbourhood radius (cells)
    pcr::DevBuf vox_rec_plane;    // float4[2n]: (mean, 0), (normal, 0)
    pcr::DevBuf vox_rec_ndt;      // float4[3n]: (mean, W00), (W01, W02, W11, W12), (W22, 0, 0, 0)
    bool has_voxels = false, has_icov = false;

    // ---- tile-stream path (alternative hot path, PCR_PATH=tile / pcr_set_path; see pcr_tile.cuh) ----
    pcr::TileIndex tile_tgt;      // over the target points (ICP / PlaneICP)
    pcr::TileIndex tile_vox;      // over the kept voxel means (VPlaneICP / NDT)
    int use_tile = 0;             // 0: list kernels (default: faster on every measured workload), 1: tile-stream kernel (PCR_PATH=tile)
    double tile_ppc_tgt = 8.0;    // desired mean points per occupied cell of the row grids
    double tile_ppc_vox = 4.0;
    int tile_cap = 384;           // points a warp can stage at once
    int tile_core_e = 8;          // lanes farther than this many cells from the leader wait for their own pass
    int tile_min_blocks = 0;      // resident blocks per SM requested (0: default)
    int tile_rows_per_unit = 0;   // warp rows per unit of work (2 or 4; 0: chosen from the scan size)
    int acc_min_blocks = 2;       // resident blocks per SM requested for the accumulate kernel (2: 128 registers, 3: 80)
    int tile_bulk_min = 9;        // staging: ranges of at least this many pair records (32 B each) go through the TMA engine
    int tile_split = 0;           // 1: correspondences and accumulation as two kernels (A/B, PCR_TILE_SPLIT)
    int tile_groups = 1;          // independent groups a warp row is searched as (1, 2, 4, 8; see tile_search_row)
    int record_matches = 0;       // 1: the tile kernel also parks the matched positions (pcr_debug_matches)
    long long normals_epoch = 0;
    float tile_first_radius = 0.5f;   // halo radius (cells) a new scan starts with
    pcr::DevBuf scan_hint;        // float[n_pad / 32]: halo radius each warp row needed last time
    pcr::DevBuf tile_scratch;     // histogram / counters of the row-grid build

    // ---- scan ----
    long long n_scan = 0;         // real points
    long long n_scan_pad = 0;     // padded to a multiple of 32 with NaN (whole tiles enter the kernel loop)
    bool scan_set = false;
    bool scan_sorted = false;     // spatially coherent order (Morton-sorted on upload, or promised by the caller)
    double target_ppc = 0.0;      // desired mean points per occupied cell of the target-point grid; 0 = by target size:
                                  // 10 while lists with a 3-cell margin fit (<= 4M points), else 24 (the margin is then
                                  // bought in metres per byte: larger cells reach farther) -- profiles/r2_notes.md
    int min_blocks = 0;           // resident blocks per SM requested for the correspondence pass (3..6; 0: per-method default)
    int ball_first = 1;           // list misses search the ball of max_dist in one pass (0: ring growth, A/B)
    int cell_order = 1;           // scan upload: order by correspondence-grid cell (0: Morton order in the scan's frame)
    int grab_rows = 0;            // rows of 32 scan slots a warp fetches at a time (0: chosen from the scan size)
    int split_passes = 1;         // 1: correspond + accumulate kernels, 0: one fused kernel (A/B)
    int lin_blocks_per_sm[4][12] = {};   // cached occupancy per (method, kernel variant)
    pcr::DevBuf scan_x, scan_y, scan_z;
    pcr::DevBuf scan_prev;        // int[n_pad]: position matched by the previous linearisation (warm start)
    int prev_which = -1;          // index the positions refer to (0 target grid, 1 voxel grid, -1 none)
    long long prev_epoch = -1, tgt_grid_epoch = 0, vox_grid_epoch = 0;
    pcr::DevBuf scan_raw;         // staging float[3n]
    pcr::DevBuf scan_order;       // uint32[n]: permutation of the last ordered scan (reused by uploads with sort = 2)
    long long order_n = -1, order_epoch = -1, order_chunk_len = -1;
    int order_chunks = 0;
    int order_method = -2;
    int order_reuse = 1;          // PCR_ORDER_REUSE=0: recompute the order on every upload

    // ---- reduction / loop state ----
    pcr::DevBuf partials;         // double[kMaxLinBlocks * PCR_NEQ_PAD]
    pcr::DevBuf state;            // LoopState
    pcr::LoopState* h_state = nullptr;   // pinned host mirror
    double* h_out = nullptr;             // pinned + mapped: rec[32] + T[16] + {iter, done} written by the kernel
    double* d_out_mapped = nullptr;      // device alias of h_out
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_copy = nullptr;       // host->device scan copy finished (the caller may reuse its buffer)
    cudaStream_t copy_stream = nullptr;  // pcr_linearize_host: host->device chunks run here while the kernels of the previous chunk run on `stream`
    cudaEvent_t ev_chunk[8] = {};        // ... one "chunk arrived" event per chunk
    int host_chunks = 4;                 // chunks of pcr_linearize_host (PCR_E2E_CHUNKS, 1 = no overlap)
    long long host_chunk_min_reuse = 2000000;  // smallest chunk of a pipelined upload whose orders are kept (PCR_E2E_MIN_CHUNK);
                                               // measured at C2 with 250k: 4 chunks of 300k points 1387 it/s, one piece 1815
    float last_ms = 0.f;
    long long launches = 0;       // kernels launched by this context (all kinds)

    // ---- scratch for builds ----
    pcr::DevBuf tmp_a, tmp_b, tmp_c, tmp_d, tmp_e, cub_tmp;

    // ---- multi-GPU ----
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
};

namespace pcr {

void set_global_error(const std::string& s);

inline int fail(pcr_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    set_global_error(msg);
    return code;
}

#define PCR_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            return pcr::fail(ctx, PCR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e) + \
                                                    " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
        }                                                                                           \
    } while (0)

#define PCR_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            return pcr::fail(ctx, PCR_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e) + \
                                                    " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
        }                                                                                           \
        ctx->launches++;                                                                            \
    } while (0)

inline bool is_device_pointer(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// internal cross-TU entry points
int build_point_grid(pcr_ctx* ctx, const float* d_xyz, long long n, Grid& g, DevBuf* sorted_payload_out);
int ensure_tile_index(pcr_ctx* ctx, int method);   // builds / refreshes the row grid + payload the method needs
int ensure_loop_buffers(pcr_ctx* ctx);

}  // namespace pcr
