// Once-per-target builds and query utilities of libpcr_b200.so (sm_100a):
//   * brick-grid NN index over a point set            (replaces pykdtree KDTree(data), kdtree.py:18-25)
//   * k-NN normals with the reference's float32 moments (estimate_normals.py:27-87)
//   * voxel statistics / inverse covariance / voxel NN index (voxel.py:69-179)
//   * k-NN and nearest-voxel queries, voxel_filter      (kdtree.py:18-25, voxel.py:171-179, 209-241)
// Sorting / prefix sums use CUB (plumbing); every geometric kernel is hand written.
#include <cub/cub.cuh>
#include <cfloat>
#include <climits>
#include <vector>

#include "pcr_context.cuh"
#include "pcr_grid.cuh"
#include "pcr_linalg.cuh"

namespace pcr {

static thread_local std::string g_err;
void set_global_error(const std::string& s) { g_err = s; }
const char* global_error() { return g_err.c_str(); }

struct BrickRec {
    unsigned long long mask;
    uint32_t base;
    uint32_t pad;
};
static_assert(sizeof(BrickRec) == 16, "brick record must be 16 bytes");

__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
static inline float ord2f_host(int i) {
    int j = i >= 0 ? i : i ^ 0x7fffffff;
    float f;
    memcpy(&f, &j, 4);
    return f;
}

// ---------------------------------------------------------------------------------------
// bounding box of finite points: mm[0..2] = min (ordered ints), mm[3..5] = max
// ---------------------------------------------------------------------------------------
__global__ void bbox_kernel(const float* __restrict__ xyz, long long n, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            int a = f2ord(x), b = f2ord(y), c = f2ord(z);
            lo[0] = min(lo[0], a); hi[0] = max(hi[0], a);
            lo[1] = min(lo[1], b); hi[1] = max(hi[1], b);
            lo[2] = min(lo[2], c); hi[2] = max(hi[2], c);
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
        for (int a = 0; a < 3; ++a) {
            atomicMin(&mm[a], lo[a]);
            atomicMax(&mm[3 + a], hi[a]);
        }
    }
}

__global__ void init_minmax_kernel(int* mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 6) mm[threadIdx.x] = INT_MIN;
}

// ---------------------------------------------------------------------------------------
// point grid: keys, heads, fill, gather
// ---------------------------------------------------------------------------------------
__global__ void point_key_kernel(const float* __restrict__ xyz, long long n, GridView G,
                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float big = 1.0e9f;
    float gx = (xyz[3 * i] - G.ox) * G.inv_h, gy = (xyz[3 * i + 1] - G.oy) * G.inv_h, gz = (xyz[3 * i + 2] - G.oz) * G.inv_h;
    if (!(gx == gx)) gx = 0.f;
    if (!(gy == gy)) gy = 0.f;
    if (!(gz == gz)) gz = 0.f;
    int cx = cell_of(fminf(fmaxf(gx, -big), big), G.cnx);
    int cy = cell_of(fminf(fmaxf(gy, -big), big), G.cny);
    int cz = cell_of(fminf(fmaxf(gz, -big), big), G.cnz);
    unsigned long long brick = ((unsigned long long)(cz >> 2) * G.bny + (cy >> 2)) * G.bnx + (cx >> 2);
    keys[i] = brick * 64ull + (unsigned long long)brick_bit(cx, cy, cz);
    vals[i] = (uint32_t)i;
}

__global__ void head_flag_kernel(const unsigned long long* __restrict__ keys, long long n, uint32_t* __restrict__ flags) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// For each segment head: cell_start[ordinal] = position; set the brick's mask bit; the first
// head of a brick records the brick's base ordinal.
__global__ void grid_fill_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ flags,
                                 const uint32_t* __restrict__ ords, long long n, uint32_t n_cells,
                                 BrickRec* __restrict__ bricks, uint32_t* __restrict__ cell_start) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i == 0) cell_start[n_cells] = (uint32_t)n;
    if (i >= n || !flags[i]) return;
    const unsigned long long key = keys[i];
    const uint32_t ord = ords[i];
    cell_start[ord] = (uint32_t)i;
    const unsigned long long brick = key >> 6;
    atomicOr(&bricks[brick].mask, 1ull << (key & 63ull));
    if (i == 0 || (keys[i - 1] >> 6) != brick) bricks[brick].base = ord;
}

__global__ void gather_points_kernel(const float* __restrict__ xyz, const uint32_t* __restrict__ vals, long long n,
                                     float4* __restrict__ pts) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = vals[i];
    pts[i] = make_float4(xyz[3 * (size_t)j], xyz[3 * (size_t)j + 1], xyz[3 * (size_t)j + 2], __uint_as_float(j));
}

// Four sentinel records behind the last indexed point (and four list entries pointing at the
// first of them): candidate loops read in groups of four without clamping; a sentinel is
// infinitely far from every query and can never win.
__global__ void pad_tail_kernel(float4* __restrict__ pts, long long n, uint32_t* __restrict__ list_idx, long long n_entries) {
    if (threadIdx.x < 4) {
        if (pts) pts[n + threadIdx.x] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, __uint_as_float(0xffffffffu));
        if (list_idx) list_idx[n_entries + threadIdx.x] = (uint32_t)n;
    }
}

static inline int blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }

static int bits_for(unsigned long long v) {   // number of bits needed to represent values < v
    int b = 1;
    while (b < 64 && (1ull << b) < v) ++b;
    return b;
}

// Sort (key64, val32) pairs by key (stable LSD radix sort).
static int sort_pairs64(pcr_ctx* ctx, unsigned long long* k_in, unsigned long long* k_out, uint32_t* v_in, uint32_t* v_out,
                        long long n, int end_bit) {
    size_t tmp = 0;
    PCR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, n, 0, end_bit, ctx->stream));
    PCR_CUDA(ctx->cub_tmp.ensure(tmp));
    PCR_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, k_in, k_out, v_in, v_out, n, 0, end_bit, ctx->stream));
    ctx->launches += 4;
    return PCR_OK;
}

static int exclusive_sum_u32(pcr_ctx* ctx, const uint32_t* in, uint32_t* out, long long n) {
    size_t tmp = 0;
    PCR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, ctx->stream));
    PCR_CUDA(ctx->cub_tmp.ensure(tmp));
    PCR_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, out, n, ctx->stream));
    ctx->launches += 2;
    return PCR_OK;
}

// segment count = ords[n-1] + flags[n-1]
static int read_segment_count(pcr_ctx* ctx, const uint32_t* flags, const uint32_t* ords, long long n, uint32_t* out) {
    uint32_t a = 0, b = 0;
    PCR_CUDA(cudaMemcpyAsync(&a, ords + (n - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaMemcpyAsync(&b, flags + (n - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = a + b;
    return PCR_OK;
}

constexpr unsigned long long kMaxBricks = 1ull << 27;   // 2 GiB of brick records

// Build a brick grid over n points (device float[3n]) with automatically chosen cell edge.
int build_point_grid(pcr_ctx* ctx, const float* d_xyz, long long n, Grid& g, DevBuf* sorted_payload_out) {
    if (n <= 0) return fail(ctx, PCR_ERR_ARG, "cannot index an empty point set");
    if (n >= (1ll << 31)) return fail(ctx, PCR_ERR_LIMIT, "point count exceeds 2^31-1");
    g.release();
    // ---- bounding box ----
    PCR_CUDA(ctx->tmp_e.ensure(64));
    int* d_mm = ctx->tmp_e.as<int>();
    init_minmax_kernel<<<1, 32, 0, ctx->stream>>>(d_mm);
    PCR_LAUNCH_CHECK();
    bbox_kernel<<<min(blocks_for(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(d_xyz, n, d_mm);
    PCR_LAUNCH_CHECK();
    int h_mm[6];
    PCR_CUDA(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_mm[0] == INT_MAX) return fail(ctx, PCR_ERR_ARG, "point set has no finite points");
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = ord2f_host(h_mm[a]); hi[a] = ord2f_host(h_mm[3 + a]); }
    double ext[3], maxext = 0, maxabs = 0;
    for (int a = 0; a < 3; ++a) {
        ext[a] = (double)hi[a] - (double)lo[a];
        maxext = std::max(maxext, ext[a]);
        maxabs = std::max(maxabs, std::max(fabs((double)lo[a]), fabs((double)hi[a])));
    }
    if (maxext <= 0) maxext = 1.0;
    double vol = 1.0;
    for (int a = 0; a < 3; ++a) vol *= std::max(ext[a], 1e-3 * maxext);
    const double target_ppc = ctx->target_ppc > 0.0 ? ctx->target_ppc : (n <= 4000000 ? 10.0 : 24.0);
    double h = cbrt(vol * target_ppc / (double)n);
    h = std::max(h, maxext * 1e-5);

    PCR_CUDA(ctx->tmp_a.ensure((size_t)n * 8));
    PCR_CUDA(ctx->tmp_b.ensure((size_t)n * 8));
    PCR_CUDA(ctx->tmp_c.ensure((size_t)n * 4 * 2));
    PCR_CUDA(ctx->tmp_d.ensure((size_t)n * 4 * 2));
    unsigned long long* k_in = ctx->tmp_a.as<unsigned long long>();
    unsigned long long* k_out = ctx->tmp_b.as<unsigned long long>();
    uint32_t* v_in = ctx->tmp_c.as<uint32_t>();
    uint32_t* v_out = v_in + n;
    uint32_t* flags = ctx->tmp_d.as<uint32_t>();
    uint32_t* ords = flags + n;

    const int max_attempts = 4;
    for (int attempt = 0; attempt < max_attempts; ++attempt) {
        GridView V{};
        unsigned long long nbricks;
        for (;;) {   // enlarge h until the dense brick table fits
            V.h = (float)h;
            V.inv_h = (float)(1.0 / h);
            // two empty cells of padding around the bounding box: scan points that leave the box by
            // up to two cells still fall into (band) cells of the grid and get their shell list
            V.ox = lo[0] - 2.25f * V.h; V.oy = lo[1] - 2.25f * V.h; V.oz = lo[2] - 2.25f * V.h;
            double c[3];
            for (int a = 0; a < 3; ++a) c[a] = floor(((double)hi[a] - (double)(lo[a] - 2.25f * V.h)) / h) + 4.0;
            V.bnx = (int)std::min(c[0] / 4.0 + 1.0, 2.0e9); V.bny = (int)std::min(c[1] / 4.0 + 1.0, 2.0e9); V.bnz = (int)std::min(c[2] / 4.0 + 1.0, 2.0e9);
            nbricks = (unsigned long long)V.bnx * V.bny * V.bnz;
            if ((double)V.bnx * V.bny * V.bnz <= (double)kMaxBricks && V.bnx < (1 << 20) && V.bny < (1 << 20) && V.bnz < (1 << 20)) break;
            h *= 1.2599210498948732;
        }
        V.cnx = V.bnx * 4; V.cny = V.bny * 4; V.cnz = V.bnz * 4;
        V.slack = 1e-3f + 1e-6f * (float)(2.0 * maxabs / h + (double)std::max(V.cnx, std::max(V.cny, V.cnz)));
        V.n_pts = (uint32_t)n;

        point_key_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, V, k_in, v_in);
        PCR_LAUNCH_CHECK();
        int rc = sort_pairs64(ctx, k_in, k_out, v_in, v_out, n, bits_for(nbricks * 64ull));
        if (rc) return rc;
        head_flag_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(k_out, n, flags);
        PCR_LAUNCH_CHECK();
        rc = exclusive_sum_u32(ctx, flags, ords, n);
        if (rc) return rc;
        uint32_t n_cells = 0;
        rc = read_segment_count(ctx, flags, ords, n, &n_cells);
        if (rc) return rc;
        const double ppc = (double)n / (double)n_cells;
        if (attempt + 1 < max_attempts && (ppc > 2.0 * target_ppc || (ppc < target_ppc / 2.4 && n_cells > 64))) {
            double ratio = sqrt(target_ppc / ppc);          // points lie on 2-D surfaces: ppc ~ h^2
            ratio = std::min(std::max(ratio, 0.2), 5.0);
            h *= ratio;
            continue;
        }
        // ---- accept this cell size: materialise the structure ----
        PCR_CUDA(g.bricks.ensure((size_t)nbricks * sizeof(BrickRec)));
        PCR_CUDA(cudaMemsetAsync(g.bricks.p, 0, (size_t)nbricks * sizeof(BrickRec), ctx->stream));
        PCR_CUDA(g.cell_start.ensure(((size_t)n_cells + 1) * 4));
        PCR_CUDA(g.pts.ensure(((size_t)n + 4) * sizeof(float4)));
        pad_tail_kernel<<<1, 32, 0, ctx->stream>>>(g.pts.as<float4>(), n, nullptr, 0);
        PCR_LAUNCH_CHECK();
        grid_fill_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(k_out, flags, ords, n, n_cells,
                                                                      g.bricks.as<BrickRec>(), g.cell_start.as<uint32_t>());
        PCR_LAUNCH_CHECK();
        gather_points_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_xyz, v_out, n, g.pts.as<float4>());
        PCR_LAUNCH_CHECK();
        if (sorted_payload_out) {
            PCR_CUDA(sorted_payload_out->ensure((size_t)n * 4));
            PCR_CUDA(cudaMemcpyAsync(sorted_payload_out->p, v_out, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        V.bricks = g.bricks.as<uint4>();
        V.cell_start = g.cell_start.as<uint32_t>();
        V.pts = g.pts.as<float4>();
        g.view = V;
        g.n_cells = n_cells;
        g.built = true;
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        return PCR_OK;
    }
    return fail(ctx, PCR_ERR_LIMIT, "grid build did not converge");
}

// ---------------------------------------------------------------------------------------
// k-NN kernels
// ---------------------------------------------------------------------------------------
template <int KCAP>
__global__ void __launch_bounds__(128) knn_query_kernel(GridView G, const float* __restrict__ q, long long m, int k,
                                                        float* __restrict__ dist, long long* __restrict__ idx) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    BestK<KCAP> best;
    best.init(k, 3.0e38f);
    grid_search(G, qx, qy, qz, best);
    for (int r = 0; r < k; ++r) {
        if (r < best.cnt) {
            dist[i * k + r] = sqrtf(best.d2s[r]);
            idx[i * k + r] = (long long)__float_as_uint(G.pts[best.poss[r]].w);
        } else {
            dist[i * k + r] = __int_as_float(0x7f800000);
            idx[i * k + r] = (long long)G.n_pts;
        }
    }
}

__global__ void __launch_bounds__(128) nn_query_kernel(GridView G, const float* __restrict__ q, long long m,
                                                       float* __restrict__ dist, long long* __restrict__ idx) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    float d2;
    const int pos = grid_nn(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], 3.0e38f, d2);
    if (pos >= 0) {
        dist[i] = sqrtf(d2);
        idx[i] = (long long)__float_as_uint(G.pts[pos].w);
    } else {
        dist[i] = __int_as_float(0x7f800000);
        idx[i] = (long long)G.n_pts;
    }
}

// Normal of every target point from its k nearest neighbours (k includes the point itself):
// float32 sums of p and p p^T in neighbour-rank order, cov = E[pp^T] - mu mu^T in float32
// (estimate_normals.py:41-72, quirk Q5), eigenvector of the smallest eigenvalue.
template <int KCAP>
__global__ void __launch_bounds__(128) normals_kernel(GridView G, int k, float4* __restrict__ nrm_sorted,
                                                      float* __restrict__ nrm_orig) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)G.n_pts) return;
    const float4 self = G.pts[i];
    BestK<KCAP> best;
    best.init(k, 3.0e38f);
    grid_search(G, self.x, self.y, self.z, best);
    float sx = 0.f, sy = 0.f, sz = 0.f, xx = 0.f, yy = 0.f, zz = 0.f, xy = 0.f, xz = 0.f, yz = 0.f;
    for (int r = 0; r < best.cnt; ++r) {
        const float4 t = G.pts[best.poss[r]];
        sx = fadd_rn(sx, t.x); sy = fadd_rn(sy, t.y); sz = fadd_rn(sz, t.z);
        xx = fadd_rn(xx, fmul_rn(t.x, t.x)); yy = fadd_rn(yy, fmul_rn(t.y, t.y)); zz = fadd_rn(zz, fmul_rn(t.z, t.z));
        xy = fadd_rn(xy, fmul_rn(t.x, t.y)); xz = fadd_rn(xz, fmul_rn(t.x, t.z)); yz = fadd_rn(yz, fmul_rn(t.y, t.z));
    }
    const float kf = (float)k;
    const float mx = fdiv_rn(sx, kf), my = fdiv_rn(sy, kf), mz = fdiv_rn(sz, kf);
    const float cxx = fsub_rn(fdiv_rn(xx, kf), fmul_rn(mx, mx));
    const float cyy = fsub_rn(fdiv_rn(yy, kf), fmul_rn(my, my));
    const float czz = fsub_rn(fdiv_rn(zz, kf), fmul_rn(mz, mz));
    const float cxy = fsub_rn(fdiv_rn(xy, kf), fmul_rn(mx, my));
    const float cxz = fsub_rn(fdiv_rn(xz, kf), fmul_rn(mx, mz));
    const float cyz = fsub_rn(fdiv_rn(yz, kf), fmul_rn(my, mz));
    double vx, vy, vz;
    smallest_eigvec_sym3(cxx, cxy, cxz, cyy, cyz, czz, vx, vy, vz, nullptr);
    nrm_sorted[i] = make_float4((float)vx, (float)vy, (float)vz, 0.f);
    const size_t j = __float_as_uint(self.w);
    nrm_orig[3 * j] = (float)vx; nrm_orig[3 * j + 1] = (float)vy; nrm_orig[3 * j + 2] = (float)vz;
}

// caller order (n,3) -> sorted float4 and back
__global__ void scatter_normals_kernel(const float* __restrict__ nrm_orig, const float4* __restrict__ pts, long long n,
                                       float4* __restrict__ nrm_sorted) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t j = __float_as_uint(pts[i].w);
    nrm_sorted[i] = make_float4(nrm_orig[3 * j], nrm_orig[3 * j + 1], nrm_orig[3 * j + 2], 0.f);
}

// PlaneICP record: (point, normal) of one target point side by side -- ONE 32-byte sector per
// gathered correspondence in the accumulate pass instead of two
__global__ void interleave_pn_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, long long n, float4* __restrict__ pn) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    pn[2 * i] = pts[i];
    pn[2 * i + 1] = nrm[i];
}

// ---------------------------------------------------------------------------------------
// voxel statistics
// ---------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }

// voxel coordinate = floor(p / voxel_size) in the INPUT precision (voxel.py:16)
template <typename T>
__global__ void voxel_coord_kernel(const T* __restrict__ xyz, long long n, T vs, int* __restrict__ coords, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            T v = floor(div_rn<T>(xyz[3 * i + a], vs));
            if (!(v == v)) v = 0;
            v = v < (T)-1.0e9 ? (T)-1.0e9 : (v > (T)1.0e9 ? (T)1.0e9 : v);
            const int c = (int)v;
            coords[3 * i + a] = c;
            lo[a] = min(lo[a], c); hi[a] = max(hi[a], c);
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
        for (int a = 0; a < 3; ++a) {
            atomicMin(&mm[a], lo[a]);
            atomicMax(&mm[3 + a], hi[a]);
        }
    }
}

__global__ void voxel_key_kernel(const int* __restrict__ coords, long long n, int mnx, int mny, int mnz, int bnx, int bny,
                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = coords[3 * i] - mnx, cy = coords[3 * i + 1] - mny, cz = coords[3 * i + 2] - mnz;
    unsigned long long brick = ((unsigned long long)(cz >> 2) * bny + (cy >> 2)) * bnx + (cx >> 2);
    keys[i] = brick * 64ull + (unsigned long long)brick_bit(cx, cy, cz);
    vals[i] = (uint32_t)i;
}

__global__ void segment_start_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ ords, long long n,
                                     uint32_t n_seg, uint32_t* __restrict__ seg) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i == 0) seg[n_seg] = (uint32_t)n;
    if (i >= n || !flags[i]) return;
    seg[ords[i]] = (uint32_t)i;
}

// One thread per occupied voxel: sequential float64 sums in original point order (the radix
// sort is stable), i.e. the order np.bincount uses (voxel.py:113-143).  Two passes about the
// mean; sample covariance / max(n-1, 1).
template <typename T>
__global__ void voxel_stats_kernel(const T* __restrict__ xyz, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ seg,
                                   uint32_t n_seg, int min_points, double* __restrict__ mean, double* __restrict__ cov6,
                                   uint32_t* __restrict__ keep) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_seg) return;
    const uint32_t s = seg[v], e = seg[v + 1];
    const double cnt = (double)(e - s);
    double sx = 0, sy = 0, sz = 0;
    for (uint32_t p = s; p < e; ++p) {
        const size_t j = vals[p];
        sx = dadd_rn(sx, (double)xyz[3 * j]); sy = dadd_rn(sy, (double)xyz[3 * j + 1]); sz = dadd_rn(sz, (double)xyz[3 * j + 2]);
    }
    const double mx = ddiv_rn(sx, cnt), my = ddiv_rn(sy, cnt), mz = ddiv_rn(sz, cnt);
    double cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0;
    for (uint32_t p = s; p < e; ++p) {
        const size_t j = vals[p];
        const double dx = dsub_rn((double)xyz[3 * j], mx), dy = dsub_rn((double)xyz[3 * j + 1], my), dz = dsub_rn((double)xyz[3 * j + 2], mz);
        cxx = dadd_rn(cxx, dmul_rn(dx, dx)); cxy = dadd_rn(cxy, dmul_rn(dx, dy)); cxz = dadd_rn(cxz, dmul_rn(dx, dz));
        cyy = dadd_rn(cyy, dmul_rn(dy, dy)); cyz = dadd_rn(cyz, dmul_rn(dy, dz)); czz = dadd_rn(czz, dmul_rn(dz, dz));
    }
    const double den = (e - s) > 1 ? (double)(e - s - 1) : 1.0;
    mean[3 * (size_t)v] = mx; mean[3 * (size_t)v + 1] = my; mean[3 * (size_t)v + 2] = mz;
    double* c = cov6 + 6 * (size_t)v;
    c[0] = ddiv_rn(cxx, den); c[1] = ddiv_rn(cxy, den); c[2] = ddiv_rn(cxz, den);
    c[3] = ddiv_rn(cyy, den); c[4] = ddiv_rn(cyz, den); c[5] = ddiv_rn(czz, den);
    keep[v] = ((int)(e - s) >= min_points) ? 1u : 0u;
}

// Compact kept voxels and derive everything the hot path needs from them.
__global__ void voxel_finalize_kernel(const unsigned long long* __restrict__ keys_sorted, const uint32_t* __restrict__ seg,
                                      const uint32_t* __restrict__ keep, const uint32_t* __restrict__ kord, uint32_t n_seg,
                                      uint32_t n_keep, const double* __restrict__ mean_all, const double* __restrict__ cov6_all,
                                      int with_icov, double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ norm,
                                      double* __restrict__ icov, long long* __restrict__ count, BrickRec* __restrict__ bricks,
                                      uint32_t* __restrict__ cell_start, float4* __restrict__ pts, float4* __restrict__ rec_plane,
                                      float4* __restrict__ rec_ndt) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v == 0) cell_start[n_keep] = n_keep;
    if (v >= n_seg || !keep[v]) return;
    const uint32_t o = kord[v];
    const double mx = mean_all[3 * (size_t)v], my = mean_all[3 * (size_t)v + 1], mz = mean_all[3 * (size_t)v + 2];
    const double* c6 = cov6_all + 6 * (size_t)v;
    mean[3 * (size_t)o] = mx; mean[3 * (size_t)o + 1] = my; mean[3 * (size_t)o + 2] = mz;
    double C[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    for (int t = 0; t < 9; ++t) cov[9 * (size_t)o + t] = C[t];
    count[o] = (long long)(seg[v + 1] - seg[v]);
    double nx, ny, nz;
    smallest_eigvec_sym3(c6[0], c6[1], c6[2], c6[3], c6[4], c6[5], nx, ny, nz, nullptr);
    norm[3 * (size_t)o] = nx; norm[3 * (size_t)o + 1] = ny; norm[3 * (size_t)o + 2] = nz;
    rec_plane[2 * (size_t)o] = make_float4((float)mx, (float)my, (float)mz, 0.f);
    rec_plane[2 * (size_t)o + 1] = make_float4((float)nx, (float)ny, (float)nz, 0.f);
    if (with_icov) {
        double W[9];
        icov_closed_form(C, W);
        for (int t = 0; t < 9; ++t) icov[9 * (size_t)o + t] = W[t];
        rec_ndt[3 * (size_t)o] = make_float4((float)mx, (float)my, (float)mz, (float)W[0]);
        rec_ndt[3 * (size_t)o + 1] = make_float4((float)W[1], (float)W[2], (float)W[4], (float)W[5]);
        rec_ndt[3 * (size_t)o + 2] = make_float4((float)W[8], 0.f, 0.f, 0.f);
    }
    // NN index over the kept means: one point per cell, cell ordinal == kept ordinal
    const unsigned long long key = keys_sorted[seg[v]];
    const unsigned long long brick = key >> 6;
    atomicOr(&bricks[brick].mask, 1ull << (key & 63ull));
    atomicMin(&bricks[brick].base, o);
    cell_start[o] = o;
    pts[o] = make_float4((float)mx, (float)my, (float)mz, __uint_as_float(o));
}

__global__ void brick_base_init_kernel(BrickRec* bricks, unsigned long long n) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) { bricks[i].mask = 0ull; bricks[i].base = 0xffffffffu; bricks[i].pad = 0u; }
}

__global__ void voxel_centroid_kernel(const double* __restrict__ mean_all, uint32_t n_seg, float* __restrict__ out) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_seg) return;
    out[3 * (size_t)v] = (float)mean_all[3 * (size_t)v];
    out[3 * (size_t)v + 1] = (float)mean_all[3 * (size_t)v + 1];
    out[3 * (size_t)v + 2] = (float)mean_all[3 * (size_t)v + 2];
}

__global__ void voxel_query_kernel(GridView G, CandLists L, int use_lists, const float* __restrict__ q, long long m,
                                   const double* __restrict__ mean, long long* __restrict__ vidx, double* __restrict__ dist) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    float d2;
    const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    int pos;
    if (!(use_lists && list_nn(G, L, qx, qy, qz, 3.0e38f, d2, pos))) pos = grid_nn(G, qx, qy, qz, 3.0e38f, d2);
    if (pos >= 0) {
        const size_t o = __float_as_uint(G.pts[pos].w);
        const double dx = (double)qx - mean[3 * o], dy = (double)qy - mean[3 * o + 1], dz = (double)qz - mean[3 * o + 2];
        vidx[i] = (long long)o;
        dist[i] = sqrt(dx * dx + dy * dy + dz * dz);
    } else {
        vidx[i] = (long long)G.n_pts;
        dist[i] = __longlong_as_double(0x7ff0000000000000ll);
    }
}

// Shared front end of pcr_build_voxels / pcr_voxel_filter: upload, voxel coordinates, sort,
// segments, per-voxel mean / covariance.  Leaves on the device:
//   tmp_b: sorted keys (u64[n]); tmp_c[n..2n): sorted point indices; seg_buf: u32[n_seg+1];
//   mean_all / cov6_all / keep (per occupied voxel).
struct VoxelFront {
    uint32_t n_seg = 0;
    int mn[3] = {0, 0, 0};
    int bn[3] = {0, 0, 0};
    DevBuf xyz, seg, mean_all, cov6_all, keep;
    void release() { xyz.release(); seg.release(); mean_all.release(); cov6_all.release(); keep.release(); }
};

template <typename T>
static int voxel_front(pcr_ctx* ctx, const void* xyz, long long n, double voxel_size, int min_points, VoxelFront& F) {
    if (n <= 0) return fail(ctx, PCR_ERR_ARG, "cannot voxelise an empty point set");
    if (n >= (1ll << 31)) return fail(ctx, PCR_ERR_LIMIT, "point count exceeds 2^31-1");
    if (!(voxel_size > 0)) return fail(ctx, PCR_ERR_ARG, "voxel_size must be positive");
    const T* d_xyz;
    if (is_device_pointer(xyz)) {
        d_xyz = (const T*)xyz;
    } else {
        PCR_CUDA(F.xyz.ensure((size_t)n * 3 * sizeof(T)));
        PCR_CUDA(cudaMemcpyAsync(F.xyz.p, xyz, (size_t)n * 3 * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        d_xyz = F.xyz.as<T>();
    }
    PCR_CUDA(ctx->tmp_e.ensure(64 + (size_t)n * 12));
    int* d_mm = ctx->tmp_e.as<int>();
    int* coords = d_mm + 16;
    init_minmax_kernel<<<1, 32, 0, ctx->stream>>>(d_mm);
    PCR_LAUNCH_CHECK();
    voxel_coord_kernel<T><<<min(blocks_for(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(d_xyz, n, (T)voxel_size, coords, d_mm);
    PCR_LAUNCH_CHECK();
    int h_mm[6];
    PCR_CUDA(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    double nb[3];
    for (int a = 0; a < 3; ++a) {
        F.mn[a] = h_mm[a];
        nb[a] = floor(((double)h_mm[3 + a] - (double)h_mm[a]) / 4.0) + 1.0;
    }
    if (nb[0] * nb[1] * nb[2] > (double)kMaxBricks || nb[0] > 1e6 || nb[1] > 1e6 || nb[2] > 1e6)
        return fail(ctx, PCR_ERR_LIMIT, "voxel grid extent too large for the dense brick table; use a larger voxel_size");
    for (int a = 0; a < 3; ++a) F.bn[a] = (int)nb[a];
    const unsigned long long nbricks = (unsigned long long)F.bn[0] * F.bn[1] * F.bn[2];

    PCR_CUDA(ctx->tmp_a.ensure((size_t)n * 8));
    PCR_CUDA(ctx->tmp_b.ensure((size_t)n * 8));
    PCR_CUDA(ctx->tmp_c.ensure((size_t)n * 4 * 2));
    PCR_CUDA(ctx->tmp_d.ensure((size_t)n * 4 * 2));
    unsigned long long* k_in = ctx->tmp_a.as<unsigned long long>();
    unsigned long long* k_out = ctx->tmp_b.as<unsigned long long>();
    uint32_t* v_in = ctx->tmp_c.as<uint32_t>();
    uint32_t* v_out = v_in + n;
    uint32_t* flags = ctx->tmp_d.as<uint32_t>();
    uint32_t* ords = flags + n;
    voxel_key_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(coords, n, F.mn[0], F.mn[1], F.mn[2], F.bn[0], F.bn[1], k_in, v_in);
    PCR_LAUNCH_CHECK();
    int rc = sort_pairs64(ctx, k_in, k_out, v_in, v_out, n, bits_for(nbricks * 64ull));
    if (rc) return rc;
    head_flag_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(k_out, n, flags);
    PCR_LAUNCH_CHECK();
    rc = exclusive_sum_u32(ctx, flags, ords, n);
    if (rc) return rc;
    rc = read_segment_count(ctx, flags, ords, n, &F.n_seg);
    if (rc) return rc;
    PCR_CUDA(F.seg.ensure(((size_t)F.n_seg + 1) * 4));
    PCR_CUDA(F.mean_all.ensure((size_t)F.n_seg * 3 * 8));
    PCR_CUDA(F.cov6_all.ensure((size_t)F.n_seg * 6 * 8));
    PCR_CUDA(F.keep.ensure((size_t)F.n_seg * 4 * 2));
    segment_start_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(flags, ords, n, F.n_seg, F.seg.as<uint32_t>());
    PCR_LAUNCH_CHECK();
    voxel_stats_kernel<T><<<blocks_for(F.n_seg, 128), 128, 0, ctx->stream>>>(d_xyz, v_out, F.seg.as<uint32_t>(), F.n_seg, min_points,
                                                                            F.mean_all.as<double>(), F.cov6_all.as<double>(),
                                                                            F.keep.as<uint32_t>());
    PCR_LAUNCH_CHECK();
    return PCR_OK;
}


// ---------------------------------------------------------------------------------------
// per-cell candidate lists over the kept voxel means (see CandLists in pcr_common.cuh)
// ---------------------------------------------------------------------------------------
// ctx->list_dilate (default 3): cells within this Chebyshev distance of a kept voxel get a list
// ctx->list_radius (default 5): the build looks at the (2R+1)^3 neighbourhood; exact while D_C < R

// mark the neighbourhood of every kept voxel as "band" (cells that get a list)
__global__ void band_mark_kernel(GridView G, BrickRec* __restrict__ lbricks, int dilate) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= G.n_pts) return;
    const float4 m = G.pts[v];
    const int cx = cell_of((m.x - G.ox) * G.inv_h, G.cnx), cy = cell_of((m.y - G.oy) * G.inv_h, G.cny), cz = cell_of((m.z - G.oz) * G.inv_h, G.cnz);
    const int x0 = max(cx - dilate, 0), x1 = min(cx + dilate, G.cnx - 1);
    const int y0 = max(cy - dilate, 0), y1 = min(cy + dilate, G.cny - 1);
    const int z0 = max(cz - dilate, 0), z1 = min(cz + dilate, G.cnz - 1);
    for (int bz = z0 >> 2; bz <= z1 >> 2; ++bz)
        for (int by = y0 >> 2; by <= y1 >> 2; ++by)
            for (int bx = x0 >> 2; bx <= x1 >> 2; ++bx) {
                const unsigned long long mk = brick_box_mask(max(x0 - bx * 4, 0), min(x1 - bx * 4, 3), max(y0 - by * 4, 0), min(y1 - by * 4, 3),
                                                             max(z0 - bz * 4, 0), min(z1 - bz * 4, 3));
                BrickRec* r = &lbricks[((size_t)bz * G.bny + by) * G.bnx + bx];
                if ((r->mask & mk) != mk) atomicOr(&r->mask, mk);
            }
}

__global__ void brick_popc_kernel(const BrickRec* __restrict__ bricks, unsigned long long n, uint32_t* __restrict__ cnt) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = (uint32_t)__popcll(bricks[i].mask);
}

__global__ void brick_base_set_kernel(BrickRec* __restrict__ bricks, unsigned long long n, const uint32_t* __restrict__ base) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) bricks[i].base = base[i];
}

// One thread per (brick, bit) of the band.  FILL = false: D2[ordinal] = D_C^2 (or -1: no list) and
// counts[ordinal] = list length; FILL = true: write the list.  The neighbourhood is walked brick
// by brick through the occupancy masks, so empty cells cost nothing.
template <bool FILL>
__global__ void list_build_kernel(GridView G, const BrickRec* __restrict__ lbricks, unsigned long long nbricks, float* __restrict__ D2s,
                                  uint32_t* __restrict__ counts, const uint32_t* __restrict__ list_start, uint32_t* __restrict__ list_idx,
                                  float4* __restrict__ list_pts, int kListRadius) {
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long b = tid >> 6;
    const int bit = (int)(tid & 63ull);
    if (b >= nbricks) return;
    const BrickRec lr = lbricks[b];
    if (!((lr.mask >> bit) & 1ull)) return;
    const uint32_t ord = lr.base + (uint32_t)__popcll(lr.mask & ((1ull << bit) - 1ull));
    const int bx = (int)(b % (unsigned long long)G.bnx), by = (int)((b / (unsigned long long)G.bnx) % (unsigned long long)G.bny),
              bz = (int)(b / ((unsigned long long)G.bnx * G.bny));
    const int cx = bx * 4 + (bit & 3), cy = by * 4 + ((bit >> 2) & 3), cz = bz * 4 + (bit >> 4);
    const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
    const int x0 = max(cx - kListRadius, 0), x1 = min(cx + kListRadius, G.cnx - 1);
    const int y0 = max(cy - kListRadius, 0), y1 = min(cy + kListRadius, G.cny - 1);
    const int z0 = max(cz - kListRadius, 0), z1 = min(cz + kListRadius, G.cnz - 1);
    float D2 = FILL ? D2s[ord] : 3.0e38f;
    if (FILL && D2 < 0.f) return;
    for (int pass = FILL ? 1 : 0; pass < 2; ++pass) {
        float lim2 = 0.f;
        uint32_t n_out = 0, w = 0;
        if (pass == 1) {
            const float D = sqrtf(D2);
            if (!FILL && !(D < (float)kListRadius - 2.0f * G.slack)) {   // neighbourhood too small to be exact here: no list
                D2s[ord] = -1.0f;
                counts[ord] = 0u;
                return;
            }
            lim2 = (D + G.slack) * (D + G.slack);
            if (FILL) w = list_start[ord];
        }
        for (int nz = z0 >> 2; nz <= z1 >> 2; ++nz)
            for (int ny = y0 >> 2; ny <= y1 >> 2; ++ny)
                for (int nx = x0 >> 2; nx <= x1 >> 2; ++nx) {
                    const uint4 rec = G.bricks[((size_t)nz * G.bny + ny) * G.bnx + nx];
                    const unsigned long long occ = ((unsigned long long)rec.y << 32) | rec.x;
                    if (occ == 0ull) continue;
                    unsigned long long m = occ & brick_box_mask(max(x0 - nx * 4, 0), min(x1 - nx * 4, 3), max(y0 - ny * 4, 0), min(y1 - ny * 4, 3),
                                                                max(z0 - nz * 4, 0), min(z1 - nz * 4, 3));
                    while (m) {
                        const int nb = __ffsll((long long)m) - 1;
                        m &= m - 1ull;
                        const uint32_t o2 = rec.z + (uint32_t)__popcll(occ & ((1ull << nb) - 1ull));
                        const uint32_t s = G.cell_start[o2], e = G.cell_start[o2 + 1];
                        for (uint32_t p = s; p < e; ++p) {
                            const float4 mp = G.pts[p];
                            const float gx = (mp.x - G.ox) * G.inv_h, gy = (mp.y - G.oy) * G.inv_h, gz = (mp.z - G.oz) * G.inv_h;
                            if (pass == 0) {
                                const float ax = fmaxf(fabsf(gx - fx), fabsf(gx - fx - 1.0f));
                                const float ay = fmaxf(fabsf(gy - fy), fabsf(gy - fy - 1.0f));
                                const float az = fmaxf(fabsf(gz - fz), fabsf(gz - fz - 1.0f));
                                D2 = fminf(D2, ax * ax + ay * ay + az * az);
                            } else {
                                const float ax = fmaxf(fmaxf(fx - gx, gx - fx - 1.0f), 0.0f);
                                const float ay = fmaxf(fmaxf(fy - gy, gy - fy - 1.0f), 0.0f);
                                const float az = fmaxf(fmaxf(fz - gz, gz - fz - 1.0f), 0.0f);
                                if (ax * ax + ay * ay + az * az <= lim2) {
                                    if (FILL) {
                                        list_idx[w] = p;
                                        if (list_pts) list_pts[w] = make_float4(mp.x, mp.y, mp.z, __uint_as_float(p));
                                        ++w;
                                    }
                                    ++n_out;
                                }
                            }
                        }
                    }
                }
        if (pass == 1 && !FILL) { counts[ord] = n_out; D2s[ord] = D2; }
    }
}

static int build_voxel_lists(pcr_ctx* ctx) {
    Grid& g = ctx->vox_grid;
    ctx->vox_lists = CandLists{};
    ctx->n_band_cells = ctx->n_list_entries = 0;
    if (!g.built || g.view.n_pts == 0) return PCR_OK;
    if (const char* e = getenv("PCR_VOXEL_LISTS")) if (atoi(e) == 0) return PCR_OK;
    const GridView& G = g.view;
    const unsigned long long nbricks = (unsigned long long)G.bnx * G.bny * G.bnz;
    PCR_CUDA(ctx->vox_lbricks.ensure((size_t)nbricks * sizeof(BrickRec)));
    PCR_CUDA(cudaMemsetAsync(ctx->vox_lbricks.p, 0, (size_t)nbricks * sizeof(BrickRec), ctx->stream));
    BrickRec* lb = ctx->vox_lbricks.as<BrickRec>();
    band_mark_kernel<<<blocks_for(G.n_pts, 128), 128, 0, ctx->stream>>>(G, lb, ctx->list_dilate);
    PCR_LAUNCH_CHECK();
    PCR_CUDA(ctx->tmp_a.ensure((size_t)(nbricks + 1) * 4 * 2));
    uint32_t* bcnt = ctx->tmp_a.as<uint32_t>();
    uint32_t* bbase = bcnt + (nbricks + 1);
    PCR_CUDA(cudaMemsetAsync(bcnt, 0, (size_t)(nbricks + 1) * 4, ctx->stream));
    brick_popc_kernel<<<blocks_for((long long)nbricks, 256), 256, 0, ctx->stream>>>(lb, nbricks, bcnt);
    PCR_LAUNCH_CHECK();
    int rc = exclusive_sum_u32(ctx, bcnt, bbase, (long long)nbricks + 1);
    if (rc) return rc;
    uint32_t n_band = 0;
    PCR_CUDA(cudaMemcpyAsync(&n_band, bbase + nbricks, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    brick_base_set_kernel<<<blocks_for((long long)nbricks, 256), 256, 0, ctx->stream>>>(lb, nbricks, bbase);
    PCR_LAUNCH_CHECK();
    PCR_CUDA(ctx->tmp_b.ensure(((size_t)n_band + 1) * 4));
    PCR_CUDA(ctx->tmp_c.ensure(((size_t)n_band + 1) * 4));
    PCR_CUDA(ctx->vox_list_start.ensure(((size_t)n_band + 1) * 4));
    uint32_t* counts = ctx->tmp_b.as<uint32_t>();
    float* D2s = ctx->tmp_c.as<float>();
    PCR_CUDA(cudaMemsetAsync(counts, 0, ((size_t)n_band + 1) * 4, ctx->stream));
    const long long nthreads = (long long)nbricks * 64;
    list_build_kernel<false><<<blocks_for(nthreads, 128), 128, 0, ctx->stream>>>(G, lb, nbricks, D2s, counts, nullptr, nullptr, nullptr, ctx->list_radius);
    PCR_LAUNCH_CHECK();
    rc = exclusive_sum_u32(ctx, counts, ctx->vox_list_start.as<uint32_t>(), (long long)n_band + 1);
    if (rc) return rc;
    uint32_t n_entries = 0;
    PCR_CUDA(cudaMemcpyAsync(&n_entries, ctx->vox_list_start.as<uint32_t>() + n_band, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    PCR_CUDA(ctx->vox_list_idx.ensure(((size_t)n_entries + 4) * 4));
    pad_tail_kernel<<<1, 32, 0, ctx->stream>>>(nullptr, (long long)G.n_pts, ctx->vox_list_idx.as<uint32_t>(), (long long)n_entries);
    PCR_LAUNCH_CHECK();
    // the candidate means inline next to their indices (one load per candidate instead of two dependent ones),
    // unless that copy would be large (PCR_LIST_INLINE=0 disables it)
    float4* inline_pts = nullptr;
    {
        const char* e = getenv("PCR_LIST_INLINE");
        if ((!e || atoi(e) != 0) && (double)n_entries * 16.0 <= 4.0 * 1024.0 * 1024.0 * 1024.0) {
            PCR_CUDA(ctx->vox_list_pts.ensure(((size_t)n_entries + 4) * sizeof(float4)));
            inline_pts = ctx->vox_list_pts.as<float4>();
        }
    }
    list_build_kernel<true><<<blocks_for(nthreads, 128), 128, 0, ctx->stream>>>(G, lb, nbricks, D2s, nullptr, ctx->vox_list_start.as<uint32_t>(),
                                                                               ctx->vox_list_idx.as<uint32_t>(), inline_pts, ctx->list_radius);
    PCR_LAUNCH_CHECK();
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->vox_lists.bricks = ctx->vox_lbricks.as<uint4>();
    ctx->vox_lists.list_start = ctx->vox_list_start.as<uint32_t>();
    ctx->vox_lists.list_idx = ctx->vox_list_idx.as<uint32_t>();
    ctx->vox_lists.list_pts = inline_pts;
    ctx->n_band_cells = n_band;
    ctx->n_list_entries = n_entries;
    return PCR_OK;
}


// ---------------------------------------------------------------------------------------
// per-cell shell lists over the target-point grid (see ShellLists in pcr_common.cuh)
// ---------------------------------------------------------------------------------------
__constant__ float kShellFrac[PCR_SHELL_LEVELS] = {PCR_SHELL_FRACS};   // level margins in cell edges, see pcr_common.cuh

// One thread per (brick, bit) of the band.  FILL = false: counts[ordinal] = list length in GROUPS of
// four entries; FILL = true: write the entries level by level, the sentinels and the
// per-group margin bounds.  dmax <= R cell edges, so the (2R+1)^3 block holds every point within dmax.
template <bool FILL>
__global__ void shell_build_kernel(GridView G, const BrickRec* __restrict__ band, unsigned long long nbricks, float dmax, int R,
                                   uint32_t* __restrict__ counts, const uint32_t* __restrict__ start, float4* __restrict__ out,
                                   float* __restrict__ margin2) {
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long b = tid >> 6;
    const int bit = (int)(tid & 63ull);
    if (b >= nbricks) return;
    const BrickRec lr = band[b];
    if (!((lr.mask >> bit) & 1ull)) return;
    const uint32_t ord = lr.base + (uint32_t)__popcll(lr.mask & ((1ull << bit) - 1ull));
    const int bx = (int)(b % (unsigned long long)G.bnx), by = (int)((b / (unsigned long long)G.bnx) % (unsigned long long)G.bny),
              bz = (int)(b / ((unsigned long long)G.bnx * G.bny));
    const int cx = bx * 4 + (bit & 3), cy = by * 4 + ((bit >> 2) & 3), cz = bz * 4 + (bit >> 4);
    const float lox = G.ox + (float)cx * G.h, loy = G.oy + (float)cy * G.h, loz = G.oz + (float)cz * G.h;
    const float hix = lox + G.h, hiy = loy + G.h, hiz = loz + G.h;
    float lim2[PCR_SHELL_LEVELS];
#pragma unroll
    for (int l = 0; l < PCR_SHELL_LEVELS; ++l) lim2[l] = (kShellFrac[l] * G.h) * (kShellFrac[l] * G.h);
    const float dmax2 = dmax * dmax;
    uint32_t cnt[PCR_SHELL_LEVELS];
#pragma unroll
    for (int l = 0; l < PCR_SHELL_LEVELS; ++l) cnt[l] = 0u;
    const size_t base = FILL ? 4 * (size_t)start[ord] : 0;      // list offsets count groups of four entries
    for (int pass = 0; pass < (FILL ? 2 : 1); ++pass) {
        long long cached = -1;
        uint4 rec = make_uint4(0, 0, 0, 0);
        const int side = 2 * R + 1, ncell = side * side * side, centre = (ncell - 1) / 2;
        for (int oo = 0; oo < ncell; ++oo) {
            // own cell first, then the rest of the block in z, y, x order
            const int o = oo == 0 ? centre : (oo <= centre ? oo - 1 : oo);
            const int nx = cx + o % side - R, ny = cy + (o / side) % side - R, nz = cz + o / (side * side) - R;
            if (nx < 0 || ny < 0 || nz < 0 || nx >= G.cnx || ny >= G.cny || nz >= G.cnz) continue;
            const long long nbk = ((long long)(nz >> 2) * G.bny + (ny >> 2)) * G.bnx + (nx >> 2);
            if (nbk != cached) { rec = G.bricks[nbk]; cached = nbk; }
            const unsigned long long occ = ((unsigned long long)rec.y << 32) | rec.x;
            const int nbit = brick_bit(nx, ny, nz);
            if (!((occ >> nbit) & 1ull)) continue;
            const uint32_t o2 = rec.z + (uint32_t)__popcll(occ & ((1ull << nbit) - 1ull));
            const uint32_t s = G.cell_start[o2], e = G.cell_start[o2 + 1];
            for (uint32_t p = s; p < e; ++p) {
                const float4 t = G.pts[p];
                int lvl = 0;
                if (oo != 0) {                                // the cell's own points are level 0 by definition
                    const float mx = fmaxf(fmaxf(lox - t.x, t.x - hix), 0.0f), my = fmaxf(fmaxf(loy - t.y, t.y - hiy), 0.0f),
                                mz = fmaxf(fmaxf(loz - t.z, t.z - hiz), 0.0f);
                    const float m2 = mx * mx + my * my + mz * mz;
                    if (m2 > dmax2) continue;                 // farther than dmax from the cell
                    lvl = PCR_SHELL_LEVELS - 1;
#pragma unroll
                    for (int l = PCR_SHELL_LEVELS - 2; l >= 1; --l) if (m2 <= lim2[l]) lvl = l;
                }
                if (pass == 0) {
#pragma unroll
                    for (int l = 0; l < PCR_SHELL_LEVELS; ++l) if (l == lvl) cnt[l]++;
                } else {
                    uint32_t w = 0;
#pragma unroll
                    for (int l = 0; l < PCR_SHELL_LEVELS; ++l) if (l == lvl) w = cnt[l]++;
                    float* grp = reinterpret_cast<float*>(out + base + (w & ~3u));   // group of four, structure of arrays
                    const uint32_t j = w & 3u;
                    grp[j] = t.x; grp[4 + j] = t.y; grp[8 + j] = t.z; grp[12 + j] = __uint_as_float(p);
                }
            }
        }
        if (pass == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int l = 0; l < PCR_SHELL_LEVELS; ++l) total += cnt[l];
            const uint32_t padded = (total + 3u) & ~3u;
            if (!FILL) { counts[ord] = padded >> 2; return; }
            // per-group lower bound of the squared margin (slack covers the query's binning error),
            // then turn the level counts into write cursors for the second pass
            const float slack_w = G.slack * G.h;
            uint32_t lvl_end = 0;
            int l_of = 0;
            uint32_t ends[PCR_SHELL_LEVELS];
#pragma unroll
            for (int l = 0; l < PCR_SHELL_LEVELS; ++l) { lvl_end += cnt[l]; ends[l] = lvl_end; }
            for (uint32_t g4 = 0; g4 < padded; g4 += 4) {
                while (l_of < PCR_SHELL_LEVELS - 1 && g4 >= ends[l_of]) ++l_of;     // level of the group's first entry
                float lb = 0.0f;
                if (g4 >= total) lb = 3.0e38f;                                        // (cannot happen: groups start below total)
                else if (l_of >= 1) lb = fmaxf(kShellFrac[l_of - 1] * G.h - slack_w, 0.0f);
                margin2[(base + g4) >> 2] = lb * lb;
            }
            for (uint32_t k = total; k < padded; ++k) {       // sentinels complete the last group
                float* grp = reinterpret_cast<float*>(out + base + (k & ~3u));
                const uint32_t j = k & 3u;
                grp[j] = 3.0e38f; grp[4 + j] = 3.0e38f; grp[8 + j] = 3.0e38f; grp[12 + j] = __uint_as_float(0xffffffffu);
            }
            uint32_t run = 0;
#pragma unroll
            for (int l = 0; l < PCR_SHELL_LEVELS; ++l) { const uint32_t c = cnt[l]; cnt[l] = run; run += c; }
        }
    }
}

struct U32ToU64 {
    __host__ __device__ unsigned long long operator()(uint32_t v) const { return (unsigned long long)v; }
};

// `which` = 0: lists over the target points (ICP / PlaneICP), 1: over the kept voxel means (VPlaneICP / NDT:
// one mean per cell at most, so a list is the margin-ordered set of means around the cell and a query
// usually stops after its first group -- the exact candidate lists of round 1 made every query evaluate
// its cell's whole candidate set, ~20 means, through an index indirection).
static int build_shell_lists(pcr_ctx* ctx, int which = 0) {
    Grid& g = which == 0 ? ctx->tgt_grid : ctx->vox_grid;
    ShellLists& view = which == 0 ? ctx->tgt_shell : ctx->vox_shell;
    DevBuf& b_bricks = which == 0 ? ctx->shell_bricks : ctx->vshell_bricks;
    DevBuf& b_start = which == 0 ? ctx->shell_start : ctx->vshell_start;
    DevBuf& b_pts = which == 0 ? ctx->shell_pts : ctx->vshell_pts;
    DevBuf& b_margin2 = which == 0 ? ctx->shell_margin2 : ctx->vshell_margin2;
    long long& n_band_out = which == 0 ? ctx->n_shell_band : ctx->n_vshell_band;
    long long& n_entries_out = which == 0 ? ctx->n_shell_entries : ctx->n_vshell_entries;
    double& dmax_used = which == 0 ? ctx->shell_dmax_used : ctx->vshell_dmax_used;
    view = ShellLists{};
    n_band_out = n_entries_out = 0;
    dmax_used = 0.0;
    if (!g.built || g.view.n_pts == 0) return PCR_OK;
    if (which == 0) {
        if (const char* e = getenv("PCR_SHELL_LISTS")) if (atoi(e) == 0) return PCR_OK;
    } else {
        // measured (profiles/r2_notes.md): same late-iteration time as the exact candidate lists, +4 % on NDT 10M,
        // but the first two iterations of VPlaneICP 10M take 1.7x longer (queries displaced beyond the margin fall
        // back to the grid walk; the candidate lists cover them) -- opt-in
        const char* e = getenv("PCR_VOXEL_SHELL");
        if (!e || atoi(e) == 0) return PCR_OK;
    }
    const GridView& G = g.view;
    const unsigned long long nbricks = (unsigned long long)G.bnx * G.bny * G.bnz;
    PCR_CUDA(b_bricks.ensure((size_t)nbricks * sizeof(BrickRec)));
    BrickRec* lb = b_bricks.as<BrickRec>();
    const long long nthreads = (long long)nbricks * 64;
    // the requested margin first, then smaller ones until the lists fit the memory cap
    const double tries[5] = {ctx->shell_dmax_frac, 2.0, 1.5, 1.0, 0.5};
    for (int t = 0; t < 5; ++t) {
        const double frac = tries[t];
        if (t > 0 && frac >= tries[0]) continue;
        const int R = frac <= 1.0 ? 1 : (frac <= 2.0 ? 2 : 3);   // build neighbourhood and band dilation (cells)
        const float dmax = (float)(frac * (double)G.h);
        PCR_CUDA(cudaMemsetAsync(lb, 0, (size_t)nbricks * sizeof(BrickRec), ctx->stream));
        band_mark_kernel<<<blocks_for(G.n_pts, 128), 128, 0, ctx->stream>>>(G, lb, R);
        PCR_LAUNCH_CHECK();
        PCR_CUDA(ctx->tmp_a.ensure((size_t)(nbricks + 1) * 4 * 2));
        uint32_t* bcnt = ctx->tmp_a.as<uint32_t>();
        uint32_t* bbase = bcnt + (nbricks + 1);
        PCR_CUDA(cudaMemsetAsync(bcnt, 0, (size_t)(nbricks + 1) * 4, ctx->stream));
        brick_popc_kernel<<<blocks_for((long long)nbricks, 256), 256, 0, ctx->stream>>>(lb, nbricks, bcnt);
        PCR_LAUNCH_CHECK();
        int rc = exclusive_sum_u32(ctx, bcnt, bbase, (long long)nbricks + 1);
        if (rc) return rc;
        uint32_t n_band = 0;
        PCR_CUDA(cudaMemcpyAsync(&n_band, bbase + nbricks, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        brick_base_set_kernel<<<blocks_for((long long)nbricks, 256), 256, 0, ctx->stream>>>(lb, nbricks, bbase);
        PCR_LAUNCH_CHECK();
        PCR_CUDA(ctx->tmp_b.ensure(((size_t)n_band + 1) * 4));
        PCR_CUDA(b_start.ensure(((size_t)n_band + 1) * 4));
        uint32_t* counts = ctx->tmp_b.as<uint32_t>();
        PCR_CUDA(cudaMemsetAsync(counts, 0, ((size_t)n_band + 1) * 4, ctx->stream));
        shell_build_kernel<false><<<blocks_for(nthreads, 128), 128, 0, ctx->stream>>>(G, lb, nbricks, dmax, R, counts, nullptr, nullptr, nullptr);
        PCR_LAUNCH_CHECK();
        // total size in 64 bits: the 32-bit offsets below must not wrap, and the lists must fit the cap
        size_t tmp = 0;
        PCR_CUDA(ctx->tmp_e.ensure(64));
        unsigned long long* d_total = ctx->tmp_e.as<unsigned long long>();
        cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> counts64(counts, U32ToU64());   // 64-bit accumulation
        PCR_CUDA(cub::DeviceReduce::Sum(nullptr, tmp, counts64, d_total, (long long)n_band, ctx->stream));
        PCR_CUDA(ctx->cub_tmp.ensure(tmp));
        PCR_CUDA(cub::DeviceReduce::Sum(ctx->cub_tmp.p, tmp, counts64, d_total, (long long)n_band, ctx->stream));
        ctx->launches += 1;
        unsigned long long total_groups = 0;
        PCR_CUDA(cudaMemcpyAsync(&total_groups, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        const unsigned long long total = 4ull * total_groups;                         // entries
        const double gib = (double)total * 17.0 / (1024.0 * 1024.0 * 1024.0);
        // the margin of three cells (2.5x the memory of two) is for targets whose lists stay small
        const double cap_gib = frac > 2.0 ? std::min(ctx->shell_max_gib, ctx->shell_wide_gib) : ctx->shell_max_gib;
        if (total_groups >= (1ull << 32) - 8ull || gib > cap_gib) continue;           // too large: try a smaller margin
        rc = exclusive_sum_u32(ctx, counts, b_start.as<uint32_t>(), (long long)n_band + 1);
        if (rc) return rc;
        const unsigned long long n_entries = total;
        PCR_CUDA(b_pts.ensure(((size_t)n_entries + 4) * sizeof(float4)));
        PCR_CUDA(b_margin2.ensure(((size_t)total_groups + 1) * 4));
        pad_tail_kernel<<<1, 32, 0, ctx->stream>>>(b_pts.as<float4>(), (long long)n_entries, nullptr, 0);
        PCR_LAUNCH_CHECK();
        shell_build_kernel<true><<<blocks_for(nthreads, 128), 128, 0, ctx->stream>>>(G, lb, nbricks, dmax, R, nullptr, b_start.as<uint32_t>(),
                                                                                    b_pts.as<float4>(), b_margin2.as<float>());
        PCR_LAUNCH_CHECK();
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        view.bricks = b_bricks.as<uint4>();
        view.start = b_start.as<uint32_t>();
        view.pts = b_pts.as<float4>();
        view.margin2 = b_margin2.as<float>();
        const float cov = fmaxf(dmax - G.slack * G.h, 0.0f);
        view.covered2 = cov * cov;
        view.block_r = (double)dmax >= 1.7320508 * (double)G.h * 1.0001 ? 1 : 0;
        n_band_out = n_band;
        n_entries_out = (long long)n_entries;
        dmax_used = frac;
        return PCR_OK;
    }
    return PCR_OK;                                            // nothing fits: the general search stays in charge
}

template <typename T>
static int build_voxels_impl(pcr_ctx* ctx, const void* xyz, long long n, double voxel_size, int min_points, int with_icov) {
    VoxelFront F;
    int rc = voxel_front<T>(ctx, xyz, n, voxel_size, min_points, F);
    if (rc) { F.release(); return rc; }
    uint32_t* keep = F.keep.as<uint32_t>();
    uint32_t* kord = keep + F.n_seg;
    rc = exclusive_sum_u32(ctx, keep, kord, F.n_seg);
    if (rc) { F.release(); return rc; }
    uint32_t n_keep = 0;
    rc = read_segment_count(ctx, keep, kord, F.n_seg, &n_keep);
    if (rc) { F.release(); return rc; }

    ctx->has_voxels = false; ctx->has_icov = false;
    ctx->vox_grid_epoch++;
    ctx->vox_grid.release();
    ctx->tile_vox.release();
    ctx->vox_shell = ShellLists{};
    ctx->n_vox = n_keep; ctx->n_vox_all = F.n_seg; ctx->voxel_size = voxel_size;
    const size_t nk = n_keep ? n_keep : 1;
    PCR_CUDA(ctx->vox_mean.ensure(nk * 3 * 8));
    PCR_CUDA(ctx->vox_cov.ensure(nk * 9 * 8));
    PCR_CUDA(ctx->vox_norm.ensure(nk * 3 * 8));
    PCR_CUDA(ctx->vox_icov.ensure(nk * 9 * 8));
    PCR_CUDA(ctx->vox_count.ensure(nk * 8));
    PCR_CUDA(ctx->vox_rec_plane.ensure(nk * 2 * sizeof(float4)));
    PCR_CUDA(ctx->vox_rec_ndt.ensure(nk * 3 * sizeof(float4)));
    Grid& g = ctx->vox_grid;
    const unsigned long long nbricks = (unsigned long long)F.bn[0] * F.bn[1] * F.bn[2];
    PCR_CUDA(g.bricks.ensure((size_t)nbricks * sizeof(BrickRec)));
    PCR_CUDA(g.cell_start.ensure((nk + 1) * 4));
    PCR_CUDA(g.pts.ensure((nk + 4) * sizeof(float4)));
    pad_tail_kernel<<<1, 32, 0, ctx->stream>>>(g.pts.as<float4>(), (long long)n_keep, nullptr, 0);
    PCR_LAUNCH_CHECK();
    brick_base_init_kernel<<<blocks_for((long long)nbricks, 256), 256, 0, ctx->stream>>>(g.bricks.as<BrickRec>(), nbricks);
    PCR_LAUNCH_CHECK();
    voxel_finalize_kernel<<<blocks_for(F.n_seg, 128), 128, 0, ctx->stream>>>(
        ctx->tmp_b.as<unsigned long long>(), F.seg.as<uint32_t>(), keep, kord, F.n_seg, n_keep, F.mean_all.as<double>(),
        F.cov6_all.as<double>(), with_icov, ctx->vox_mean.as<double>(), ctx->vox_cov.as<double>(), ctx->vox_norm.as<double>(),
        ctx->vox_icov.as<double>(), ctx->vox_count.as<long long>(), g.bricks.as<BrickRec>(), g.cell_start.as<uint32_t>(),
        g.pts.as<float4>(), ctx->vox_rec_plane.as<float4>(), ctx->vox_rec_ndt.as<float4>());
    PCR_LAUNCH_CHECK();
    GridView V{};
    V.h = (float)voxel_size;
    V.inv_h = (float)(1.0 / voxel_size);
    V.ox = (float)((double)F.mn[0] * voxel_size); V.oy = (float)((double)F.mn[1] * voxel_size); V.oz = (float)((double)F.mn[2] * voxel_size);
    V.bnx = F.bn[0]; V.bny = F.bn[1]; V.bnz = F.bn[2];
    V.cnx = V.bnx * 4; V.cny = V.bny * 4; V.cnz = V.bnz * 4;
    double maxabs = 0;
    for (int a = 0; a < 3; ++a)
        maxabs = std::max(maxabs, std::max(fabs((double)F.mn[a]), fabs((double)F.mn[a] + 4.0 * F.bn[a])) * voxel_size);
    V.slack = 1e-3f + 1e-6f * (float)(2.0 * maxabs / voxel_size + (double)std::max(V.cnx, std::max(V.cny, V.cnz)));
    V.n_pts = n_keep;
    V.bricks = g.bricks.as<uint4>();
    V.cell_start = g.cell_start.as<uint32_t>();
    V.pts = g.pts.as<float4>();
    g.view = V;
    g.n_cells = n_keep;
    g.built = true;
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    F.release();
    if (ctx->use_tile) rc = ensure_tile_index(ctx, PCR_VPLANE);   // row grid over the kept means + both payloads
    else rc = build_voxel_lists(ctx);                             // exact candidate lists (VoxelGrid.query)
    if (rc) return rc;
    if (!ctx->use_tile) {
        rc = build_shell_lists(ctx, 1);                           // margin-ordered lists streamed by the correspondence pass
        if (rc) return rc;
    }
    ctx->has_voxels = true;
    ctx->has_icov = with_icov != 0;
    return PCR_OK;
}

// ---------------------------------------------------------------------------------------
// row grid + row-major point copy for the tile-stream kernel (pcr_tile.cuh)
// ---------------------------------------------------------------------------------------
__global__ void bbox4_kernel(const float4* __restrict__ pts, long long n, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
            int a = f2ord(p.x), b = f2ord(p.y), c = f2ord(p.z);
            lo[0] = min(lo[0], a); hi[0] = max(hi[0], a);
            lo[1] = min(lo[1], b); hi[1] = max(hi[1], b);
            lo[2] = min(lo[2], c); hi[2] = max(hi[2], c);
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

// cell number (x fastest) of every source point + histogram of the cells
__global__ void tile_key_kernel(const float4* __restrict__ src, long long n, TileGrid G, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals, uint32_t* __restrict__ hist) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = src[i];
    float gx = (p.x - G.ox) * G.inv_c, gy = (p.y - G.oy) * G.inv_c, gz = (p.z - G.oz) * G.inv_c;
    if (!(gx == gx)) gx = 0.f;
    if (!(gy == gy)) gy = 0.f;
    if (!(gz == gz)) gz = 0.f;
    const float big = 1.0e9f;
    const int cx = cell_of(fminf(fmaxf(gx, -big), big), G.nx);
    const int cy = cell_of(fminf(fmaxf(gy, -big), big), G.ny);
    const int cz = cell_of(fminf(fmaxf(gz, -big), big), G.nz);
    const uint32_t key = (uint32_t)(((size_t)cz * G.ny + cy) * G.nx + cx);
    keys[i] = key;
    vals[i] = (uint32_t)i;
    atomicAdd(&hist[key], 1u);
}

__global__ void tile_count_occupied_kernel(const uint32_t* __restrict__ hist, unsigned long long ncells, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < ncells; i += (unsigned long long)gridDim.x * blockDim.x)
        c += hist[i] != 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// sorted points as PAIR records (see TileGrid::pairs): record p = points 2p, 2p+1 as
// (x0, x1, y0, y1), (z0, z1, w0, w1); an odd tail gets a sentinel that can never win
__global__ void tile_gather_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ vals, long long n,
                                   float* __restrict__ pairs, uint32_t* __restrict__ perm) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n_even = (n + 1) / 2 * 2;
    if (i >= n_even) return;
    float4 p = make_float4(3.0e38f, 3.0e38f, 3.0e38f, __uint_as_float(0xffffffffu));
    if (i < n) {
        const uint32_t j = vals[i];
        const float4 t = src[j];
        p = make_float4(t.x, t.y, t.z, __uint_as_float((uint32_t)i));
        perm[i] = j;
    }
    float* rec = pairs + (i >> 1) * 8 + (i & 1);
    rec[0] = p.x; rec[2] = p.y; rec[4] = p.z; rec[6] = p.w;
}

// payload in row-grid order: mode 0 = one float4 per source point (normals in source order);
// mode 1 = VPlaneICP record (2 float4 / voxel: mean, normal) -> normal; mode 2 = NDT record
// (3 float4 / voxel: (mean, W00), (W01, W02, W11, W12), (W22, ..)) -> (W00, W01, W02, W11), (W12, W22, 0, 0)
__global__ void tile_payload_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ perm, long long n, int mode,
                                    float4* __restrict__ pay) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t j = perm[i];
    if (mode == 0) {
        pay[i] = src[j];
    } else if (mode == 1) {
        pay[i] = src[2 * j + 1];
    } else {
        const float4 a = src[3 * j], b = src[3 * j + 1], c = src[3 * j + 2];
        pay[2 * i] = make_float4(a.w, b.x, b.y, b.z);
        pay[2 * i + 1] = make_float4(b.w, c.x, 0.f, 0.f);
    }
}

constexpr unsigned long long kMaxTileCells = 1ull << 29;   // 2 GiB of cell starts

// Build the row grid over n float4 source points (xyz + anything); ppc = desired mean points per
// occupied cell.  out.perm maps a position of the new order to the index in `src`.
static int build_tile_index(pcr_ctx* ctx, const float4* src, long long n, double ppc, TileIndex& out) {
    out.release();
    TileGrid V{};
    if (n <= 0) {
        // empty index (e.g. no voxel survived min_points): one empty cell, every query misses
        V.ox = V.oy = V.oz = 0.f; V.c = 1.f; V.inv_c = 1.f; V.inv_c2 = 1.f; V.slack = 1e-3f;
        V.nx = V.ny = V.nz = 1; V.n = 0;
        PCR_CUDA(out.cs.ensure(2 * 4));
        PCR_CUDA(cudaMemsetAsync(out.cs.p, 0, 8, ctx->stream));
        PCR_CUDA(out.pts.ensure(32));
        PCR_CUDA(out.perm.ensure(4));
        V.cs = out.cs.as<uint32_t>(); V.pairs = out.pts.as<float4>();
        out.view = V; out.built = true; out.n_cells_occupied = 0;
        return PCR_OK;
    }
    if (n >= (1ll << 31)) return fail(ctx, PCR_ERR_LIMIT, "point count exceeds 2^31-1");
    PCR_CUDA(ctx->tmp_e.ensure(64));
    int* d_mm = ctx->tmp_e.as<int>();
    init_minmax_kernel<<<1, 32, 0, ctx->stream>>>(d_mm);
    PCR_LAUNCH_CHECK();
    bbox4_kernel<<<min(blocks_for(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(src, n, d_mm);
    PCR_LAUNCH_CHECK();
    int h_mm[6];
    PCR_CUDA(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_mm[0] == INT_MAX) return fail(ctx, PCR_ERR_ARG, "point set has no finite points");
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = ord2f_host(h_mm[a]); hi[a] = ord2f_host(h_mm[3 + a]); }
    double ext[3], maxext = 0, maxabs = 0;
    for (int a = 0; a < 3; ++a) {
        ext[a] = (double)hi[a] - (double)lo[a];
        maxext = std::max(maxext, ext[a]);
        maxabs = std::max(maxabs, std::max(fabs((double)lo[a]), fabs((double)hi[a])));
    }
    if (maxext <= 0) maxext = 1.0;
    double vol = 1.0;
    for (int a = 0; a < 3; ++a) vol *= std::max(ext[a], 1e-3 * maxext);
    double c = cbrt(vol * ppc / (double)n);
    c = std::max(c, maxext * 1e-5);

    PCR_CUDA(ctx->tmp_a.ensure((size_t)n * 8));
    PCR_CUDA(ctx->tmp_b.ensure((size_t)n * 8));
    uint32_t* k_in = ctx->tmp_a.as<uint32_t>(); uint32_t* k_out = k_in + n;
    uint32_t* v_in = ctx->tmp_b.as<uint32_t>(); uint32_t* v_out = v_in + n;
    const int max_attempts = 4;
    for (int attempt = 0; attempt < max_attempts; ++attempt) {
        unsigned long long ncells;
        for (;;) {   // enlarge the cell until the dense table fits
            V.c = (float)c; V.inv_c = (float)(1.0 / c); V.inv_c2 = V.inv_c * V.inv_c;
            // one and a half empty cells around the bounding box: no indexed point is ever clamped into a border cell
            V.ox = lo[0] - 1.5f * V.c; V.oy = lo[1] - 1.5f * V.c; V.oz = lo[2] - 1.5f * V.c;
            double d[3];
            for (int a = 0; a < 3; ++a) d[a] = floor(((double)hi[a] - (double)(lo[a] - 1.5f * V.c)) / c) + 3.0;
            if (d[0] < 2.0e6 && d[1] < 2.0e6 && d[2] < 2.0e6 && d[0] * d[1] * d[2] <= (double)kMaxTileCells) {
                V.nx = (int)d[0]; V.ny = (int)d[1]; V.nz = (int)d[2];
                break;
            }
            c *= 1.2599210498948732;
        }
        ncells = (unsigned long long)V.nx * V.ny * V.nz;
        V.slack = 1e-3f + 1e-6f * (float)(2.0 * maxabs / c + (double)std::max(V.nx, std::max(V.ny, V.nz)));
        V.n = (uint32_t)n;
        PCR_CUDA(out.cs.ensure((size_t)(ncells + 1) * 4));
        PCR_CUDA(cudaMemsetAsync(out.cs.p, 0, (size_t)(ncells + 1) * 4, ctx->stream));
        PCR_CUDA(ctx->tile_scratch.ensure(64));
        PCR_CUDA(cudaMemsetAsync(ctx->tile_scratch.p, 0, 64, ctx->stream));
        tile_key_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(src, n, V, k_in, v_in, out.cs.as<uint32_t>());
        PCR_LAUNCH_CHECK();
        tile_count_occupied_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(out.cs.as<uint32_t>(), ncells, ctx->tile_scratch.as<unsigned long long>());
        PCR_LAUNCH_CHECK();
        unsigned long long occ = 0;
        PCR_CUDA(cudaMemcpyAsync(&occ, ctx->tile_scratch.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        const double got = (double)n / (double)std::max<unsigned long long>(occ, 1ull);
        if (attempt + 1 < max_attempts && (got > 1.6 * ppc || (got < ppc / 1.6 && occ > 64))) {
            double ratio = sqrt(ppc / got);                 // points lie on 2-D surfaces: ppc ~ c^2
            ratio = std::min(std::max(ratio, 0.25), 4.0);
            c *= ratio;
            continue;
        }
        out.n_cells_occupied = (long long)occ;
        // ---- accept: sort by cell number, cell starts = exclusive prefix sum of the histogram ----
        size_t tmp = 0;
        const int end_bit = bits_for(ncells);
        PCR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, n, 0, end_bit, ctx->stream));
        PCR_CUDA(ctx->cub_tmp.ensure(tmp));
        PCR_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, k_in, k_out, v_in, v_out, n, 0, end_bit, ctx->stream));
        ctx->launches += 4;
        int rc = exclusive_sum_u32(ctx, out.cs.as<uint32_t>(), out.cs.as<uint32_t>(), (long long)(ncells + 1));
        if (rc) return rc;
        PCR_CUDA(out.pts.ensure((size_t)(n + 1) * sizeof(float4)));
        PCR_CUDA(out.perm.ensure((size_t)n * 4));
        tile_gather_kernel<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(src, v_out, n, out.pts.as<float>(), out.perm.as<uint32_t>());
        PCR_LAUNCH_CHECK();
        V.cs = out.cs.as<uint32_t>();
        V.pairs = out.pts.as<float4>();
        out.view = V;
        out.built = true;
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        return PCR_OK;
    }
    return fail(ctx, PCR_ERR_LIMIT, "row grid build did not converge");
}

static int build_tile_payload(pcr_ctx* ctx, TileIndex& t, const float4* src, int mode, DevBuf& pay) {
    const long long n = t.view.n;
    const size_t per = mode == 2 ? 2 : 1;
    PCR_CUDA(pay.ensure(std::max<size_t>((size_t)n * per, 1) * sizeof(float4)));
    if (n > 0) {
        tile_payload_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(src, t.perm.as<uint32_t>(), n, mode, pay.as<float4>());
        PCR_LAUNCH_CHECK();
    }
    return PCR_OK;
}

// Build / refresh what the tile-stream kernel needs for `method` (called by the set_target entry
// points and, as a safety net, by the first linearisation).
int ensure_tile_index(pcr_ctx* ctx, int method) {
    if (method == PCR_ICP || method == PCR_PLANE) {
        if (!ctx->tgt_grid.built) return fail(ctx, PCR_ERR_STATE, "row grid: target NN index not built");
        if (!ctx->tile_tgt.built) {
            int rc = build_tile_index(ctx, ctx->tgt_grid.view.pts, ctx->n_tgt, ctx->tile_ppc_tgt, ctx->tile_tgt);
            if (rc) return rc;
        }
        if (method == PCR_PLANE && ctx->has_normals && ctx->tile_tgt.pay_epoch != ctx->normals_epoch) {
            int rc = build_tile_payload(ctx, ctx->tile_tgt, ctx->tgt_nrm_sorted.as<float4>(), 0, ctx->tile_tgt.pay);
            if (rc) return rc;
            ctx->tile_tgt.pay_epoch = ctx->normals_epoch;
            PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    } else {
        if (!ctx->vox_grid.built) return fail(ctx, PCR_ERR_STATE, "row grid: voxels not built");
        if (!ctx->tile_vox.built) {
            int rc = build_tile_index(ctx, ctx->vox_grid.view.pts, ctx->n_vox, ctx->tile_ppc_vox, ctx->tile_vox);
            if (rc) return rc;
            rc = build_tile_payload(ctx, ctx->tile_vox, ctx->vox_rec_plane.as<float4>(), 1, ctx->tile_vox.pay);
            if (rc) return rc;
            rc = build_tile_payload(ctx, ctx->tile_vox, ctx->vox_rec_ndt.as<float4>(), 2, ctx->tile_vox.pay2);
            if (rc) return rc;
            PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    return PCR_OK;
}

template <typename T>
static int voxel_filter_impl(pcr_ctx* ctx, const void* xyz, long long n, double voxel_size, float* out, long long* n_out) {
    VoxelFront F;
    int rc = voxel_front<T>(ctx, xyz, n, voxel_size, 0, F);
    if (rc) { F.release(); return rc; }
    DevBuf d_out;
    PCR_CUDA(d_out.ensure((size_t)F.n_seg * 12));
    voxel_centroid_kernel<<<blocks_for(F.n_seg, 256), 256, 0, ctx->stream>>>(F.mean_all.as<double>(), F.n_seg, d_out.as<float>());
    PCR_LAUNCH_CHECK();
    const bool dev_out = is_device_pointer(out);
    PCR_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)F.n_seg * 12, dev_out ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_out = F.n_seg;
    d_out.release();
    F.release();
    return PCR_OK;
}

// per point (caller order): ordinal of its voxel; per voxel: its integer coordinate floor(p / size)
__global__ void voxel_label_kernel(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ ords,
                                   const int* __restrict__ coords, long long n, long long* __restrict__ labels, int* __restrict__ gcoords) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t j = vals[s], grp = ords[s] + flags[s] - 1u;
    labels[j] = (long long)grp;
    if (flags[s]) { gcoords[3 * (size_t)grp] = coords[3 * (size_t)j]; gcoords[3 * (size_t)grp + 1] = coords[3 * (size_t)j + 1]; gcoords[3 * (size_t)grp + 2] = coords[3 * (size_t)j + 2]; }
}

template <typename T>
static int voxel_labels_impl(pcr_ctx* ctx, const void* xyz, long long n, double voxel_size, long long* labels, int* gcoords, long long* n_groups) {
    VoxelFront F;
    int rc = voxel_front<T>(ctx, xyz, n, voxel_size, 0, F);
    if (rc) { F.release(); return rc; }
    // buffers left behind by voxel_front: sorted original indices, head flags, head ordinals, coordinates
    const uint32_t* v_out = ctx->tmp_c.as<uint32_t>() + n;
    const uint32_t* flags = ctx->tmp_d.as<uint32_t>();
    const uint32_t* ords = flags + n;
    const int* coords = ctx->tmp_e.as<int>() + 16;
    DevBuf d_lab, d_gc;
    PCR_CUDA(d_lab.ensure((size_t)n * 8));
    PCR_CUDA(d_gc.ensure((size_t)F.n_seg * 12));
    voxel_label_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(v_out, flags, ords, coords, n, d_lab.as<long long>(), d_gc.as<int>());
    PCR_LAUNCH_CHECK();
    PCR_CUDA(cudaMemcpyAsync(labels, d_lab.p, (size_t)n * 8, is_device_pointer(labels) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaMemcpyAsync(gcoords, d_gc.p, (size_t)F.n_seg * 12, is_device_pointer(gcoords) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_groups = F.n_seg;
    d_lab.release(); d_gc.release();
    F.release();
    return PCR_OK;
}

}  // namespace pcr

using namespace pcr;

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int pcr_version(void) { return 100; }

int pcr_device_count(int* count) {
    pcr_ctx* ctx = nullptr;
    PCR_CUDA(cudaGetDeviceCount(count));
    return PCR_OK;
}

const char* pcr_last_error(const pcr_ctx* ctx) { return ctx ? ctx->err.c_str() : global_error(); }

int pcr_create(int device_id, pcr_ctx** out) {
    if (!out) return fail(nullptr, PCR_ERR_ARG, "pcr_create: out is NULL");
    *out = nullptr;
    pcr_ctx* ctx = nullptr;
    int ndev = 0;
    PCR_CUDA(cudaGetDeviceCount(&ndev));
    if (device_id < 0 || device_id >= ndev)
        return fail(nullptr, PCR_ERR_ARG, "pcr_create: no CUDA device " + std::to_string(device_id) + " (" + std::to_string(ndev) + " visible)");
    PCR_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    PCR_CUDA(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10)
        return fail(nullptr, PCR_ERR_CUDA, std::string("pcr_create: device '") + prop.name + "' is not sm_100+; this library ships sm_100a code only");
    ctx = new pcr_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        std::string m = std::string("pcr_create: ") + cudaGetErrorString(e);
        delete ctx;
        return fail(nullptr, PCR_ERR_CUDA, m);
    }
    if (const char* e = getenv("PCR_MIN_BLOCKS")) ctx->min_blocks = atoi(e) >= 3 && atoi(e) <= 6 ? atoi(e) : 0;
    if (const char* e = getenv("PCR_TARGET_PPC")) ctx->target_ppc = atof(e) > 0.5 ? atof(e) : 0.0;
    if (const char* e = getenv("PCR_SHELL_DMAX")) ctx->shell_dmax_frac = atof(e) > 0.0 && atof(e) <= PCR_SHELL_MAX_MARGIN ? atof(e) : 1.0;
    if (const char* e = getenv("PCR_SHELL_MAX_GIB")) ctx->shell_max_gib = atof(e) > 0.0 ? atof(e) : 96.0;
    if (const char* e = getenv("PCR_SHELL_WIDE_GIB")) ctx->shell_wide_gib = atof(e) >= 0.0 ? atof(e) : 8.0;
    if (const char* e = getenv("PCR_LIST_DILATE")) ctx->list_dilate = atoi(e) >= 1 && atoi(e) <= 4 ? atoi(e) : 3;
    if (const char* e = getenv("PCR_LIST_RADIUS")) ctx->list_radius = atoi(e) >= 2 && atoi(e) <= 6 ? atoi(e) : 5;
    if (const char* e = getenv("PCR_BALL_FIRST")) ctx->ball_first = atoi(e) != 0;
    if (const char* e = getenv("PCR_CELL_ORDER")) ctx->cell_order = atoi(e) != 0;
    if (const char* e = getenv("PCR_GRAB_ROWS")) ctx->grab_rows = atoi(e) >= 0 && atoi(e) <= 64 ? atoi(e) : 0;
    if (const char* e = getenv("PCR_SPLIT")) ctx->split_passes = atoi(e) != 0;
    if (const char* e = getenv("PCR_PATH")) ctx->use_tile = strcmp(e, "tile") == 0;
    if (const char* e = getenv("PCR_TILE_PPC")) ctx->tile_ppc_tgt = atof(e) > 0.25 ? atof(e) : 8.0;
    if (const char* e = getenv("PCR_TILE_PPC_VOX")) ctx->tile_ppc_vox = atof(e) > 0.25 ? atof(e) : 4.0;
    if (const char* e = getenv("PCR_TILE_CAP")) ctx->tile_cap = atoi(e) >= 256 && atoi(e) <= 4096 ? atoi(e) / 64 * 64 : 384;
    if (const char* e = getenv("PCR_TILE_CORE")) ctx->tile_core_e = atoi(e) >= 0 && atoi(e) <= 16 ? atoi(e) : 8;
    if (const char* e = getenv("PCR_TILE_MINB")) ctx->tile_min_blocks = atoi(e) >= 3 && atoi(e) <= 6 ? atoi(e) : 0;
    if (const char* e = getenv("PCR_TILE_R0")) ctx->tile_first_radius = atof(e) > 0.0 ? (float)atof(e) : 0.5f;
    if (const char* e = getenv("PCR_TILE_G")) ctx->tile_groups = (atoi(e) == 1 || atoi(e) == 2 || atoi(e) == 4 || atoi(e) == 8) ? atoi(e) : 1;
    if (const char* e = getenv("PCR_ACC_MINB")) ctx->acc_min_blocks = atoi(e) == 3 ? 3 : 2;
    if (const char* e = getenv("PCR_TILE_BULK_MIN")) ctx->tile_bulk_min = atoi(e) >= 1 ? atoi(e) : 9;
    if (const char* e = getenv("PCR_ORDER_REUSE")) ctx->order_reuse = atoi(e) != 0;
    if (const char* e = getenv("PCR_E2E_CHUNKS")) ctx->host_chunks = atoi(e) >= 1 && atoi(e) <= 8 ? atoi(e) : 4;
    if (const char* e = getenv("PCR_E2E_MIN_CHUNK")) ctx->host_chunk_min_reuse = atoll(e) >= 64 ? atoll(e) : 2000000;
    if (const char* e = getenv("PCR_TILE_SPLIT")) ctx->tile_split = atoi(e) != 0;
    if (const char* e = getenv("PCR_TILE_KR")) ctx->tile_rows_per_unit = atoi(e) == 2 || atoi(e) == 4 ? atoi(e) : 0;
    int rc = ensure_loop_buffers(ctx);
    if (rc) { std::string m = ctx->err; pcr_destroy(ctx); return fail(nullptr, rc, m); }
    *out = ctx;
    return PCR_OK;
}

int pcr_destroy(pcr_ctx* ctx) {
    if (!ctx) return PCR_OK;
    cudaSetDevice(ctx->device);
    pcr_comm_destroy(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->tgt_xyz.release(); ctx->tgt_grid.release(); ctx->tgt_nrm_sorted.release(); ctx->tgt_nrm_orig.release(); ctx->tgt_pn.release();
    ctx->vox_mean.release(); ctx->vox_cov.release(); ctx->vox_norm.release(); ctx->vox_icov.release(); ctx->vox_count.release();
    ctx->vox_grid.release(); ctx->vox_rec_plane.release(); ctx->vox_rec_ndt.release();
    ctx->vox_lbricks.release(); ctx->vox_list_start.release(); ctx->vox_list_idx.release(); ctx->vox_list_pts.release();
    ctx->shell_bricks.release(); ctx->shell_start.release(); ctx->shell_pts.release(); ctx->shell_margin2.release();
    ctx->vshell_bricks.release(); ctx->vshell_start.release(); ctx->vshell_pts.release(); ctx->vshell_margin2.release();
    ctx->scan_x.release(); ctx->scan_y.release(); ctx->scan_z.release(); ctx->scan_raw.release(); ctx->scan_prev.release();
    ctx->tile_tgt.release(); ctx->tile_vox.release(); ctx->scan_hint.release(); ctx->tile_scratch.release();
    ctx->scan_order.release();
    ctx->partials.release(); ctx->state.release();
    ctx->tmp_a.release(); ctx->tmp_b.release(); ctx->tmp_c.release(); ctx->tmp_d.release(); ctx->tmp_e.release(); ctx->cub_tmp.release();
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    for (cudaEvent_t e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PCR_OK;
}

int pcr_set_target_points(pcr_ctx* ctx, const float* xyz, int64_t n) {
    if (!ctx) return PCR_ERR_ARG;
    if (!xyz || n <= 0) return fail(ctx, PCR_ERR_ARG, "pcr_set_target_points: empty target");
    PCR_CUDA(cudaSetDevice(ctx->device));
    ctx->tgt_grid.release();
    ctx->tile_tgt.release();
    ctx->tgt_shell = ShellLists{};              // lists refer to the released grid
    ctx->n_shell_band = ctx->n_shell_entries = 0;
    ctx->shell_tried = false;
    ctx->has_normals = false;
    PCR_CUDA(ctx->tgt_xyz.ensure((size_t)n * 12));
    PCR_CUDA(cudaMemcpyAsync(ctx->tgt_xyz.p, xyz, (size_t)n * 12,
                             is_device_pointer(xyz) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->n_tgt = n;
    return PCR_OK;
}

int pcr_append_target_points(pcr_ctx* ctx, const float* xyz, int64_t n) {
    if (!ctx) return PCR_ERR_ARG;
    if (!xyz || n <= 0) return fail(ctx, PCR_ERR_ARG, "pcr_append_target_points: nothing to append");
    if (ctx->n_tgt <= 0) return fail(ctx, PCR_ERR_STATE, "pcr_append_target_points: no target to append to");
    PCR_CUDA(cudaSetDevice(ctx->device));
    const long long n_old = ctx->n_tgt, n_new = n_old + n;
    if (n_new >= (1ll << 31)) return fail(ctx, PCR_ERR_LIMIT, "point count exceeds 2^31-1");
    pcr::DevBuf grown;
    PCR_CUDA(grown.ensure((size_t)n_new * 12));
    PCR_CUDA(cudaMemcpyAsync(grown.p, ctx->tgt_xyz.p, (size_t)n_old * 12, cudaMemcpyDeviceToDevice, ctx->stream));   // the old part never leaves the GPU
    PCR_CUDA(cudaMemcpyAsync((char*)grown.p + (size_t)n_old * 12, xyz, (size_t)n * 12,
                             is_device_pointer(xyz) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->tgt_xyz.release();
    ctx->tgt_xyz = grown;
    ctx->n_tgt = n_new;
    // every structure over the old target is void (rebuilt by the usual build calls)
    ctx->tgt_grid.release();
    ctx->tile_tgt.release();
    ctx->tgt_shell = ShellLists{};
    ctx->n_shell_band = ctx->n_shell_entries = 0;
    ctx->shell_tried = false;
    ctx->has_normals = false;
    return PCR_OK;
}

int pcr_build_nn_index(pcr_ctx* ctx) {
    if (!ctx) return PCR_ERR_ARG;
    if (ctx->n_tgt <= 0) return fail(ctx, PCR_ERR_STATE, "pcr_build_nn_index: target points not set");
    PCR_CUDA(cudaSetDevice(ctx->device));
    ctx->tgt_grid_epoch++;
    ctx->tile_tgt.release();
    ctx->tgt_shell = ShellLists{};              // lists of the previous grid are void; rebuilt on demand
    ctx->n_shell_band = ctx->n_shell_entries = 0;
    ctx->shell_dmax_used = 0.0;
    ctx->shell_tried = false;
    return build_point_grid(ctx, ctx->tgt_xyz.as<float>(), ctx->n_tgt, ctx->tgt_grid, nullptr);
}

int pcr_build_correspondence_lists(pcr_ctx* ctx) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->tgt_grid.built) return fail(ctx, PCR_ERR_STATE, "pcr_build_correspondence_lists: NN index not built");
    PCR_CUDA(cudaSetDevice(ctx->device));
    ctx->shell_tried = true;
    if (ctx->use_tile) return ensure_tile_index(ctx, ctx->has_normals ? PCR_PLANE : PCR_ICP);   // row grid (+ normals in its order)
    return build_shell_lists(ctx);
}

int pcr_estimate_normals(pcr_ctx* ctx, int k) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->tgt_grid.built) return fail(ctx, PCR_ERR_STATE, "pcr_estimate_normals: NN index not built");
    if (k < 1 || k > 64) return fail(ctx, PCR_ERR_ARG, "pcr_estimate_normals: k must be in 1..64");
    PCR_CUDA(cudaSetDevice(ctx->device));
    const long long n = ctx->n_tgt;
    PCR_CUDA(ctx->tgt_nrm_sorted.ensure((size_t)n * sizeof(float4)));
    PCR_CUDA(ctx->tgt_nrm_orig.ensure((size_t)n * 12));
    const int thr = 128, blk = blocks_for(n, thr);
    float4* ns = ctx->tgt_nrm_sorted.as<float4>();
    float* no = ctx->tgt_nrm_orig.as<float>();
    if (k <= 8) normals_kernel<8><<<blk, thr, 0, ctx->stream>>>(ctx->tgt_grid.view, k, ns, no);
    else if (k <= 16) normals_kernel<16><<<blk, thr, 0, ctx->stream>>>(ctx->tgt_grid.view, k, ns, no);
    else if (k <= 32) normals_kernel<32><<<blk, thr, 0, ctx->stream>>>(ctx->tgt_grid.view, k, ns, no);
    else normals_kernel<64><<<blk, thr, 0, ctx->stream>>>(ctx->tgt_grid.view, k, ns, no);
    PCR_LAUNCH_CHECK();
    PCR_CUDA(ctx->tgt_pn.ensure((size_t)n * 2 * sizeof(float4)));
    interleave_pn_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ctx->tgt_grid.view.pts, ns, n, ctx->tgt_pn.as<float4>());
    PCR_LAUNCH_CHECK();
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->has_normals = true;
    ctx->normals_epoch++;
    if (ctx->use_tile && ctx->tile_tgt.built) return ensure_tile_index(ctx, PCR_PLANE);
    return PCR_OK;
}

int pcr_set_normals(pcr_ctx* ctx, const float* normals) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->tgt_grid.built) return fail(ctx, PCR_ERR_STATE, "pcr_set_normals: NN index not built");
    if (!normals) return fail(ctx, PCR_ERR_ARG, "pcr_set_normals: NULL normals");
    PCR_CUDA(cudaSetDevice(ctx->device));
    const long long n = ctx->n_tgt;
    PCR_CUDA(ctx->tgt_nrm_sorted.ensure((size_t)n * sizeof(float4)));
    PCR_CUDA(ctx->tgt_nrm_orig.ensure((size_t)n * 12));
    PCR_CUDA(cudaMemcpyAsync(ctx->tgt_nrm_orig.p, normals, (size_t)n * 12,
                             is_device_pointer(normals) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    scatter_normals_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ctx->tgt_nrm_orig.as<float>(), ctx->tgt_grid.view.pts, n,
                                                                        ctx->tgt_nrm_sorted.as<float4>());
    PCR_LAUNCH_CHECK();
    PCR_CUDA(ctx->tgt_pn.ensure((size_t)n * 2 * sizeof(float4)));
    interleave_pn_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ctx->tgt_grid.view.pts, ctx->tgt_nrm_sorted.as<float4>(), n,
                                                                      ctx->tgt_pn.as<float4>());
    PCR_LAUNCH_CHECK();
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->has_normals = true;
    ctx->normals_epoch++;
    if (ctx->use_tile && ctx->tile_tgt.built) return ensure_tile_index(ctx, PCR_PLANE);
    return PCR_OK;
}

int pcr_get_normals(pcr_ctx* ctx, float* normals) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->has_normals) return fail(ctx, PCR_ERR_STATE, "pcr_get_normals: no normals");
    PCR_CUDA(cudaSetDevice(ctx->device));
    PCR_CUDA(cudaMemcpyAsync(normals, ctx->tgt_nrm_orig.p, (size_t)ctx->n_tgt * 12,
                             is_device_pointer(normals) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    return PCR_OK;
}

int pcr_build_voxels(pcr_ctx* ctx, const void* xyz, int64_t n, int is_f64, double voxel_size, int min_points, int with_icov) {
    if (!ctx) return PCR_ERR_ARG;
    if (!xyz && n == 0 && ctx->n_tgt > 0) {                   // the resident target points (pcr_set_target_points / pcr_append_target_points)
        xyz = ctx->tgt_xyz.p; n = ctx->n_tgt; is_f64 = 0;
    }
    if (!xyz || n <= 0) return fail(ctx, PCR_ERR_ARG, "pcr_build_voxels: empty input");
    PCR_CUDA(cudaSetDevice(ctx->device));
    return is_f64 ? build_voxels_impl<double>(ctx, xyz, n, voxel_size, min_points, with_icov)
                  : build_voxels_impl<float>(ctx, xyz, n, voxel_size, min_points, with_icov);
}

int pcr_get_voxel_count(pcr_ctx* ctx, int64_t* n_kept, int64_t* n_occupied) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->has_voxels) return fail(ctx, PCR_ERR_STATE, "pcr_get_voxel_count: voxels not built");
    if (n_kept) *n_kept = ctx->n_vox;
    if (n_occupied) *n_occupied = ctx->n_vox_all;
    return PCR_OK;
}

int pcr_get_voxels(pcr_ctx* ctx, double* mean, double* cov, double* norm, double* icov, int64_t* count) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->has_voxels) return fail(ctx, PCR_ERR_STATE, "pcr_get_voxels: voxels not built");
    if (icov && !ctx->has_icov) return fail(ctx, PCR_ERR_STATE, "pcr_get_voxels: inverse covariances were not requested at build time");
    PCR_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->n_vox;
    if (n) {
        if (mean) PCR_CUDA(cudaMemcpyAsync(mean, ctx->vox_mean.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
        if (cov) PCR_CUDA(cudaMemcpyAsync(cov, ctx->vox_cov.p, n * 72, cudaMemcpyDeviceToHost, ctx->stream));
        if (norm) PCR_CUDA(cudaMemcpyAsync(norm, ctx->vox_norm.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
        if (icov) PCR_CUDA(cudaMemcpyAsync(icov, ctx->vox_icov.p, n * 72, cudaMemcpyDeviceToHost, ctx->stream));
        if (count) PCR_CUDA(cudaMemcpyAsync(count, ctx->vox_count.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    return PCR_OK;
}

int pcr_knn(pcr_ctx* ctx, const float* queries, int64_t m, int k, float* dist, int64_t* idx) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->tgt_grid.built) return fail(ctx, PCR_ERR_STATE, "pcr_knn: NN index not built");
    if (k < 1 || k > 64) return fail(ctx, PCR_ERR_ARG, "pcr_knn: k must be in 1..64");
    if (m <= 0) return PCR_OK;
    PCR_CUDA(cudaSetDevice(ctx->device));
    DevBuf dq, dd, di;
    const float* q = queries;
    if (!is_device_pointer(queries)) {
        PCR_CUDA(dq.ensure((size_t)m * 12));
        PCR_CUDA(cudaMemcpyAsync(dq.p, queries, (size_t)m * 12, cudaMemcpyHostToDevice, ctx->stream));
        q = dq.as<float>();
    }
    PCR_CUDA(dd.ensure((size_t)m * k * 4));
    PCR_CUDA(di.ensure((size_t)m * k * 8));
    const int thr = 128, blk = blocks_for(m, thr);
    const GridView& G = ctx->tgt_grid.view;
    if (k == 1) nn_query_kernel<<<blk, thr, 0, ctx->stream>>>(G, q, m, dd.as<float>(), di.as<long long>());
    else if (k <= 8) knn_query_kernel<8><<<blk, thr, 0, ctx->stream>>>(G, q, m, k, dd.as<float>(), di.as<long long>());
    else if (k <= 16) knn_query_kernel<16><<<blk, thr, 0, ctx->stream>>>(G, q, m, k, dd.as<float>(), di.as<long long>());
    else if (k <= 32) knn_query_kernel<32><<<blk, thr, 0, ctx->stream>>>(G, q, m, k, dd.as<float>(), di.as<long long>());
    else knn_query_kernel<64><<<blk, thr, 0, ctx->stream>>>(G, q, m, k, dd.as<float>(), di.as<long long>());
    PCR_LAUNCH_CHECK();
    PCR_CUDA(cudaMemcpyAsync(dist, dd.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaMemcpyAsync(idx, di.p, (size_t)m * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    dq.release(); dd.release(); di.release();
    return PCR_OK;
}

int pcr_voxel_query(pcr_ctx* ctx, const float* queries, int64_t m, int64_t* vidx, double* dist) {
    if (!ctx) return PCR_ERR_ARG;
    if (!ctx->has_voxels) return fail(ctx, PCR_ERR_STATE, "pcr_voxel_query: voxels not built");
    if (m <= 0) return PCR_OK;
    PCR_CUDA(cudaSetDevice(ctx->device));
    DevBuf dq, dd, di;
    const float* q = queries;
    if (!is_device_pointer(queries)) {
        PCR_CUDA(dq.ensure((size_t)m * 12));
        PCR_CUDA(cudaMemcpyAsync(dq.p, queries, (size_t)m * 12, cudaMemcpyHostToDevice, ctx->stream));
        q = dq.as<float>();
    }
    PCR_CUDA(dd.ensure((size_t)m * 8));
    PCR_CUDA(di.ensure((size_t)m * 8));
    voxel_query_kernel<<<blocks_for(m, 128), 128, 0, ctx->stream>>>(ctx->vox_grid.view, ctx->vox_lists,
                                                                    ctx->use_voxel_lists && ctx->vox_lists.bricks != nullptr, q, m,
                                                                    ctx->vox_mean.as<double>(), di.as<long long>(), dd.as<double>());
    PCR_LAUNCH_CHECK();
    PCR_CUDA(cudaMemcpyAsync(dist, dd.p, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaMemcpyAsync(vidx, di.p, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    dq.release(); dd.release(); di.release();
    return PCR_OK;
}

int pcr_voxel_filter(pcr_ctx* ctx, const void* xyz, int64_t n, int is_f64, double voxel_size, float* out, int64_t* n_out) {
    if (!ctx) return PCR_ERR_ARG;
    if (!xyz || !out || !n_out || n <= 0) return fail(ctx, PCR_ERR_ARG, "pcr_voxel_filter: bad arguments");
    PCR_CUDA(cudaSetDevice(ctx->device));
    long long no = 0;
    int rc = is_f64 ? voxel_filter_impl<double>(ctx, xyz, n, voxel_size, out, &no) : voxel_filter_impl<float>(ctx, xyz, n, voxel_size, out, &no);
    *n_out = no;
    return rc;
}

int pcr_voxel_labels(pcr_ctx* ctx, const void* xyz, int64_t n, int is_f64, double voxel_size, int64_t* labels, int32_t* coords, int64_t* n_voxels) {
    if (!ctx) return PCR_ERR_ARG;
    if (!xyz || !labels || !coords || !n_voxels || n <= 0) return fail(ctx, PCR_ERR_ARG, "pcr_voxel_labels: bad arguments");
    PCR_CUDA(cudaSetDevice(ctx->device));
    long long ng = 0;
    int rc = is_f64 ? voxel_labels_impl<double>(ctx, xyz, n, voxel_size, (long long*)labels, coords, &ng)
                    : voxel_labels_impl<float>(ctx, xyz, n, voxel_size, (long long*)labels, coords, &ng);
    *n_voxels = ng;
    return rc;
}

int pcr_index_stats(pcr_ctx* ctx, int which, double* cell_edge, int64_t* n_cells, int64_t* n_bricks, int64_t* n_points) {
    if (!ctx) return PCR_ERR_ARG;
    const Grid& g = which == 0 ? ctx->tgt_grid : ctx->vox_grid;
    if (!g.built) return fail(ctx, PCR_ERR_STATE, "pcr_index_stats: index not built");
    if (cell_edge) *cell_edge = g.view.h;
    if (n_cells) *n_cells = g.n_cells;
    if (n_bricks) *n_bricks = (int64_t)g.view.bnx * g.view.bny * g.view.bnz;
    if (n_points) *n_points = g.view.n_pts;
    return PCR_OK;
}

int pcr_set_voxel_lists(pcr_ctx* ctx, int enable) {
    if (!ctx) return PCR_ERR_ARG;
    ctx->use_voxel_lists = enable ? 1 : 0;
    return PCR_OK;
}

int pcr_set_shell_lists(pcr_ctx* ctx, int enable) {
    if (!ctx) return PCR_ERR_ARG;
    ctx->use_shell_lists = enable ? 1 : 0;
    return PCR_OK;
}

int pcr_shell_list_stats(pcr_ctx* ctx, int64_t* band_cells, int64_t* entries, double* margin_cells) {
    if (!ctx) return PCR_ERR_ARG;
    if (band_cells) *band_cells = ctx->n_shell_band;
    if (entries) *entries = ctx->n_shell_entries;
    if (margin_cells) *margin_cells = ctx->shell_dmax_used;
    return PCR_OK;
}

int pcr_voxel_list_stats(pcr_ctx* ctx, int64_t* band_cells, int64_t* entries) {
    if (!ctx) return PCR_ERR_ARG;
    if (band_cells) *band_cells = ctx->n_band_cells;
    if (entries) *entries = ctx->n_list_entries;
    return PCR_OK;
}

int pcr_voxel_shell_stats(pcr_ctx* ctx, int64_t* band_cells, int64_t* entries, double* margin_cells) {
    if (!ctx) return PCR_ERR_ARG;
    if (band_cells) *band_cells = ctx->n_vshell_band;
    if (entries) *entries = ctx->n_vshell_entries;
    if (margin_cells) *margin_cells = ctx->vshell_dmax_used;
    return PCR_OK;
}

int pcr_set_path(pcr_ctx* ctx, int path) {
    if (!ctx) return PCR_ERR_ARG;
    if (path != 0 && path != 1) return fail(ctx, PCR_ERR_ARG, "pcr_set_path: 0 = lists, 1 = tile stream");
    ctx->use_tile = path == 1;
    return PCR_OK;
}

int pcr_set_record_matches(pcr_ctx* ctx, int enable) {
    if (!ctx) return PCR_ERR_ARG;
    ctx->record_matches = enable ? 1 : 0;
    return PCR_OK;
}

int pcr_tile_stats(pcr_ctx* ctx, int which, double* cell_edge, int64_t* cells, int64_t* occupied, int64_t* bytes) {
    if (!ctx) return PCR_ERR_ARG;
    const TileIndex& t = which == 0 ? ctx->tile_tgt : ctx->tile_vox;
    if (!t.built) return fail(ctx, PCR_ERR_STATE, "pcr_tile_stats: row grid not built");
    if (cell_edge) *cell_edge = t.view.c;
    if (cells) *cells = (int64_t)t.view.nx * t.view.ny * t.view.nz;
    if (occupied) *occupied = t.n_cells_occupied;
    if (bytes) *bytes = (int64_t)(t.cs.bytes + t.pts.bytes + t.perm.bytes + t.pay.bytes + t.pay2.bytes);
    return PCR_OK;
}

int pcr_launch_count(pcr_ctx* ctx, int64_t* launches) {
    if (!ctx || !launches) return PCR_ERR_ARG;
    *launches = ctx->launches;
    return PCR_OK;
}

int pcr_stream(pcr_ctx* ctx, void** stream) {
    if (!ctx || !stream) return PCR_ERR_ARG;
    *stream = (void*)ctx->stream;
    return PCR_OK;
}

}  // extern "C"
