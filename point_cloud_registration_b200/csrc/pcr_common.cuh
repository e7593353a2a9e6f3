// Shared definitions for the B200 point-cloud-registration hot path.
//
// Everything marked PCR_HD is plain C++ that compiles both as device code (the product) and
// as host code (ONLY for tests/hostsim, which replays single-query / single-point logic on
// the CPU so that the index arithmetic can be debugged without a GPU).  The product path
// never executes the host instantiations.
#pragma once

#include <cstdint>
#include <cmath>
#include <vector_types.h>

#if defined(__CUDACC__)
#define PCR_HD __host__ __device__ __forceinline__
#else
#define PCR_HD inline
#endif

// ---- method ids (match include/pcr_b200.h) -------------------------------------------
#define PCR_METHOD_ICP 0
#define PCR_METHOD_PLANE 1
#define PCR_METHOD_VPLANE 2
#define PCR_METHOD_NDT 3

// Normal-equation record produced per linearisation:
//   [0..20]  upper triangle of H, row major (H00 H01 .. H05 H11 .. H55)
//   [21..26] g
//   [27]     e2
//   [28]     number of inlier correspondences
#define PCR_NEQ 29
#define PCR_NEQ_PAD 32

namespace pcr {

PCR_HD int popc64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

PCR_HD int ffs64(unsigned long long v) {   // 1-based index of least significant set bit, 0 if none
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v);
#else
    return __builtin_ffsll((long long)v);
#endif
}

// IEEE single-rounding float ops that must not be contracted into FMAs (used where the
// reference's NumPy float32 arithmetic is replayed operation by operation).
PCR_HD float fmul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
PCR_HD float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
PCR_HD float fsub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
PCR_HD float fdiv_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
PCR_HD double dmul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
PCR_HD double dadd_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
PCR_HD double dsub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    volatile double r = a - b; return r;
#endif
}
PCR_HD double ddiv_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}

// Squared distance with ONE fixed rounding sequence, ex*ex, then fma(ey, ey, .), then fma(ez, ez, .):
// every search path (brick-grid walk, candidate lists, shell lists with packed f32x2 arithmetic)
// computes bit-identical distances, so they all pick the same neighbour.
PCR_HD float dist2_rn(float ex, float ey, float ez) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));
#else
    volatile float xx = ex * ex;
    return fmaf(ez, ez, fmaf(ey, ey, xx));
#endif
}

// --------------------------------------------------------------------------------------
// Two-level sparse uniform grid ("brick grid") used for every exact nearest-neighbour
// search (target points for ICP/PlaneICP/kNN normals, kept voxel means for VPlaneICP/NDT).
//
//   * fine cells of edge h; 4x4x4 cells form a brick
//   * dense brick table: one 16-byte record per brick = 64-bit occupancy mask of its cells
//     + ordinal of its first occupied cell in `cell_start`
//   * cell_start[ordinal] .. cell_start[ordinal+1] = range of the cell's points in `pts`
//   * pts: float4 (x, y, z, payload index bits), sorted by (brick, cell-in-brick)
//
// An empty cell costs no memory traffic beyond its brick record; a query touches
// 1 brick record + 1-2 cell_start words + the candidate points of the few occupied cells
// that intersect its current search ball.
// --------------------------------------------------------------------------------------
struct GridView {
    float ox, oy, oz;        // world position of cell (0,0,0)'s low corner
    float h, inv_h;          // cell edge and its reciprocal
    float slack;             // conservative inflation (grid units) covering f32 binning error
    int cnx, cny, cnz;       // cells per axis (multiples of 4)
    int bnx, bny, bnz;       // bricks per axis
    const uint4* bricks;     // (mask_lo, mask_hi, first_cell_ordinal, unused)
    const uint32_t* cell_start;
    const float4* pts;
    uint32_t n_pts;
};

// Per-cell exact candidate lists over a brick grid whose cells hold at most one point (the kept
// voxel means): for every "band" cell C near a point, the list holds every indexed point that can
// be the nearest neighbour of SOME location inside C:
//     K_C = { m : mindist(C, m) <= D_C },   D_C = min_m maxdist(C, m)
// (any query in C has its nearest neighbour within D_C, and that neighbour is at least
// mindist(C, .) away).  A query only evaluates the list of its own cell -- no search -- and all
// queries of one cell read the same addresses.  Cells without a list (far from every point, or
// D_C >= 2 cells so that the 5x5x5 build neighbourhood would not suffice) fall back to the
// general search.
struct CandLists {
    const uint4* bricks;         // (band mask lo, hi, ordinal of first band cell, unused), same brick layout as the grid
    const uint32_t* list_start;  // [n_band + 1]
    const uint32_t* list_idx;    // positions in GridView::pts
    const float4* list_pts;      // optional: the same entries with the point inline (x, y, z, position bits) -- one
                                 // load per candidate instead of an index load and a dependent gather
};

// Per-cell "shell lists" over the target-point grid.  For every band cell C (within Chebyshev
// distance 1 of an occupied cell) ONE contiguous list holds every indexed point whose distance to
// the box of C (its "margin") is <= dmax, ordered by margin level: level 0 = the points binned in
// C itself, levels 1.. = shells of growing margin.  Lists are padded to a multiple of four with
// sentinels and stored in groups of four entries, structure-of-arrays inside the group:
//     pts[4g] = (x0 x1 x2 x3), pts[4g+1] = (y0 ..), pts[4g+2] = (z0 ..), pts[4g+3] = positions of the
//     four points in GridView::pts (uint32 bits)
// so that three 16-byte loads feed four distance evaluations in packed f32x2 arithmetic and the
// positions are only read for the winner.  For every group `margin2[group]` is a lower bound of the
// squared margin of the group's and of all later entries.  A query in C needs no geometry at all:
//     for each group: if (margin2[group] >= best) stop;  evaluate the four entries
// -- a point not yet looked at is at least its margin away from any location in C -- and every
// query of one cell streams the SAME addresses for (nearly) the SAME number of steps, which is
// what the per-lane searches over cell walks could not offer (10-13 active lanes per warp
// instruction, the rest waiting for stragglers).  If the list ends while the best is still
// farther than dmax, the general search takes over from that bound.
#define PCR_SHELL_LEVELS 24
// level j >= 1 holds margins in (frac[j-1], frac[j]] cell edges (cut at dmax), frac growing by a
// constant factor (1.2113) per level up to 3 cell edges; level 0 = the cell's own points
#define PCR_SHELL_FRACS                                                                                                     \
    0.0f, 0.0442f, 0.0535f, 0.0649f, 0.0786f, 0.0952f, 0.1153f, 0.1396f, 0.1691f, 0.2049f, 0.2482f, 0.3006f, 0.3641f, 0.4411f, \
        0.5343f, 0.6472f, 0.7839f, 0.9496f, 1.1502f, 1.3933f, 1.6877f, 2.0443f, 2.4763f, 3.0f
#define PCR_SHELL_MAX_MARGIN 3.0
struct ShellLists {
    const uint4* bricks;         // (band mask lo, hi, ordinal of first band cell, unused), brick layout of the grid
    const uint32_t* start;       // [n_band + 1], in GROUPS of four entries (2^32 groups = 2^34 entries: a 100M-point target fits)
    const float4* pts;           // entries (+ 4 sentinels): group g = pts[4g .. 4g+3]
    const float* margin2;        // [groups]
    float covered2;              // (dmax - slack)^2: a best within it after the whole list is final
    int block_r;                 // the lists hold EVERY point of the (2 block_r + 1)^3 cell block around their cell
                                 // (1 when dmax >= sqrt(3) cell edges, else 0: only the cell itself)
};

PCR_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

PCR_HD int cell_of(float g, int n) {          // grid coordinate -> clamped cell index
    int c = (int)floorf(g);
    return clampi(c, 0, n - 1);
}

// bit index of cell (x,y,z) inside its brick
PCR_HD int brick_bit(int cx, int cy, int cz) { return ((cz & 3) << 4) | ((cy & 3) << 2) | (cx & 3); }

// 64-bit mask of the cells of one brick whose local coordinates lie in [x0,x1]x[y0,y1]x[z0,z1]
// (all in 0..3, inclusive, lo <= hi).
PCR_HD unsigned long long brick_box_mask(int x0, int x1, int y0, int y1, int z0, int z1) {
    // 32-bit arithmetic only (64-bit multiplies are slow on the GPU): one z slab is 16 bits
    const uint32_t ax = (2u << x1) - (1u << x0);                      // x pattern of one row (4 bits)
    const uint32_t ay = (2u << (4 * y1 + 3)) - (1u << (4 * y0));      // rows y0..y1 of one slab (16 bits)
    const uint32_t slab = (ax * 0x1111u) & ay;
    const uint32_t two = slab | (slab << 16);                         // the same pattern in two slabs
    const uint32_t lo = (z0 <= 0 ? 0x0000ffffu : 0u) | ((z0 <= 1 && z1 >= 1) ? 0xffff0000u : 0u);   // slabs 0, 1
    const uint32_t hi = ((z0 <= 2 && z1 >= 2) ? 0x0000ffffu : 0u) | (z1 >= 3 ? 0xffff0000u : 0u);   // slabs 2, 3
    return ((unsigned long long)(two & hi) << 32) | (unsigned long long)(two & lo);
}

}  // namespace pcr
