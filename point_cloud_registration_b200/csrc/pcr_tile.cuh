// Tile-stream correspondence search: the per-lane half (plain PCR_HD code, replayed on the host by
// tests/hostsim) of the fused linearise kernel in pcr_tile_kernel.cuh.
//
// Structure ("row grid", TileGrid): a DENSE uniform grid over the indexed points (target points
// for ICP / PlaneICP, kept voxel means for VPlaneICP / NDT) whose cells are numbered x-fastest,
// and the points sorted by that cell number.  Two properties carry the whole design:
//
//   * any run of cells [xa..xb] of one (y,z) row is ONE contiguous range of the point array,
//     cs[row*nx + xa] .. cs[row*nx + xb + 1]: a box of cells is rny*rnz contiguous byte ranges,
//     which is exactly what a 1-D bulk copy (cp.async.bulk, TMA engine) moves into shared memory;
//   * once a box is staged, EVERY lane of the warp evaluates EVERY staged candidate: all lanes read
//     the same shared-memory addresses (broadcast), nobody diverges, no per-lane cell walk exists
//     (measured: a pruned per-lane walk of the same box ran at 9 of 32 active lanes and needed
//     three times the instructions, profiles/r2_notes.md).  Candidates are stored as pair records
//     so that two distances cost six packed f32x2 instructions.
//
// A warp takes 32 consecutive scan points (the scan is uploaded in cell order, so they sit in a
// handful of neighbouring cells), stages the cells that meet the box [min - rho, max + rho] around
// them (rho = halo radius, a fraction of a cell once the scan is nearly aligned) in shared memory
// and every lane searches that box; a lane is SETTLED when its best distance does not exceed its
// distance to the faces of the staged box -- nothing outside can be closer.  A lane that is not
// settled asks for the ball of its candidate (which settles it for certain) or, without a
// candidate, twice the radius: exact for any displacement, no separate fallback structure.  The
// radius a row needed is remembered as the next iteration's first guess.
//
// Replaces KDTree.query at icp.py:33, plane_icp.py:40, voxel.py:176 (exact 1-NN, strict
// `dist < max_dist`, quirk Q4).
#pragma once
#include <climits>
#include <cstring>
#include "pcr_common.cuh"

namespace pcr {

struct TileGrid {
    float ox, oy, oz;        // world position of the low corner of cell (0,0,0)
    float c, inv_c;          // cell edge and reciprocal
    float inv_c2;            // inv_c^2
    float slack;             // conservative inflation (grid units) covering float32 binning error
    int nx, ny, nz;          // cells per axis
    const uint32_t* cs;      // [nx*ny*nz + 1] first point of every cell, x fastest, then y, then z
    const float4* pairs;     // [2 * ceil(n / 2)] points sorted by cell number, stored as PAIR records of 32 bytes:
                             //   pairs[2p] = (x0, x1, y0, y1), pairs[2p+1] = (z0, z1, w0, w1), w = own position (uint32 bits);
                             //   an odd tail is filled with a sentinel that is infinitely far from every query
    uint32_t n;
};

struct TileBox { int x0, x1, y0, y1, z0, z1; };   // inclusive cell box

struct TileQuery {
    float qx, qy, qz;        // posed scan point (world, float32: quirk Q7)
    float gx, gy, gz;        // grid coordinates (cell units), clamped to +-2^20
    int ix, iy, iz;          // cell (NOT clamped to the grid: a query may lie outside it)
};

struct TileBest {
    float d2;                // best squared distance so far (starts at max_dist^2, strict <)
    float x, y, z;           // the matched point
    uint32_t pos;            // its position in TileGrid::pts, 0xffffffff = none
};

constexpr uint32_t kTileNone = 0xffffffffu;
// host replay only (tests/hostsim): work counters of the per-lane search (never touched by device code)
inline long long g_tile_evals = 0;
constexpr float kTileCoordLimit = 1048576.0f;   // 2^20 cells

// floor to a cell number, safe for any float (huge halo radii of an unbounded max_dist included)
PCR_HD int tile_cell_floor(float v) {
    return (int)floorf(fminf(fmaxf(v, -2.0f * kTileCoordLimit), 2.0f * kTileCoordLimit));
}

// false: NaN query (padding or a NaN scan point) -- no correspondence
PCR_HD bool tile_make_query(const TileGrid& G, float qx, float qy, float qz, TileQuery& q) {
    q.qx = qx; q.qy = qy; q.qz = qz;
    float gx = (qx - G.ox) * G.inv_c, gy = (qy - G.oy) * G.inv_c, gz = (qz - G.oz) * G.inv_c;
    if (!(gx == gx) || !(gy == gy) || !(gz == gz)) return false;
    gx = fminf(fmaxf(gx, -kTileCoordLimit), kTileCoordLimit);
    gy = fminf(fmaxf(gy, -kTileCoordLimit), kTileCoordLimit);
    gz = fminf(fmaxf(gz, -kTileCoordLimit), kTileCoordLimit);
    q.gx = gx; q.gy = gy; q.gz = gz;
    q.ix = (int)floorf(gx); q.iy = (int)floorf(gy); q.iz = (int)floorf(gz);
    return true;
}

// true: the whole grid is at least sqrt(max_d2) away from the query -- nothing can match
PCR_HD bool tile_query_far_outside(const TileGrid& G, const TileQuery& q, float max_d2) {
    const float ex = fmaxf(fmaxf(-q.gx, q.gx - (float)G.nx), 0.0f);
    const float ey = fmaxf(fmaxf(-q.gy, q.gy - (float)G.ny), 0.0f);
    const float ez = fmaxf(fmaxf(-q.gz, q.gz - (float)G.nz), 0.0f);
    const float e = fmaxf(sqrtf(ex * ex + ey * ey + ez * ez) - G.slack, 0.0f);
    return e * e >= max_d2 * G.inv_c2 * 1.00001f;
}

// Squared distances of the two candidates of one pair record (see TileGrid::pairs) to one query,
// with the rounding sequence of dist2_rn (packed f32x2 arithmetic on sm_100: six instructions).
PCR_HD void tile_pair_d2(const float4& A, float bz0, float bz1, const TileQuery& q, float& d0, float& d1) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    const float2 ex = __fadd2_rn(make_float2(A.x, A.y), make_float2(-q.qx, -q.qx));
    const float2 ey = __fadd2_rn(make_float2(A.z, A.w), make_float2(-q.qy, -q.qy));
    const float2 ez = __fadd2_rn(make_float2(bz0, bz1), make_float2(-q.qz, -q.qz));
    const float2 r = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
    d0 = r.x; d1 = r.y;
#else
    d0 = dist2_rn(A.x - q.qx, A.z - q.qy, bz0 - q.qz);
    d1 = dist2_rn(A.y - q.qx, A.w - q.qy, bz1 - q.qz);
#endif
}

// Every lane evaluates EVERY staged candidate: `np` pair records at `pairs` (shared or global memory,
// the same addresses for all lanes of the warp: broadcast loads, no divergence), four candidates per
// trip; only the minimum of a trip is compared with the best, and only the trip that improved it is
// remembered (three instructions) -- tile_take finds the winner inside that trip afterwards.
// Returns the first pair of the winning trip or kTileNone; `best` is updated.
PCR_HD uint32_t tile_scan_pairs(const float4* pairs, uint32_t np, const TileQuery& q, float& best) {
    uint32_t trip = kTileNone;
    uint32_t p = 0;
    for (; p + 2 <= np; p += 2) {
        const float4 A0 = pairs[2 * p], B0 = pairs[2 * p + 1], A1 = pairs[2 * p + 2], B1 = pairs[2 * p + 3];
        float d0, d1, d2, d3;
        tile_pair_d2(A0, B0.x, B0.y, q, d0, d1);
        tile_pair_d2(A1, B1.x, B1.y, q, d2, d3);
        const float dm = fminf(fminf(d0, d1), fminf(d2, d3));
        if (dm < best) { best = dm; trip = p; }
    }
    if (p < np) {
        const float4 A0 = pairs[2 * p], B0 = pairs[2 * p + 1];
        float d0, d1;
        tile_pair_d2(A0, B0.x, B0.y, q, d0, d1);
        const float dm = fminf(d0, d1);
        if (dm < best) { best = dm; trip = p; }
    }
#if !defined(__CUDA_ARCH__)
    g_tile_evals += 2 * (long long)np;
#endif
    return trip;
}

// the candidate of the trip starting at pair `trip` whose distance is `best` -> b (the distances are
// recomputed with the same roundings, so the comparison is exact; first of equals wins)
PCR_HD void tile_take(const float4* pairs, uint32_t trip, uint32_t np, const TileQuery& q, float best, TileBest& b) {
    const uint32_t pe = trip + 2 <= np ? trip + 2 : np;
    for (uint32_t p = trip; p < pe; ++p) {
        const float4 A = pairs[2 * p], B = pairs[2 * p + 1];
        float d0, d1;
        tile_pair_d2(A, B.x, B.y, q, d0, d1);
        if (d0 == best || d1 == best) {
            const bool hi = d0 != best;
            b.d2 = best; b.x = hi ? A.y : A.x; b.y = hi ? A.w : A.z; b.z = hi ? B.y : B.x;
            const float w = hi ? B.w : B.z;
#if defined(__CUDA_ARCH__)
            b.pos = __float_as_uint(w);
#else
            memcpy(&b.pos, &w, 4);
#endif
            return;
        }
    }
}

// Distance (grid units) from the query to the nearest face of the UNCLIPPED staged box U: every
// indexed point outside the box is at least this far away (outside the grid there are no points).
PCR_HD float tile_guarantee(const TileGrid& G, const TileQuery& q, const TileBox& U) {
    float g = q.gx - (float)U.x0;
    g = fminf(g, (float)(U.x1 + 1) - q.gx);
    g = fminf(g, q.gy - (float)U.y0);
    g = fminf(g, (float)(U.y1 + 1) - q.gy);
    g = fminf(g, q.gz - (float)U.z0);
    g = fminf(g, (float)(U.z1 + 1) - q.gz);
    return g - G.slack;
}

PCR_HD bool tile_box_holds_grid(const TileGrid& G, const TileBox& U) {
    return U.x0 <= 0 && U.y0 <= 0 && U.z0 <= 0 && U.x1 >= G.nx - 1 && U.y1 >= G.ny - 1 && U.z1 >= G.nz - 1;
}

// the best is final: nothing outside the staged box can be closer (also true for "no match": b.d2
// is then max_dist^2 itself), or the box held the whole grid
PCR_HD bool tile_settled(const TileGrid& G, const TileQuery& q, const TileBest& b, const TileBox& U) {
    if (tile_box_holds_grid(G, U)) return true;
    const float g = tile_guarantee(G, q, U);
    return g > 0.0f && b.d2 * G.inv_c2 * 1.00001f <= g * g;
}

// Halo radius that settles a lane for certain in the NEXT pass: the ball of its current candidate.
PCR_HD float tile_radius_for(const TileGrid& G, const TileBest& b) {
    return sqrtf(b.d2) * G.inv_c * 1.00001f + 2.0f * G.slack;
}

// Halo radius (grid units) at which every joined lane is settled whatever it found: a lane is at
// least radius - slack away from every face of the box staged around the joined lanes.
PCR_HD float tile_rmax(const TileGrid& G, float max_d2) {
    return sqrtf(max_d2) * G.inv_c * 1.00001f + 2.0f * G.slack + 1.0e-3f;
}

}  // namespace pcr
