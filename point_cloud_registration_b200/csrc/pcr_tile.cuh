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
//   * the ball of the current best distance meets a row in one x-interval, so the per-lane search
//     is "for each row near the query: one range loop" -- no per-cell walk.
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
    const float4* pts;       // [n] (x, y, z, own position as uint32 bits), sorted by cell number
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

// candidates [lo, hi) of one range
PCR_HD void tile_eval_range(const float4* pts, uint32_t lo, uint32_t hi, const TileQuery& q, TileBest& b) {
    for (uint32_t s = lo; s < hi; ++s) {
        const float4 t = pts[s];
        const float d = dist2_rn(t.x - q.qx, t.y - q.qy, t.z - q.qz);
        if (d < b.d2) {
            b.d2 = d; b.x = t.x; b.y = t.y; b.z = t.z;
#if defined(__CUDA_ARCH__)
            b.pos = __float_as_uint(t.w);
#else
            memcpy(&b.pos, &t.w, 4);
#endif
        }
    }
}

// One (jy, jz) row whose cells X0..X1 are addressable: csrow[j] = index in `pts` of the first point
// of cell X0 + j (j = 0 .. X1 - X0 + 1).  Evaluates the cells the ball of the current best meets;
// skip_ix = a cell of this row that has been evaluated already (INT_MIN: none).
PCR_HD void tile_visit_row(const TileGrid& G, const TileQuery& q, TileBest& b, int jy, int jz, int X0, int X1,
                           const uint32_t* csrow, const float4* pts, int skip_ix) {
    const float dy = fmaxf(fmaxf((float)jy - q.gy, q.gy - (float)(jy + 1)) - G.slack, 0.0f);
    const float dz = fmaxf(fmaxf((float)jz - q.gz, q.gz - (float)(jz + 1)) - G.slack, 0.0f);
    const float r2 = b.d2 * G.inv_c2 * 1.00001f;                 // pruning radius^2 in grid units, rounded up
    const float rem = r2 - (dy * dy + dz * dz);
    if (!(rem > 0.0f)) return;                                   // the row's (y,z) rectangle is beyond the best
    const float rx = sqrtf(rem) * 1.000001f + G.slack;
    int xa = tile_cell_floor(q.gx - rx), xb = tile_cell_floor(q.gx + rx);
    xa = xa > X0 ? xa : X0;
    xb = xb < X1 ? xb : X1;
    if (xa > xb) return;
    if (skip_ix >= xa && skip_ix <= xb) {
        tile_eval_range(pts, csrow[xa - X0], csrow[skip_ix - X0], q, b);
        tile_eval_range(pts, csrow[skip_ix + 1 - X0], csrow[xb + 1 - X0], q, b);
    } else {
        tile_eval_range(pts, csrow[xa - X0], csrow[xb + 1 - X0], q, b);
    }
}

// e-th row offset (dy, dz) of the Chebyshev ring k >= 1 around the own row (8k rows)
PCR_HD void tile_ring_offset(int k, int e, int& dy, int& dz) {
    const int s = 2 * k + 1;
    if (e < s) { dz = -k; dy = e - k; }
    else if (e < 2 * s) { dz = k; dy = e - s - k; }
    else if (e < 3 * s - 2) { dy = -k; dz = e - 2 * s - k + 1; }
    else { dy = k; dz = e - (3 * s - 2) - k + 1; }
}

// The rows of box R with row numbers [ra, ra + nfit) are resident (row number r = (jy - R.y0) +
// rny * (jz - R.z0); scs[(r - ra) * W + j] = index in spts of the first point of cell R.x0 + j).
// RING ORDER: own cell, own row, then the rings of rows around it, nearest first, stopping as soon
// as a whole ring is beyond the best -- for a box that is resident in one piece.
PCR_HD void tile_search_rings(const TileGrid& G, const TileQuery& q, TileBest& b, const TileBox& R, int ra, int nfit, int W,
                              const uint32_t* scs, const float4* spts) {
    const int rny = R.y1 - R.y0 + 1;
    const bool own_row_in = q.iy >= R.y0 && q.iy <= R.y1 && q.iz >= R.z0 && q.iz <= R.z1;
    if (own_row_in) {
        const int lr = (q.iy - R.y0) + rny * (q.iz - R.z0) - ra;
        if (lr >= 0 && lr < nfit) {
            const uint32_t* csrow = scs + lr * W;
            int skip = INT_MIN;
            if (q.ix >= R.x0 && q.ix <= R.x1) {
                tile_eval_range(spts, csrow[q.ix - R.x0], csrow[q.ix + 1 - R.x0], q, b);
                skip = q.ix;
            }
            tile_visit_row(G, q, b, q.iy, q.iz, R.x0, R.x1, csrow, spts, skip);
        }
    }
    // distance (grid units) from the query to the nearest (y,z) face of its own row: a lower
    // bound of ring k's distance is (k - 1) + that
    const float fy = q.gy - (float)q.iy, fz = q.gz - (float)q.iz;
    const float myz = fminf(fminf(fy, 1.0f - fy), fminf(fz, 1.0f - fz));
    int kmax = q.iy - R.y0;
    kmax = (R.y1 - q.iy) > kmax ? (R.y1 - q.iy) : kmax;
    kmax = (q.iz - R.z0) > kmax ? (q.iz - R.z0) : kmax;
    kmax = (R.z1 - q.iz) > kmax ? (R.z1 - q.iz) : kmax;
    for (int k = 1; k <= kmax; ++k) {
        const float lb = (float)(k - 1) + myz - G.slack;
        if (lb > 0.0f && lb * lb >= b.d2 * G.inv_c2 * 1.00001f) break;
        for (int e = 0; e < 8 * k; ++e) {
            int dy, dz;
            tile_ring_offset(k, e, dy, dz);
            const int jy = q.iy + dy, jz = q.iz + dz;
            if (jy < R.y0 || jy > R.y1 || jz < R.z0 || jz > R.z1) continue;
            const int lr = (jy - R.y0) + rny * (jz - R.z0) - ra;
            if (lr < 0 || lr >= nfit) continue;
            tile_visit_row(G, q, b, jy, jz, R.x0, R.x1, scs + lr * W, spts, INT_MIN);
        }
    }
}

// MEMORY ORDER: every resident row once, pruned by the best -- for one batch of a box that is
// staged in several pieces (large halos).
PCR_HD void tile_search_linear(const TileGrid& G, const TileQuery& q, TileBest& b, const TileBox& R, int ra, int nfit, int W,
                               const uint32_t* scs, const float4* spts) {
    const int rny = R.y1 - R.y0 + 1;
    for (int lr = 0; lr < nfit; ++lr) {
        const int r = ra + lr;
        const int jz = R.z0 + r / rny, jy = R.y0 + r % rny;
        tile_visit_row(G, q, b, jy, jz, R.x0, R.x1, scs + lr * W, spts, INT_MIN);
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

// the best is final: nothing outside the staged box can be closer (also true for "no match": the
// pruning radius is then max_dist itself), or the box held the whole grid
PCR_HD bool tile_settled(const TileGrid& G, const TileQuery& q, const TileBest& b, const TileBox& U) {
    if (U.x0 <= 0 && U.y0 <= 0 && U.z0 <= 0 && U.x1 >= G.nx - 1 && U.y1 >= G.ny - 1 && U.z1 >= G.nz - 1) return true;
    const float g = tile_guarantee(G, q, U);
    return g > 0.0f && b.d2 * G.inv_c2 * 1.00001f <= g * g;
}

// Halo radius (grid units) at which every joined lane is settled whatever it found: a lane is at
// least radius - slack away from every face of the box staged around the joined lanes.
PCR_HD float tile_rmax(const TileGrid& G, float max_d2) {
    return sqrtf(max_d2) * G.inv_c * 1.00001f + 2.0f * G.slack + 1.0e-3f;
}

// Halo radius that settles a lane for certain in the NEXT pass: the ball of its current candidate.
PCR_HD float tile_radius_for(const TileGrid& G, const TileBest& b) {
    return sqrtf(b.d2) * G.inv_c * 1.00001f + 2.0f * G.slack;
}

}  // namespace pcr
