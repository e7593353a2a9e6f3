// Tile-cooperative exact nearest-neighbour search (device only).
//
// The per-lane search of pcr_grid.cuh is exact but every lane walks its own bricks / cells /
// points, so a warp executes the union of 32 different control flows (ncu, round 1: 6.4 of 32
// lanes active per instruction).  Here G consecutive, spatially sorted queries (a "tile" of G
// lanes, G = 8/16/32) search TOGETHER:
//
//   1. every lane proposes a cell box that must be searched for it (first its cell +- r0, later
//      the box enclosing the ball (query, best distance so far));
//   2. the tile takes the union box of a cluster of nearby proposals (a Morton curve jumps now
//      and then: far-away lanes simply wait for the next cluster);
//   3. the lanes enumerate the OCCUPIED cells of that box in parallel (one brick record per
//      lane, 64-bit masks), write the (start, length) ranges to shared memory, then copy the
//      candidate points -- which are contiguous per cell -- to a shared candidate list;
//   4. every lane evaluates the SAME candidate list (shared-memory broadcast reads, no
//      divergence) and keeps its own best;
//   5. a lane is finished when its best distance is not larger than its distance to the
//      boundary of the box searched so far (or that distance exceeds max_dist, or the box
//      covers the grid); otherwise it proposes a larger box and the tile iterates.
//
// Exactness argument: identical to pcr_grid.cuh -- a lane stops only when every unvisited
// cell is provably (slack-inflated) farther than its current best.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cooperative_groups/scan.h>

#include "pcr_common.cuh"

namespace pcr {
namespace cg = cooperative_groups;

constexpr int kTileQuota = 8;      // candidate points a lane may stage per pass

template <int G>
struct TileScratch {
    float4 cand[G * kTileQuota];   // staged candidates, .w = position in GridView::pts (int bits)
    uint2 cells[4 * G];            // (start, length) of occupied cells still to be staged
    float4 pad[2];                 // sizeof % 128 == 32: the tiles of one warp broadcast-read
                                   // cand[k] from different banks (no 4-way conflict for G = 8)
};

__device__ __forceinline__ int cell_clamped(float g, int n) {
    g = fminf(fmaxf(g, -1.0e9f), 1.0e9f);
    int c = (int)floorf(g);
    return c < 0 ? 0 : (c > n - 1 ? n - 1 : c);
}

// Fast path for a query that already has a tight upper bound (warm start from the previous
// Gauss-Newton iteration): visit the few cells that meet the ball (query, sqrt(best_d2)) --
// at most 3 per axis when the radius is <= one cell -- pruned by their box distance.  Purely
// per lane; all lanes run the same flattened loop over their (<= 27) cells.
__device__ __forceinline__ void local_nn_search(const GridView& Gv, float qx, float qy, float qz, float gx, float gy, float gz,
                                                float r, float& best_d2, int& best_pos) {
    const int lx = cell_clamped(gx - r, Gv.cnx), hx = cell_clamped(gx + r, Gv.cnx);
    const int ly = cell_clamped(gy - r, Gv.cny), hy = cell_clamped(gy + r, Gv.cny);
    const int lz = cell_clamped(gz - r, Gv.cnz), hz = cell_clamped(gz + r, Gv.cnz);
    const int nx = hx - lx + 1, ny = hy - ly + 1;
    const int n = nx * ny * (hz - lz + 1);
    const float h2 = Gv.h * Gv.h;
    int cur_brick = -1;
    unsigned long long occ = 0ull;
    uint32_t base = 0u;
    int cx = lx, cy = ly, cz = lz;
    for (int k = 0; k < n; ++k) {
        const int b = ((cz >> 2) * Gv.bny + (cy >> 2)) * Gv.bnx + (cx >> 2);
        if (b != cur_brick) {
            const uint4 rec = __ldg(Gv.bricks + b);
            occ = ((unsigned long long)rec.y << 32) | rec.x;
            base = rec.z;
            cur_brick = b;
        }
        const int bit = brick_bit(cx, cy, cz);
        if ((occ >> bit) & 1ull) {
            const float dx = fmaxf(fmaxf((float)cx - gx, gx - (float)(cx + 1)) - Gv.slack, 0.0f);
            const float dy = fmaxf(fmaxf((float)cy - gy, gy - (float)(cy + 1)) - Gv.slack, 0.0f);
            const float dz = fmaxf(fmaxf((float)cz - gz, gz - (float)(cz + 1)) - Gv.slack, 0.0f);
            if ((dx * dx + dy * dy + dz * dz) * h2 < best_d2) {
                const uint32_t ord = base + (uint32_t)__popcll(occ & ((1ull << bit) - 1ull));
                const uint32_t s = __ldg(Gv.cell_start + ord), e = __ldg(Gv.cell_start + ord + 1);
                for (uint32_t p = s; p < e; ++p) {
                    const float4 t = __ldg(Gv.pts + p);
                    const float ex = t.x - qx, ey = t.y - qy, ez = t.z - qz;
                    const float d2 = ex * ex + ey * ey + ez * ez;
                    if (d2 < best_d2) { best_d2 = d2; best_pos = (int)p; }
                }
            }
        }
        if (++cx > hx) { cx = lx; if (++cy > hy) { cy = ly; ++cz; } }
    }
}

// Searches for the nearest point of every lane's query.  All lanes of the tile must call it
// (lanes without a query pass valid = false).  On entry (best_d2, best_pos) is either
// (max_dist^2, -1) or a WARM START: the squared distance to / position of any indexed point
// (e.g. the previous iteration's match) -- an upper bound that lets the lane search the box
// of that ball at once.  r0 = first search radius in grid units for lanes without a warm start.
template <int G>
__device__ __forceinline__ void tile_nn_search(const cg::thread_block_tile<G>& tile, const GridView& Gv, TileScratch<G>& S,
                                               bool valid, float qx, float qy, float qz, float r0, float max_d2,
                                               float& best_d2, int& best_pos) {
    constexpr int LISTCAP = 4 * G;
    const float gx = (qx - Gv.ox) * Gv.inv_h, gy = (qy - Gv.oy) * Gv.inv_h, gz = (qz - Gv.oz) * Gv.inv_h;
    bool pending = valid && (gx == gx) && (gy == gy) && (gz == gz) && Gv.n_pts != 0;
    if (pending) {   // farther from the whole grid than max_dist: no match possible
        const float ex = fmaxf(fmaxf(-gx, gx - (float)Gv.cnx), 0.0f);
        const float ey = fmaxf(fmaxf(-gy, gy - (float)Gv.cny), 0.0f);
        const float ez = fmaxf(fmaxf(-gz, gz - (float)Gv.cnz), 0.0f);
        const float e = fmaxf(sqrtf(ex * ex + ey * ey + ez * ez) - Gv.slack, 0.0f) * Gv.h;
        if (e * e >= max_d2) pending = false;
    }
    if (best_pos >= 0) r0 = sqrtf(best_d2) * Gv.inv_h * 1.000001f + Gv.slack;     // warm start: box of the ball
    int lo0 = cell_clamped(gx - r0, Gv.cnx), hi0 = cell_clamped(gx + r0, Gv.cnx);
    int lo1 = cell_clamped(gy - r0, Gv.cny), hi1 = cell_clamped(gy + r0, Gv.cny);
    int lo2 = cell_clamped(gz - r0, Gv.cnz), hi2 = cell_clamped(gz + r0, Gv.cnz);
    const int rank = tile.thread_rank();
    // box visited by the previous round (every lane of the tile evaluated all of its points)
    int p0 = 0, p1 = 0, p2 = 0, w0 = -1, w1 = -1, w2 = -1;

    for (;;) {
        const unsigned act = tile.ballot(pending);
        if (act == 0u) break;
        const int leader = __ffs(act) - 1;
        // ---- cluster = pending lanes whose box lies within 2 cells of the leader's box ----
        const int L0 = tile.shfl(lo0, leader), H0 = tile.shfl(hi0, leader);
        const int L1 = tile.shfl(lo1, leader), H1 = tile.shfl(hi1, leader);
        const int L2 = tile.shfl(lo2, leader), H2 = tile.shfl(hi2, leader);
        const int gap = max(max(max(L0 - hi0, lo0 - H0), max(L1 - hi1, lo1 - H1)), max(L2 - hi2, lo2 - H2));
        const bool member = pending && gap <= 2;
        const int u0 = cg::reduce(tile, member ? lo0 : INT_MAX, cg::less<int>());
        const int u1 = cg::reduce(tile, member ? lo1 : INT_MAX, cg::less<int>());
        const int u2 = cg::reduce(tile, member ? lo2 : INT_MAX, cg::less<int>());
        const int v0 = cg::reduce(tile, member ? hi0 : INT_MIN, cg::greater<int>());
        const int v1 = cg::reduce(tile, member ? hi1 : INT_MIN, cg::greater<int>());
        const int v2 = cg::reduce(tile, member ? hi2 : INT_MIN, cg::greater<int>());
        // cells of the previous round's box need no second visit (if it lies inside this one)
        const bool skip_prev = w0 >= p0 && p0 >= u0 && w0 <= v0 && p1 >= u1 && w1 <= v1 && p2 >= u2 && w2 <= v2;

        // ---- visit every occupied cell of the union box [u, v] ----
        const int bx0 = u0 >> 2, by0 = u1 >> 2, bz0 = u2 >> 2;
        const int nbx = (v0 >> 2) - bx0 + 1, nby = (v1 >> 2) - by0 + 1, nbz = (v2 >> 2) - bz0 + 1;
        const int nb = nbx * nby * nbz;
        for (int b0 = 0; b0 < nb; b0 += G) {
            const int b = b0 + rank;
            unsigned long long m = 0ull, occ = 0ull;
            uint32_t base = 0u;
            if (b < nb) {
                const int ix = b % nbx, t = b / nbx;
                const int bx = bx0 + ix, by = by0 + t % nby, bz = bz0 + t / nby;
                const int x0 = max(u0 - bx * 4, 0), x1 = min(v0 - bx * 4, 3);
                const int y0 = max(u1 - by * 4, 0), y1 = min(v1 - by * 4, 3);
                const int z0 = max(u2 - bz * 4, 0), z1 = min(v2 - bz * 4, 3);
                unsigned long long keep = brick_box_mask(x0, x1, y0, y1, z0, z1);
                if (skip_prev) {
                    const int a0 = max(p0 - bx * 4, 0), a1 = min(w0 - bx * 4, 3);
                    const int b0_ = max(p1 - by * 4, 0), b1 = min(w1 - by * 4, 3);
                    const int c0_ = max(p2 - bz * 4, 0), c1 = min(w2 - bz * 4, 3);
                    if (a0 <= a1 && b0_ <= b1 && c0_ <= c1) keep &= ~brick_box_mask(a0, a1, b0_, b1, c0_, c1);
                }
                if (keep) {
                    const uint4 rec = __ldg(Gv.bricks + ((size_t)bz * Gv.bny + by) * Gv.bnx + bx);
                    occ = ((unsigned long long)rec.y << 32) | rec.x;
                    m = occ & keep;
                    base = rec.z;
                }
            }
            while (tile.any(m != 0ull)) {
                // phase 1: occupied cells -> (start, length) list in shared memory
                const int c = __popcll(m);
                int off = cg::exclusive_scan(tile, c);
                const int total = tile.shfl(off + c, G - 1);
                while (m != 0ull && off < LISTCAP) {
                    const int bit = __ffsll((long long)m) - 1;
                    m &= m - 1ull;
                    const uint32_t ord = base + (uint32_t)__popcll(occ & ((1ull << bit) - 1ull));
                    const uint32_t s = __ldg(Gv.cell_start + ord), e = __ldg(Gv.cell_start + ord + 1);
                    S.cells[off++] = make_uint2(s, e - s);
                }
                const int n_list = total < LISTCAP ? total : LISTCAP;
                tile.sync();
                // phase 2: stage candidate points; phase 3: everybody evaluates them
                for (int c0 = 0; c0 < n_list; c0 += G) {
                    uint32_t s = 0u, len = 0u;
                    if (c0 + rank < n_list) { const uint2 cl = S.cells[c0 + rank]; s = cl.x; len = cl.y; }
                    while (tile.any(len != 0u)) {
                        const int t = (int)(len < (uint32_t)kTileQuota ? len : (uint32_t)kTileQuota);
                        const int o = cg::exclusive_scan(tile, t);
                        const int n = tile.shfl(o + t, G - 1);
                        for (int k = 0; k < t; ++k) {
                            float4 p = __ldg(Gv.pts + s + k);
                            p.w = __int_as_float((int)(s + k));
                            S.cand[o + k] = p;
                        }
                        s += t; len -= t;
                        tile.sync();
#pragma unroll 4
                        for (int k = 0; k < n; ++k) {
                            const float4 cpt = S.cand[k];
                            const float ex = cpt.x - qx, ey = cpt.y - qy, ez = cpt.z - qz;
                            const float d2 = ex * ex + ey * ey + ez * ez;
                            if (d2 < best_d2) { best_d2 = d2; best_pos = __float_as_int(cpt.w); }
                        }
                        tile.sync();
                    }
                }
            }
        }

        p0 = u0; p1 = u1; p2 = u2; w0 = v0; w1 = v1; w2 = v2;
        // ---- members decide whether they are finished ----
        if (member) {
            float bound = 3.0e38f;
            bool open = false;
            if (u0 > 0) { bound = fminf(bound, gx - (float)u0); open = true; }
            if (v0 < Gv.cnx - 1) { bound = fminf(bound, (float)(v0 + 1) - gx); open = true; }
            if (u1 > 0) { bound = fminf(bound, gy - (float)u1); open = true; }
            if (v1 < Gv.cny - 1) { bound = fminf(bound, (float)(v1 + 1) - gy); open = true; }
            if (u2 > 0) { bound = fminf(bound, gz - (float)u2); open = true; }
            if (v2 < Gv.cnz - 1) { bound = fminf(bound, (float)(v2 + 1) - gz); open = true; }
            bound -= Gv.slack;
            const float rad = sqrtf(best_d2) * Gv.inv_h;
            if (!open || rad <= bound) {
                pending = false;
            } else if (best_pos >= 0) {
                const float r = rad * 1.000001f + Gv.slack;          // box enclosing the ball, merged with the visited box
                lo0 = min(u0, cell_clamped(gx - r, Gv.cnx)); hi0 = max(v0, cell_clamped(gx + r, Gv.cnx));
                lo1 = min(u1, cell_clamped(gy - r, Gv.cny)); hi1 = max(v1, cell_clamped(gy + r, Gv.cny));
                lo2 = min(u2, cell_clamped(gz - r, Gv.cnz)); hi2 = max(v2, cell_clamped(gz + r, Gv.cnz));
                // every cell meeting the (inflated) ball has been visited: finished.  Also what
                // guarantees progress when the two float tests above disagree by an ulp.
                if (lo0 == u0 && hi0 == v0 && lo1 == u1 && hi1 == v1 && lo2 == u2 && hi2 == v2) pending = false;
            } else {
                // nothing found yet: double the margin around the visited box
                const int ex = max(1, (v0 - u0 + 1) >> 1), ey = max(1, (v1 - u1 + 1) >> 1), ez = max(1, (v2 - u2 + 1) >> 1);
                lo0 = max(u0 - ex, 0); hi0 = min(v0 + ex, Gv.cnx - 1);
                lo1 = max(u1 - ey, 0); hi1 = min(v1 + ey, Gv.cny - 1);
                lo2 = max(u2 - ez, 0); hi2 = min(v2 + ez, Gv.cnz - 1);
            }
        }
    }
}

}  // namespace pcr
