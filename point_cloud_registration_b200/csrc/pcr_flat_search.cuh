// Resumable ("flat") form of the exact brick-grid nearest-neighbour search of pcr_grid.cuh.
//
// grid_search() runs one query to completion inside nested loops; on a GPU the 32 lanes of a
// warp then wait for the slowest query of every row of 32 (measured: 10-13 active lanes per
// warp instruction, profiles/r1_notes.md).  Here the same search is a small per-lane state
// machine with three operations
//
//     flat_begin      start a query (own cell first)
//     flat_next_cell  advance to the next occupied, unpruned cell of the current pass
//     flat_next_pass  plan the next pass (grow one ring / jump to the ball of the best) or finish
//     flat_eval       evaluate a bounded number of candidates of the current cell
//
// so that a warp can run ONE loop { find a cell | refill finished lanes with new queries |
// evaluate <= CH candidates } in which a lane that finishes early immediately starts its next
// query instead of idling (persistent-lane scheduling).  The sequence of passes, the pruning
// rules and therefore the exactness argument are those of grid_search():
//   pass 0 = the query's own cell; while nothing is found grow the visited block by one ring;
//   once a candidate exists visit the cell box of the ball (query, best) once and stop; a cell is
//   skipped only when its slack-inflated box is not closer than the current best.
// Replaces pykdtree's KDTree.query at icp.py:33 / plane_icp.py:40 / voxel.py:176.
#pragma once
#include "pcr_grid.cuh"

namespace pcr {

struct FlatLane {
    float qx, qy, qz;          // query (world)
    float gx, gy, gz;          // query (grid units)
    float best_d2;             // pruning radius^2 (starts at max_dist^2, strict <)
    int best_pos;              // position in G.pts of the best so far, -1 = none
    Block3 cur;                // block of cells already visited (x1 < x0: none)
    Block3 nb;                 // block being visited in this pass
    int last;                  // this pass covers the ball of the best: the query ends after it
    int bx, by, bz;            // brick iterator inside nb
    unsigned long long m;      // cells of the current brick still to be looked at
    unsigned long long occ;    // occupancy mask of the current brick
    uint32_t base;             // ordinal of the current brick's first occupied cell
    uint32_t p, e;             // candidates of the current cell still to be evaluated: [p, e)
};

PCR_HD void flat_reset_iter(FlatLane& L) {
    L.bx = (L.nb.x0 >> 2) - 1; L.by = L.nb.y0 >> 2; L.bz = L.nb.z0 >> 2;
    L.m = 0ull;
}

// Start a query.  Returns false when no point can match (empty index, NaN query, query farther
// than the search radius from the grid's bounding box); best_pos is -1 in that case.
PCR_HD bool flat_begin(const GridView& G, FlatLane& L, float qx, float qy, float qz, float max_d2) {
    L.qx = qx; L.qy = qy; L.qz = qz;
    L.best_d2 = max_d2; L.best_pos = -1;
    L.p = L.e = 0u;
    if (G.n_pts == 0) return false;
    const float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
    if (!(gx == gx) || !(gy == gy) || !(gz == gz)) return false;
    {
        const float ex = fmaxf(fmaxf(-gx, gx - (float)G.cnx), 0.0f);
        const float ey = fmaxf(fmaxf(-gy, gy - (float)G.cny), 0.0f);
        const float ez = fmaxf(fmaxf(-gz, gz - (float)G.cnz), 0.0f);
        const float e = fmaxf(sqrtf(ex * ex + ey * ey + ez * ez) - G.slack, 0.0f) * G.h;
        if (e * e >= max_d2) return false;
    }
    L.gx = gx; L.gy = gy; L.gz = gz;
    const float big = 1.0e9f;
    L.cur.x0 = L.cur.y0 = L.cur.z0 = 0; L.cur.x1 = L.cur.y1 = L.cur.z1 = -1;
    L.nb.x0 = L.nb.x1 = cell_of(fminf(fmaxf(gx, -big), big), G.cnx);
    L.nb.y0 = L.nb.y1 = cell_of(fminf(fmaxf(gy, -big), big), G.cny);
    L.nb.z0 = L.nb.z1 = cell_of(fminf(fmaxf(gz, -big), big), G.cnz);
    L.last = 0;
    flat_reset_iter(L);
    return true;
}

// Advance to the next occupied cell of nb \ cur whose box is closer than the best; sets [p, e).
// Returns false when the pass is exhausted.
PCR_HD bool flat_next_cell(const GridView& G, FlatLane& L) {
    const float h2 = G.h * G.h;
    for (;;) {
        while (L.m == 0ull) {
            const int bx0 = L.nb.x0 >> 2, bx1 = L.nb.x1 >> 2, by0 = L.nb.y0 >> 2, by1 = L.nb.y1 >> 2, bz1 = L.nb.z1 >> 2;
            if (++L.bx > bx1) {
                L.bx = bx0;
                if (++L.by > by1) {
                    L.by = by0;
                    if (++L.bz > bz1) return false;
                }
            }
            const int x4 = L.bx * 4, y4 = L.by * 4, z4 = L.bz * 4;
            const int lx0 = L.nb.x0 > x4 ? L.nb.x0 - x4 : 0, lx1 = L.nb.x1 < x4 + 3 ? L.nb.x1 - x4 : 3;
            const int ly0 = L.nb.y0 > y4 ? L.nb.y0 - y4 : 0, ly1 = L.nb.y1 < y4 + 3 ? L.nb.y1 - y4 : 3;
            const int lz0 = L.nb.z0 > z4 ? L.nb.z0 - z4 : 0, lz1 = L.nb.z1 < z4 + 3 ? L.nb.z1 - z4 : 3;
            unsigned long long keep = brick_box_mask(lx0, lx1, ly0, ly1, lz0, lz1);
            // cells visited by earlier passes (cur intersected with this brick; empty when cur is none)
            const int ox0 = L.cur.x0 > x4 ? L.cur.x0 - x4 : 0, ox1 = L.cur.x1 < x4 + 3 ? L.cur.x1 - x4 : 3;
            const int oy0 = L.cur.y0 > y4 ? L.cur.y0 - y4 : 0, oy1 = L.cur.y1 < y4 + 3 ? L.cur.y1 - y4 : 3;
            const int oz0 = L.cur.z0 > z4 ? L.cur.z0 - z4 : 0, oz1 = L.cur.z1 < z4 + 3 ? L.cur.z1 - z4 : 3;
            if (ox0 <= ox1 && oy0 <= oy1 && oz0 <= oz1) keep &= ~brick_box_mask(ox0, ox1, oy0, oy1, oz0, oz1);
            if (keep == 0ull) continue;
            const uint4 rec = G.bricks[((size_t)L.bz * G.bny + L.by) * G.bnx + L.bx];
            L.occ = ((unsigned long long)rec.y << 32) | rec.x;
            L.base = rec.z;
            L.m = L.occ & keep;
        }
        const int bit = ffs64(L.m) - 1;
        L.m &= L.m - 1ull;
        const int cx = L.bx * 4 + (bit & 3), cy = L.by * 4 + ((bit >> 2) & 3), cz = L.bz * 4 + (bit >> 4);
        const float dx = fmaxf(fmaxf((float)cx - L.gx, L.gx - (float)(cx + 1)) - G.slack, 0.0f);
        const float dy = fmaxf(fmaxf((float)cy - L.gy, L.gy - (float)(cy + 1)) - G.slack, 0.0f);
        const float dz = fmaxf(fmaxf((float)cz - L.gz, L.gz - (float)(cz + 1)) - G.slack, 0.0f);
        if ((dx * dx + dy * dy + dz * dz) * h2 >= L.best_d2) continue;
        const uint32_t ord = L.base + (uint32_t)popc64(L.occ & ((1ull << bit) - 1ull));
        L.p = G.cell_start[ord];
        L.e = G.cell_start[ord + 1];
        return true;
    }
}

// The current pass is exhausted: plan the next one.  Returns false when the query is finished.
PCR_HD bool flat_next_pass(const GridView& G, FlatLane& L) {
    L.cur = L.nb;
    if (L.last) return false;
    float bound = 3.0e38f;
    bool open = false;
    if (L.cur.x0 > 0) { bound = fminf(bound, L.gx - (float)L.cur.x0); open = true; }
    if (L.cur.x1 < G.cnx - 1) { bound = fminf(bound, (float)(L.cur.x1 + 1) - L.gx); open = true; }
    if (L.cur.y0 > 0) { bound = fminf(bound, L.gy - (float)L.cur.y0); open = true; }
    if (L.cur.y1 < G.cny - 1) { bound = fminf(bound, (float)(L.cur.y1 + 1) - L.gy); open = true; }
    if (L.cur.z0 > 0) { bound = fminf(bound, L.gz - (float)L.cur.z0); open = true; }
    if (L.cur.z1 < G.cnz - 1) { bound = fminf(bound, (float)(L.cur.z1 + 1) - L.gz); open = true; }
    if (!open) return false;                                  // whole grid visited
    bound -= G.slack;
    const float rad = sqrtf(L.best_d2) * G.inv_h;             // pruning radius in grid units
    if (rad <= bound) return false;                           // nothing unvisited can be closer
    if (L.best_pos >= 0) {
        const float big = 1.0e9f;
        const float r = rad * 1.000001f + G.slack;            // ball of the best, conservatively inflated
        int v;
        v = cell_of(fminf(fmaxf(L.gx - r, -big), big), G.cnx); L.nb.x0 = v < L.cur.x0 ? v : L.cur.x0;
        v = cell_of(fminf(fmaxf(L.gx + r, -big), big), G.cnx); L.nb.x1 = v > L.cur.x1 ? v : L.cur.x1;
        v = cell_of(fminf(fmaxf(L.gy - r, -big), big), G.cny); L.nb.y0 = v < L.cur.y0 ? v : L.cur.y0;
        v = cell_of(fminf(fmaxf(L.gy + r, -big), big), G.cny); L.nb.y1 = v > L.cur.y1 ? v : L.cur.y1;
        v = cell_of(fminf(fmaxf(L.gz - r, -big), big), G.cnz); L.nb.z0 = v < L.cur.z0 ? v : L.cur.z0;
        v = cell_of(fminf(fmaxf(L.gz + r, -big), big), G.cnz); L.nb.z1 = v > L.cur.z1 ? v : L.cur.z1;
        L.last = 1;
    } else {
        L.nb.x0 = L.cur.x0 > 0 ? L.cur.x0 - 1 : 0; L.nb.x1 = L.cur.x1 < G.cnx - 1 ? L.cur.x1 + 1 : L.cur.x1;
        L.nb.y0 = L.cur.y0 > 0 ? L.cur.y0 - 1 : 0; L.nb.y1 = L.cur.y1 < G.cny - 1 ? L.cur.y1 + 1 : L.cur.y1;
        L.nb.z0 = L.cur.z0 > 0 ? L.cur.z0 - 1 : 0; L.nb.z1 = L.cur.z1 < G.cnz - 1 ? L.cur.z1 + 1 : L.cur.z1;
    }
    flat_reset_iter(L);
    return true;
}

// Phase A of the warp loop for one lane: make [p, e) non-empty or finish the query.
// Returns false when the query is finished (result in best_pos / best_d2).
PCR_HD bool flat_find_work(const GridView& G, FlatLane& L) {
    while (L.p == L.e) {
        if (flat_next_cell(G, L)) continue;       // cells are never empty, but be safe
        if (!flat_next_pass(G, L)) return false;
    }
    return true;
}

// Phase B: evaluate up to `ch` candidates of the current range, four at a time.  `list_idx` null:
// the range [p, e) addresses G.pts directly (a cell); otherwise it addresses list_idx (a per-cell
// candidate list, CandLists).  Groups of four may read past the end of the range: what lies there
// is either another real indexed point (harmless: it can only win if it really is closer) or one
// of the four sentinel records that terminate G.pts / list_idx (infinitely far, never win).
PCR_HD void flat_eval(const GridView& G, FlatLane& L, int ch, const uint32_t* list_idx = nullptr) {
    const uint32_t avail = L.e - L.p;
    const uint32_t n = avail < (uint32_t)ch ? avail : (uint32_t)ch;
    for (uint32_t j = 0; j < n; j += 4) {
        uint32_t p0 = L.p + j, p1 = p0 + 1u, p2 = p0 + 2u, p3 = p0 + 3u;
        if (list_idx) { p0 = list_idx[p0]; p1 = list_idx[p1]; p2 = list_idx[p2]; p3 = list_idx[p3]; }
        const float4 t0 = G.pts[p0], t1 = G.pts[p1], t2 = G.pts[p2], t3 = G.pts[p3];
        float ex, ey, ez, d;
        ex = t0.x - L.qx; ey = t0.y - L.qy; ez = t0.z - L.qz; d = ex * ex + ey * ey + ez * ez;
        if (d < L.best_d2) { L.best_d2 = d; L.best_pos = (int)p0; }
        ex = t1.x - L.qx; ey = t1.y - L.qy; ez = t1.z - L.qz; d = ex * ex + ey * ey + ez * ez;
        if (d < L.best_d2) { L.best_d2 = d; L.best_pos = (int)p1; }
        ex = t2.x - L.qx; ey = t2.y - L.qy; ez = t2.z - L.qz; d = ex * ex + ey * ey + ez * ez;
        if (d < L.best_d2) { L.best_d2 = d; L.best_pos = (int)p2; }
        ex = t3.x - L.qx; ey = t3.y - L.qy; ez = t3.z - L.qz; d = ex * ex + ey * ey + ez * ez;
        if (d < L.best_d2) { L.best_d2 = d; L.best_pos = (int)p3; }
    }
    L.p += n;
}

// Start a query through the per-cell candidate lists (voxel means): on success [p, e) is the
// list of the query's cell and no search is needed; false = no list, use flat_begin().
PCR_HD bool flat_begin_list(const GridView& G, const CandLists& C, FlatLane& L, float qx, float qy, float qz, float max_d2) {
    const float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
    if (!(gx >= 0.0f && gy >= 0.0f && gz >= 0.0f && gx < (float)G.cnx && gy < (float)G.cny && gz < (float)G.cnz)) return false;
    const int cx = (int)gx, cy = (int)gy, cz = (int)gz;
    const uint4 rec = C.bricks[((size_t)(cz >> 2) * G.bny + (cy >> 2)) * G.bnx + (cx >> 2)];
    const unsigned long long band = ((unsigned long long)rec.y << 32) | rec.x;
    const int bit = brick_bit(cx, cy, cz);
    if (!((band >> bit) & 1ull)) return false;
    const uint32_t ord = rec.z + (uint32_t)popc64(band & ((1ull << bit) - 1ull));
    const uint32_t s = C.list_start[ord], e = C.list_start[ord + 1];
    if (s == e) return false;
    L.qx = qx; L.qy = qy; L.qz = qz;
    L.best_d2 = max_d2; L.best_pos = -1;
    L.p = s; L.e = e;
    return true;
}

// Single-lane driver (host replay / reference for the tests): same result as grid_nn().
PCR_HD int flat_nn(const GridView& G, float qx, float qy, float qz, float max_d2, float& out_d2, int ch = 8) {
    FlatLane L;
    if (flat_begin(G, L, qx, qy, qz, max_d2)) {
        while (flat_find_work(G, L)) flat_eval(G, L, ch);
    }
    out_d2 = L.best_d2;
    return L.best_pos;
}

}  // namespace pcr
