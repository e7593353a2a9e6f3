// Exact nearest-neighbour / k-nearest-neighbour search over a brick grid (see GridView in
// pcr_common.cuh).  Replaces pykdtree's KDTree.query at every reference call site:
//   icp.py:33, plane_icp.py:40 (1-NN over target points), voxel.py:176 (1-NN over kept voxel
//   means, quirk Q2), estimate_normals.py:39 (k-NN incl. the point itself, quirk Q5).
//
// Search = "block growing with ball jump":
//   1. visit the query's own cell;
//   2. while nothing has been found, grow the visited block of cells by one ring;
//   3. as soon as a candidate exists, every closer point must lie in a cell that intersects
//      the ball (query, best distance): jump straight to the cell box enclosing that ball,
//      visit only its OCCUPIED cells not visited before (brick masks), prune each by its
//      box distance, and stop.
// It is exact: a cell is skipped only if its (slack-inflated) box is farther than the
// current best, and the search stops only when the unvisited space is farther than the
// best (or than max_dist, quirk Q4: a strict `dist < max_dist` bounds the search).
#pragma once
#include <cstring>
#include "pcr_common.cuh"

namespace pcr {

struct Block3 { int x0, x1, y0, y1, z0, z1; };   // inclusive cell box

// Policy for 1-NN: tracks the best squared distance and the position of the winner in G.pts.
struct Best1 {
    float d2;      // current pruning radius^2 (starts at max_dist^2, strict <)
    int pos;       // position in the sorted point array, -1 = none
    PCR_HD bool have() const { return pos >= 0; }
    PCR_HD float radius2() const { return d2; }
    PCR_HD void offer(float cand_d2, int cand_pos) {
        if (cand_d2 < d2) { d2 = cand_d2; pos = cand_pos; }
    }
};

// Policy for the continuation of a search whose first part was answered by a list: `d2` / `pos`
// hold the list's best, which prunes like a found candidate, but the search keeps growing ring by
// ring (have() is false) until the walk itself finds something closer -- the ball of the list's
// candidate is usually much larger than the one the true neighbour needs.
struct BestSeeded {
    float d2;
    int pos;
    bool improved;
    PCR_HD bool have() const { return improved; }
    PCR_HD float radius2() const { return d2; }
    PCR_HD void offer(float cand_d2, int cand_pos) {
        if (cand_d2 < d2) { d2 = cand_d2; pos = cand_pos; improved = true; }
    }
};

// Policy for k-NN: ascending sorted list in caller-provided storage.
template <int KCAP>
struct BestK {
    float d2s[KCAP];
    int poss[KCAP];
    int k;         // requested neighbours (<= KCAP)
    int cnt;       // found so far
    float lim2;    // max radius^2 (strict <)
    PCR_HD void init(int k_, float lim2_) {
        k = k_; cnt = 0; lim2 = lim2_;
    }
    PCR_HD bool have() const { return cnt >= k; }
    PCR_HD float radius2() const { return cnt >= k ? d2s[k - 1] : lim2; }
    PCR_HD void offer(float cand_d2, int cand_pos) {
        if (!(cand_d2 < radius2())) return;
        int i = (cnt < k) ? cnt : (k - 1);
        if (cnt < k) ++cnt;
        while (i > 0 && d2s[i - 1] > cand_d2) {
            d2s[i] = d2s[i - 1]; poss[i] = poss[i - 1]; --i;
        }
        d2s[i] = cand_d2; poss[i] = cand_pos;
    }
};

// Visit every occupied cell of `nb` that is not inside `ob` (ob may be empty: x1 < x0).
template <class Best>
PCR_HD void visit_block(const GridView& G, float qx, float qy, float qz, float gx, float gy, float gz,
                        const Block3& nb, const Block3& ob, bool have_old, Best& best) {
    const int bx0 = nb.x0 >> 2, bx1 = nb.x1 >> 2;
    const int by0 = nb.y0 >> 2, by1 = nb.y1 >> 2;
    const int bz0 = nb.z0 >> 2, bz1 = nb.z1 >> 2;
    const float h2 = G.h * G.h;
    for (int bz = bz0; bz <= bz1; ++bz) {
        const int lz0 = (nb.z0 > bz * 4 ? nb.z0 - bz * 4 : 0), lz1 = (nb.z1 < bz * 4 + 3 ? nb.z1 - bz * 4 : 3);
        for (int by = by0; by <= by1; ++by) {
            const int ly0 = (nb.y0 > by * 4 ? nb.y0 - by * 4 : 0), ly1 = (nb.y1 < by * 4 + 3 ? nb.y1 - by * 4 : 3);
            for (int bx = bx0; bx <= bx1; ++bx) {
                const int lx0 = (nb.x0 > bx * 4 ? nb.x0 - bx * 4 : 0), lx1 = (nb.x1 < bx * 4 + 3 ? nb.x1 - bx * 4 : 3);
                unsigned long long keep = brick_box_mask(lx0, lx1, ly0, ly1, lz0, lz1);
                if (have_old) {
                    // remove cells already visited (old block intersected with this brick)
                    const int ox0 = (ob.x0 > bx * 4 ? ob.x0 - bx * 4 : 0), ox1 = (ob.x1 < bx * 4 + 3 ? ob.x1 - bx * 4 : 3);
                    const int oy0 = (ob.y0 > by * 4 ? ob.y0 - by * 4 : 0), oy1 = (ob.y1 < by * 4 + 3 ? ob.y1 - by * 4 : 3);
                    const int oz0 = (ob.z0 > bz * 4 ? ob.z0 - bz * 4 : 0), oz1 = (ob.z1 < bz * 4 + 3 ? ob.z1 - bz * 4 : 3);
                    if (ox0 <= ox1 && oy0 <= oy1 && oz0 <= oz1)
                        keep &= ~brick_box_mask(ox0, ox1, oy0, oy1, oz0, oz1);
                    if (keep == 0ull) continue;              // brick lies inside the visited block
                }
                const uint4 rec = G.bricks[((size_t)bz * G.bny + by) * G.bnx + bx];
                const unsigned long long occ = ((unsigned long long)rec.y << 32) | rec.x;
                unsigned long long m = occ & keep;
                while (m) {
                    const int bit = ffs64(m) - 1;
                    m &= m - 1ull;
                    const int cx = bx * 4 + (bit & 3), cy = by * 4 + ((bit >> 2) & 3), cz = bz * 4 + (bit >> 4);
                    // squared distance (grid units) from the query to the slack-inflated cell box
                    float dx = fmaxf(fmaxf((float)cx - gx, gx - (float)(cx + 1)) - G.slack, 0.0f);
                    float dy = fmaxf(fmaxf((float)cy - gy, gy - (float)(cy + 1)) - G.slack, 0.0f);
                    float dz = fmaxf(fmaxf((float)cz - gz, gz - (float)(cz + 1)) - G.slack, 0.0f);
                    if ((dx * dx + dy * dy + dz * dz) * h2 >= best.radius2()) continue;
                    const uint32_t ord = rec.z + (uint32_t)popc64(occ & ((1ull << bit) - 1ull));
                    const uint32_t s = G.cell_start[ord], e = G.cell_start[ord + 1];
                    for (uint32_t p = s; p < e; ++p) {
                        const float4 t = G.pts[p];
                        const float ex = t.x - qx, ey = t.y - qy, ez = t.z - qz;
                        best.offer(dist2_rn(ex, ey, ez), (int)p);
                    }
                }
            }
        }
    }
}

// Same contract as visit_block for SMALL boxes (a few cells per axis): one flattened loop over
// the cells, the brick record cached while consecutive cells share a brick.  Much lighter than
// the brick-mask walk when the ball of the current best meets only a handful of cells -- the
// common case once a candidate is known (late Gauss-Newton iterations, voxel-mean grids).
template <class Best>
PCR_HD void visit_small_box(const GridView& G, float qx, float qy, float qz, float gx, float gy, float gz,
                            const Block3& nb, const Block3& ob, bool have_old, Best& best) {
    const int nx = nb.x1 - nb.x0 + 1, ny = nb.y1 - nb.y0 + 1;
    const int n = nx * ny * (nb.z1 - nb.z0 + 1);
    const float h2 = G.h * G.h;
    long long cur_brick = -1;
    unsigned long long occ = 0ull;
    uint32_t base = 0u;
    int cx = nb.x0, cy = nb.y0, cz = nb.z0;
    for (int k = 0; k < n; ++k) {
        const bool seen = have_old && cx >= ob.x0 && cx <= ob.x1 && cy >= ob.y0 && cy <= ob.y1 && cz >= ob.z0 && cz <= ob.z1;
        if (!seen) {
            const long long b = ((long long)(cz >> 2) * G.bny + (cy >> 2)) * G.bnx + (cx >> 2);
            if (b != cur_brick) {
                const uint4 rec = G.bricks[b];
                occ = ((unsigned long long)rec.y << 32) | rec.x;
                base = rec.z;
                cur_brick = b;
            }
            const int bit = brick_bit(cx, cy, cz);
            if ((occ >> bit) & 1ull) {
                const float dx = fmaxf(fmaxf((float)cx - gx, gx - (float)(cx + 1)) - G.slack, 0.0f);
                const float dy = fmaxf(fmaxf((float)cy - gy, gy - (float)(cy + 1)) - G.slack, 0.0f);
                const float dz = fmaxf(fmaxf((float)cz - gz, gz - (float)(cz + 1)) - G.slack, 0.0f);
                if ((dx * dx + dy * dy + dz * dz) * h2 < best.radius2()) {
                    const uint32_t ord = base + (uint32_t)popc64(occ & ((1ull << bit) - 1ull));
                    const uint32_t s = G.cell_start[ord], e = G.cell_start[ord + 1];
                    for (uint32_t p = s; p < e; ++p) {
                        const float4 t = G.pts[p];
                        const float ex = t.x - qx, ey = t.y - qy, ez = t.z - qz;
                        best.offer(dist2_rn(ex, ey, ez), (int)p);
                    }
                }
            }
        }
        if (++cx > nb.x1) { cx = nb.x0; if (++cy > nb.y1) { cy = nb.y0; ++cz; } }
    }
}

PCR_HD bool is_small_box(const Block3& b) {
    return (b.x1 - b.x0) <= 2 && (b.y1 - b.y0) <= 2 && (b.z1 - b.z0) <= 2;
}

// Second half of the exact search: the cells of `cur` have been visited, `best` holds what was
// found there.  Grows the visited block ring by ring while nothing is known, then visits the cell
// box of the ball (query, best) once.
template <class Best>
PCR_HD void grid_search_continue(const GridView& G, float qx, float qy, float qz, float gx, float gy, float gz, Block3 cur, Best& best) {
    const float big = 1.0e9f;
    for (;;) {
        // distance (grid units) from the query to the nearest face of the visited block that
        // still has unvisited grid cells behind it
        float bound = 3.0e38f;
        bool open = false;
        if (cur.x0 > 0) { bound = fminf(bound, gx - (float)cur.x0); open = true; }
        if (cur.x1 < G.cnx - 1) { bound = fminf(bound, (float)(cur.x1 + 1) - gx); open = true; }
        if (cur.y0 > 0) { bound = fminf(bound, gy - (float)cur.y0); open = true; }
        if (cur.y1 < G.cny - 1) { bound = fminf(bound, (float)(cur.y1 + 1) - gy); open = true; }
        if (cur.z0 > 0) { bound = fminf(bound, gz - (float)cur.z0); open = true; }
        if (cur.z1 < G.cnz - 1) { bound = fminf(bound, (float)(cur.z1 + 1) - gz); open = true; }
        if (!open) return;                                   // whole grid visited
        bound -= G.slack;
        const float rad = sqrtf(best.radius2()) * G.inv_h;   // current pruning radius in grid units
        if (rad <= bound) return;                            // nothing unvisited can be closer
        Block3 nb;
        if (best.have()) {
            // enclose the ball (query, radius) -- conservative by slack and a relative epsilon
            const float r = rad * 1.000001f + G.slack;
            nb.x0 = cell_of(fminf(fmaxf(gx - r, -big), big), G.cnx); nb.x1 = cell_of(fminf(fmaxf(gx + r, -big), big), G.cnx);
            nb.y0 = cell_of(fminf(fmaxf(gy - r, -big), big), G.cny); nb.y1 = cell_of(fminf(fmaxf(gy + r, -big), big), G.cny);
            nb.z0 = cell_of(fminf(fmaxf(gz - r, -big), big), G.cnz); nb.z1 = cell_of(fminf(fmaxf(gz + r, -big), big), G.cnz);
            nb.x0 = nb.x0 < cur.x0 ? nb.x0 : cur.x0; nb.x1 = nb.x1 > cur.x1 ? nb.x1 : cur.x1;
            nb.y0 = nb.y0 < cur.y0 ? nb.y0 : cur.y0; nb.y1 = nb.y1 > cur.y1 ? nb.y1 : cur.y1;
            nb.z0 = nb.z0 < cur.z0 ? nb.z0 : cur.z0; nb.z1 = nb.z1 > cur.z1 ? nb.z1 : cur.z1;
            if (is_small_box(nb)) visit_small_box(G, qx, qy, qz, gx, gy, gz, nb, cur, true, best);
            else visit_block(G, qx, qy, qz, gx, gy, gz, nb, cur, true, best);
            return;                                          // ball fully covered
        }
        nb.x0 = cur.x0 > 0 ? cur.x0 - 1 : 0; nb.x1 = cur.x1 < G.cnx - 1 ? cur.x1 + 1 : cur.x1;
        nb.y0 = cur.y0 > 0 ? cur.y0 - 1 : 0; nb.y1 = cur.y1 < G.cny - 1 ? cur.y1 + 1 : cur.y1;
        nb.z0 = cur.z0 > 0 ? cur.z0 - 1 : 0; nb.z1 = cur.z1 < G.cnz - 1 ? cur.z1 + 1 : cur.z1;
        if (is_small_box(nb)) visit_small_box(G, qx, qy, qz, gx, gy, gz, nb, cur, true, best);
        else visit_block(G, qx, qy, qz, gx, gy, gz, nb, cur, true, best);
        cur = nb;
    }
}

// Generic exact search.  `best` carries the initial radius (max_dist^2) and receives results.
// ball_first: the caller knows that nothing lies near the query (its cell's list was exhausted, or
// the cell is not even close to an occupied one): skip the ring-by-ring growth, which would re-walk
// the neighbourhood once per ring, and visit the cell box of the ball (query, radius) in ONE pruned
// pass -- provided that box is small (max_dist of a few cells); otherwise grow rings as usual.
template <class Best>
PCR_HD void grid_search(const GridView& G, float qx, float qy, float qz, Best& best, bool ball_first = false) {
    if (G.n_pts == 0) return;
    const float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
    if (!(gx == gx) || !(gy == gy) || !(gz == gz)) return;            // NaN query: no match
    // distance from the query to the grid's bounding box; nothing can match beyond the radius
    {
        float ex = fmaxf(fmaxf(-gx, gx - (float)G.cnx), 0.0f);
        float ey = fmaxf(fmaxf(-gy, gy - (float)G.cny), 0.0f);
        float ez = fmaxf(fmaxf(-gz, gz - (float)G.cnz), 0.0f);
        float e = fmaxf(sqrtf(ex * ex + ey * ey + ez * ez) - G.slack, 0.0f) * G.h;
        if (e * e >= best.radius2()) return;
    }
    // clamp huge coordinates before the int conversion
    const float big = 1.0e9f;
    const float cgx = fminf(fmaxf(gx, -big), big), cgy = fminf(fmaxf(gy, -big), big), cgz = fminf(fmaxf(gz, -big), big);
    Block3 cur;
    Block3 none; none.x0 = none.y0 = none.z0 = 0; none.x1 = none.y1 = none.z1 = -1;
    if (best.have() || ball_first) {
        // warm start: the caller already holds a candidate (an upper bound), or asks for the ball of
        // the search radius.  Every closer point lies in a cell meeting the ball (query, bound):
        // visit exactly that box and stop.
        const float r = sqrtf(best.radius2()) * G.inv_h * 1.000001f + G.slack;
        cur.x0 = cell_of(fminf(fmaxf(gx - r, -big), big), G.cnx); cur.x1 = cell_of(fminf(fmaxf(gx + r, -big), big), G.cnx);
        cur.y0 = cell_of(fminf(fmaxf(gy - r, -big), big), G.cny); cur.y1 = cell_of(fminf(fmaxf(gy + r, -big), big), G.cny);
        cur.z0 = cell_of(fminf(fmaxf(gz - r, -big), big), G.cnz); cur.z1 = cell_of(fminf(fmaxf(gz + r, -big), big), G.cnz);
        // measured (profiles/r1_sweep11_ball_first.log): pays when the ball spans <= 6 cells per axis (NDT, 1 m
        // voxels, max_dist 2 m: 0.54 -> 0.46 ms in iteration 1); larger balls are cheaper ring by ring
        const bool small_ball = (cur.x1 - cur.x0) < 6 && (cur.y1 - cur.y0) < 6 && (cur.z1 - cur.z0) < 6;
        if (best.have() || small_ball) {
            if (is_small_box(cur)) visit_small_box(G, qx, qy, qz, gx, gy, gz, cur, none, false, best);
            else visit_block(G, qx, qy, qz, gx, gy, gz, cur, none, false, best);
            return;
        }
    }
    cur.x0 = cur.x1 = cell_of(cgx, G.cnx);
    cur.y0 = cur.y1 = cell_of(cgy, G.cny);
    cur.z0 = cur.z1 = cell_of(cgz, G.cnz);
    visit_small_box(G, qx, qy, qz, gx, gy, gz, cur, none, false, best);
    grid_search_continue(G, qx, qy, qz, gx, gy, gz, cur, best);
}

// 1-NN convenience wrapper: returns position in G.pts (or -1) and the squared distance.
PCR_HD int grid_nn(const GridView& G, float qx, float qy, float qz, float max_d2, float& out_d2, bool ball_first = false) {
    Best1 b; b.d2 = max_d2; b.pos = -1;
    grid_search(G, qx, qy, qz, b, ball_first);
    out_d2 = b.d2;
    return b.pos;
}

// Four squared distances of one structure-of-arrays group (see ShellLists) to the query, in packed
// f32x2 arithmetic on sm_100 (same roundings as dist2_rn); updates the best and its list offset.
PCR_HD void shell_eval_group(const float4& X, const float4& Y, const float4& Z, float qx, float qy, float qz, uint32_t k,
                             float& best, uint32_t& best_k, uint32_t& best_j) {
    float d0, d1, d2, d3;
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    const float2 nx = make_float2(-qx, -qx), ny = make_float2(-qy, -qy), nz = make_float2(-qz, -qz);
    const float2 ex01 = __fadd2_rn(make_float2(X.x, X.y), nx), ex23 = __fadd2_rn(make_float2(X.z, X.w), nx);
    const float2 ey01 = __fadd2_rn(make_float2(Y.x, Y.y), ny), ey23 = __fadd2_rn(make_float2(Y.z, Y.w), ny);
    const float2 ez01 = __fadd2_rn(make_float2(Z.x, Z.y), nz), ez23 = __fadd2_rn(make_float2(Z.z, Z.w), nz);
    const float2 r01 = __ffma2_rn(ez01, ez01, __ffma2_rn(ey01, ey01, __fmul2_rn(ex01, ex01)));
    const float2 r23 = __ffma2_rn(ez23, ez23, __ffma2_rn(ey23, ey23, __fmul2_rn(ex23, ex23)));
    d0 = r01.x; d1 = r01.y; d2 = r23.x; d3 = r23.y;
#else
    d0 = dist2_rn(X.x - qx, Y.x - qy, Z.x - qz); d1 = dist2_rn(X.y - qx, Y.y - qy, Z.y - qz);
    d2 = dist2_rn(X.z - qx, Y.z - qy, Z.z - qz); d3 = dist2_rn(X.w - qx, Y.w - qy, Z.w - qz);
#endif
    const float dm = fminf(fminf(d0, d1), fminf(d2, d3));
    if (dm < best) {                                          // rare after the first groups
        best = dm;
        best_k = k;
        best_j = dm == d0 ? 0u : (dm == d1 ? 1u : (dm == d2 ? 2u : 3u));
    }
}

// Cursor over the shell list (see ShellLists) of one query's cell.
struct ShellCursor {
    uint32_t k;                // next group of the list (offsets count GROUPS of four entries)
    uint32_t e;                // end of the list
    float m;                   // margin bound of the next group (already loaded)
    float best;                // best squared distance so far (starts at max_dist^2, strict <)
    uint32_t best_k;           // group of the best entry, 0xffffffff = none
    uint32_t best_j;           // ... and its place (0..3) in that group
    bool active;               // more groups to evaluate
    bool exhausted;            // the list ended (or will end) without a margin bound stopping the scan
};

// Open the list of the query's cell.  false: the cell has no list (nothing was looked at).
PCR_HD bool shell_open(const GridView& G, const ShellLists& S, float qx, float qy, float qz, float max_d2, ShellCursor& c) {
    c.active = false; c.exhausted = true; c.best = max_d2; c.best_k = 0xffffffffu; c.best_j = 0u;
    const float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
    if (!(gx >= 0.0f && gy >= 0.0f && gz >= 0.0f && gx < (float)G.cnx && gy < (float)G.cny && gz < (float)G.cnz)) return false;
    const int cx = (int)gx, cy = (int)gy, cz = (int)gz;
    const uint4 rec = S.bricks[((size_t)(cz >> 2) * G.bny + (cy >> 2)) * G.bnx + (cx >> 2)];
    const unsigned long long band = ((unsigned long long)rec.y << 32) | rec.x;
    const int bit = brick_bit(cx, cy, cz);
    if (!((band >> bit) & 1ull)) return false;
    const uint32_t ord = rec.z + (uint32_t)popc64(band & ((1ull << bit) - 1ull));
    const uint32_t s = S.start[ord], e = S.start[ord + 1];
    c.k = s; c.e = e;
    c.m = S.margin2[s];                                       // (one bound past the end of the array is allocated)
    if (s < e) {
        if (c.m >= c.best) c.exhausted = false;               // not even the first group can hold a match
        else c.active = true;
    }
    return true;
}

// Advance an active cursor past the group just evaluated; termination tests.
PCR_HD void shell_advance(ShellCursor& c, float m_next) {
    c.k += 1; c.m = m_next;
    if (!(c.k < c.e)) c.active = false;                       // list ended
    else if (c.m >= c.best) { c.active = false; c.exhausted = false; }   // everything from here on is at least this far
}

// Result of a finished cursor: 1 = final (out_pos = position in GridView::pts or -1: none within
// max_dist), 2 = the list ended while the best is still beyond the covered margin: out_* hold an
// upper bound and the general search has to finish the job.
PCR_HD int shell_close(const ShellLists& S, const ShellCursor& c, float& out_d2, int& out_pos) {
    out_d2 = c.best;
    out_pos = -1;
    if (c.best_k != 0xffffffffu) {
        const float4 W = S.pts[4 * (size_t)c.best_k + 3];
        const uint32_t j = c.best_j;
        const float w = j == 0u ? W.x : (j == 1u ? W.y : (j == 2u ? W.z : W.w));
#if defined(__CUDA_ARCH__)
        out_pos = __float_as_int(w);
#else
        memcpy(&out_pos, &w, 4);
#endif
    }
    return (c.exhausted && !(c.best <= S.covered2)) ? 2 : 1;
}

// Stream the shell list of ONE query's cell: 0 = no list, else see shell_close.
PCR_HD int shell_scan(const GridView& G, const ShellLists& S, float qx, float qy, float qz, float max_d2, float& out_d2, int& out_pos) {
    ShellCursor c;
    if (!shell_open(G, S, qx, qy, qz, max_d2, c)) return 0;
    while (c.active) {
        const float4* g4 = S.pts + 4 * (size_t)c.k;
        const float4 X = g4[0], Y = g4[1], Z = g4[2];
        const float mn = S.margin2[c.k + 1];
        shell_eval_group(X, Y, Z, qx, qy, qz, c.k, c.best, c.best_k, c.best_j);
        shell_advance(c, mn);
    }
    return shell_close(S, c, out_d2, out_pos);
}

// Continuation of a list scan that ended with status 2: the list evaluated every point of the
// (2 block_r + 1)^3 cell block around the query's cell, so the brick-grid search resumes OUTSIDE
// that block, ring by ring, pruned by the list's best (in/out: out_d2, out_pos).
PCR_HD void shell_continue(const GridView& G, const ShellLists& S, float qx, float qy, float qz, float& out_d2, int& out_pos) {
    const float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
    const int cx = (int)gx, cy = (int)gy, cz = (int)gz;       // inside the grid: shell_open accepted the query
    const int r = S.block_r;
    Block3 cur;
    cur.x0 = cx - r > 0 ? cx - r : 0; cur.x1 = cx + r < G.cnx - 1 ? cx + r : G.cnx - 1;
    cur.y0 = cy - r > 0 ? cy - r : 0; cur.y1 = cy + r < G.cny - 1 ? cy + r : G.cny - 1;
    cur.z0 = cz - r > 0 ? cz - r : 0; cur.z1 = cz + r < G.cnz - 1 ? cz + r : G.cnz - 1;
    BestSeeded b;
    b.d2 = out_d2; b.pos = out_pos; b.improved = false;
    grid_search_continue(G, qx, qy, qz, gx, gy, gz, cur, b);
    out_d2 = b.d2; out_pos = b.pos;
}

// 1-NN through the shell lists with the brick-grid search as continuation (what the kernel does
// per scan slot; also the host replay).  false: the cell has no list.
PCR_HD bool shell_nn(const GridView& G, const ShellLists& S, float qx, float qy, float qz, float max_d2, float& out_d2, int& out_pos) {
    const int st = shell_scan(G, S, qx, qy, qz, max_d2, out_d2, out_pos);
    if (st == 0) return false;
    if (st == 2) shell_continue(G, S, qx, qy, qz, out_d2, out_pos);
    return true;
}

// 1-NN through the per-cell candidate lists (see CandLists); returns false when the query's cell
// has no list and the caller must run the general search.
PCR_HD bool list_nn(const GridView& G, const CandLists& L, float qx, float qy, float qz, float max_d2,
                    float& out_d2, int& out_pos) {
    const float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
    if (!(gx >= 0.0f && gy >= 0.0f && gz >= 0.0f && gx < (float)G.cnx && gy < (float)G.cny && gz < (float)G.cnz)) return false;
    const int cx = (int)gx, cy = (int)gy, cz = (int)gz;
    const uint4 rec = L.bricks[((size_t)(cz >> 2) * G.bny + (cy >> 2)) * G.bnx + (cx >> 2)];
    const unsigned long long band = ((unsigned long long)rec.y << 32) | rec.x;
    const int bit = brick_bit(cx, cy, cz);
    if (!((band >> bit) & 1ull)) return false;
    const uint32_t ord = rec.z + (uint32_t)popc64(band & ((1ull << bit) - 1ull));
    const uint32_t s = L.list_start[ord], e = L.list_start[ord + 1];
    if (s == e) return false;
    float best = max_d2;
    int pos = -1;
    if (L.list_pts) {
        uint32_t bk = 0xffffffffu;
        for (uint32_t k = s; k < e; ++k) {
            const float4 t = L.list_pts[k];
            const float d2 = dist2_rn(t.x - qx, t.y - qy, t.z - qz);
            if (d2 < best) { best = d2; bk = k; }
        }
        if (bk != 0xffffffffu) pos = (int)L.list_idx[bk];
    } else {
        for (uint32_t k = s; k < e; ++k) {
            const uint32_t p = L.list_idx[k];
            const float4 t = G.pts[p];
            const float ex = t.x - qx, ey = t.y - qy, ez = t.z - qz;
            const float d2 = dist2_rn(ex, ey, ez);
            if (d2 < best) { best = d2; pos = (int)p; }
        }
    }
    out_d2 = best;
    out_pos = pos;
    return true;
}

// 1-NN with a warm start: `warm_pos` is the position of any indexed point (e.g. the previous
// iteration's match, -1 = none); its distance bounds the search.
PCR_HD int grid_nn_warm(const GridView& G, float qx, float qy, float qz, float max_d2, int warm_pos, float& out_d2) {
    Best1 b; b.d2 = max_d2; b.pos = -1;
    if (warm_pos >= 0) {
        const float4 t = G.pts[warm_pos];
        const float ex = t.x - qx, ey = t.y - qy, ez = t.z - qz;
        const float dw = dist2_rn(ex, ey, ez);
        if (dw < max_d2) { b.d2 = dw; b.pos = warm_pos; }
    }
    grid_search(G, qx, qy, qz, b);
    out_d2 = b.d2;
    return b.pos;
}

}  // namespace pcr
