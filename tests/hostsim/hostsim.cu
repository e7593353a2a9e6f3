// TEST INFRASTRUCTURE ONLY.  Host-side replay of the PCR_HD device functions (grid search,
// per-point terms, small linear algebra) so that their index arithmetic and algebra can be
// checked against the oracle on a machine without a GPU.  Built by tests/hostsim/build.py into
// tests/hostsim/_hostsim.so; never loaded by the product package.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../point_cloud_registration_b200/csrc/pcr_common.cuh"
#include "../../point_cloud_registration_b200/csrc/pcr_grid.cuh"
#include "../../point_cloud_registration_b200/csrc/pcr_linalg.cuh"
#include "../../point_cloud_registration_b200/csrc/pcr_terms.cuh"
#include "../../point_cloud_registration_b200/csrc/pcr_tile.cuh"

using namespace pcr;

struct HostGrid {
    GridView v{};
    std::vector<uint4> bricks;
    std::vector<uint32_t> cell_start;
    std::vector<float4> pts;
    // per-cell shell lists (host mirror of build_shell_lists in pcr_build.cu)
    ShellLists shell{};
    std::vector<uint4> shell_bricks;
    std::vector<uint32_t> shell_start;
    std::vector<float4> shell_pts;
    std::vector<float> shell_margin2;
};

extern "C" {

// Build a brick grid on the host with the same binning arithmetic as point_key_kernel.
void* hs_grid_build(const float* xyz, int64_t n, double h) {
    HostGrid* g = new HostGrid();
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    double maxabs = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], xyz[3 * i + a]); hi[a] = std::max(hi[a], xyz[3 * i + a]);
            maxabs = std::max(maxabs, (double)fabsf(xyz[3 * i + a]));
        }
    GridView& V = g->v;
    V.h = (float)h; V.inv_h = (float)(1.0 / h);
    V.ox = lo[0] - 2.25f * V.h; V.oy = lo[1] - 2.25f * V.h; V.oz = lo[2] - 2.25f * V.h;
    double c[3];
    for (int a = 0; a < 3; ++a) c[a] = floor(((double)hi[a] - (double)(lo[a] - 2.25f * V.h)) / h) + 4.0;
    V.bnx = (int)(c[0] / 4.0 + 1.0); V.bny = (int)(c[1] / 4.0 + 1.0); V.bnz = (int)(c[2] / 4.0 + 1.0);
    V.cnx = V.bnx * 4; V.cny = V.bny * 4; V.cnz = V.bnz * 4;
    V.slack = 1e-3f + 1e-6f * (float)(2.0 * maxabs / h + (double)std::max(V.cnx, std::max(V.cny, V.cnz)));
    V.n_pts = (uint32_t)n;
    std::vector<unsigned long long> keys(n);
    for (int64_t i = 0; i < n; ++i) {
        float gx = (xyz[3 * i] - V.ox) * V.inv_h, gy = (xyz[3 * i + 1] - V.oy) * V.inv_h, gz = (xyz[3 * i + 2] - V.oz) * V.inv_h;
        int cx = cell_of(gx, V.cnx), cy = cell_of(gy, V.cny), cz = cell_of(gz, V.cnz);
        unsigned long long brick = ((unsigned long long)(cz >> 2) * V.bny + (cy >> 2)) * V.bnx + (cx >> 2);
        keys[i] = brick * 64ull + (unsigned long long)brick_bit(cx, cy, cz);
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    g->bricks.assign((size_t)V.bnx * V.bny * V.bnz, make_uint4(0, 0, 0, 0));
    g->pts.resize(n);
    uint32_t ord = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t j = order[i];
        float w;
        memcpy(&w, &j, 4);
        g->pts[i] = make_float4(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2], w);
        const unsigned long long key = keys[j];
        if (i == 0 || key != keys[order[i - 1]]) {
            g->cell_start.push_back((uint32_t)i);
            const unsigned long long brick = key >> 6;
            unsigned long long mask = ((unsigned long long)g->bricks[brick].y << 32) | g->bricks[brick].x;
            if (mask == 0ull) g->bricks[brick].z = ord;
            mask |= 1ull << (key & 63ull);
            g->bricks[brick].x = (uint32_t)mask; g->bricks[brick].y = (uint32_t)(mask >> 32);
            ++ord;
        }
    }
    g->cell_start.push_back((uint32_t)n);
    for (int k = 0; k < 4; ++k) g->pts.push_back(make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f));   // sentinels (see pad_tail_kernel)
    V.bricks = g->bricks.data(); V.cell_start = g->cell_start.data(); V.pts = g->pts.data();
    return g;
}

void hs_grid_free(void* g) { delete (HostGrid*)g; }
int64_t hs_grid_cells(void* g) { return (int64_t)((HostGrid*)g)->cell_start.size() - 1; }

void hs_nn(void* gp, const float* q, int64_t m, double max_dist, int64_t* idx, float* dist) {
    const GridView& G = ((HostGrid*)gp)->v;
    const float md = (float)max_dist;
    for (int64_t i = 0; i < m; ++i) {
        float d2;
        int pos = grid_nn(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], md * md, d2);
        if (pos >= 0) {
            uint32_t j;
            memcpy(&j, &G.pts[pos].w, 4);
            idx[i] = j; dist[i] = sqrtf(d2);
        } else { idx[i] = -1; dist[i] = INFINITY; }
    }
}

// grid_nn with ball_first (the kernels' path for list misses)
void hs_nn_ball_first(void* gp, const float* q, int64_t m, double max_dist, int64_t* idx, float* dist) {
    const GridView& G = ((HostGrid*)gp)->v;
    const float md = (float)max_dist;
    for (int64_t i = 0; i < m; ++i) {
        float d2;
        int pos = grid_nn(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], md * md, d2, true);
        if (pos >= 0) {
            uint32_t j;
            memcpy(&j, &G.pts[pos].w, 4);
            idx[i] = j; dist[i] = sqrtf(d2);
        } else { idx[i] = -1; dist[i] = INFINITY; }
    }
}

void hs_knn(void* gp, const float* q, int64_t m, int k, int64_t* idx, float* dist) {
    const GridView& G = ((HostGrid*)gp)->v;
    for (int64_t i = 0; i < m; ++i) {
        BestK<64> best;
        best.init(k, 3.0e38f);
        grid_search(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], best);
        for (int r = 0; r < k; ++r) {
            if (r < best.cnt) {
                uint32_t j;
                memcpy(&j, &G.pts[best.poss[r]].w, 4);
                idx[i * k + r] = j; dist[i * k + r] = sqrtf(best.d2s[r]);
            } else { idx[i * k + r] = -1; dist[i * k + r] = INFINITY; }
        }
    }
}



// Host mirror of band_mark_kernel(dilate 1) + shell_build_kernel: same band, levels, layout.
int64_t hs_shell_build(void* gp, double dmax_frac) {
    HostGrid* g = (HostGrid*)gp;
    const GridView& G = g->v;
    static const float frac[PCR_SHELL_LEVELS] = {PCR_SHELL_FRACS};
    const int R = dmax_frac <= 1.0 ? 1 : (dmax_frac <= 2.0 ? 2 : 3);
    const int side = 2 * R + 1, ncell = side * side * side, centre = (ncell - 1) / 2;
    const float dmax = (float)(dmax_frac * (double)G.h);
    const size_t nb = (size_t)G.bnx * G.bny * G.bnz;
    std::vector<unsigned long long> band(nb, 0ull);
    auto occupied = [&](int x, int y, int z, uint32_t* ord) {
        const uint4 rec = g->bricks[((size_t)(z >> 2) * G.bny + (y >> 2)) * G.bnx + (x >> 2)];
        const unsigned long long occ = ((unsigned long long)rec.y << 32) | rec.x;
        const int bit = brick_bit(x, y, z);
        if (!((occ >> bit) & 1ull)) return false;
        if (ord) *ord = rec.z + (uint32_t)popc64(occ & ((1ull << bit) - 1ull));
        return true;
    };
    for (int z = 0; z < G.cnz; ++z) for (int y = 0; y < G.cny; ++y) for (int x = 0; x < G.cnx; ++x) {
        if (!occupied(x, y, z, nullptr)) continue;
        for (int dz = -R; dz <= R; ++dz) for (int dy = -R; dy <= R; ++dy) for (int dx = -R; dx <= R; ++dx) {
            const int nx = x + dx, ny = y + dy, nz = z + dz;
            if (nx < 0 || ny < 0 || nz < 0 || nx >= G.cnx || ny >= G.cny || nz >= G.cnz) continue;
            band[((size_t)(nz >> 2) * G.bny + (ny >> 2)) * G.bnx + (nx >> 2)] |= 1ull << brick_bit(nx, ny, nz);
        }
    }
    g->shell_bricks.assign(nb, make_uint4(0, 0, 0, 0));
    g->shell_start.clear(); g->shell_pts.clear(); g->shell_margin2.clear();
    const float slack_w = G.slack * G.h;
    uint32_t ord = 0;
    for (size_t b = 0; b < nb; ++b) {
        g->shell_bricks[b] = make_uint4((uint32_t)band[b], (uint32_t)(band[b] >> 32), ord, 0);
        const int bx = (int)(b % G.bnx), by = (int)((b / G.bnx) % G.bny), bz = (int)(b / ((size_t)G.bnx * G.bny));
        for (int bit = 0; bit < 64; ++bit) {
            if (!((band[b] >> bit) & 1ull)) continue;
            const int cx = bx * 4 + (bit & 3), cy = by * 4 + ((bit >> 2) & 3), cz = bz * 4 + (bit >> 4);
            const float lox = G.ox + (float)cx * G.h, loy = G.oy + (float)cy * G.h, loz = G.oz + (float)cz * G.h;
            const float hix = lox + G.h, hiy = loy + G.h, hiz = loz + G.h;
            std::vector<float4> lv[PCR_SHELL_LEVELS];
            for (int oo = 0; oo < ncell; ++oo) {
                const int o = oo == 0 ? centre : (oo <= centre ? oo - 1 : oo);
                const int nx = cx + o % side - R, ny = cy + (o / side) % side - R, nz = cz + o / (side * side) - R;
                if (nx < 0 || ny < 0 || nz < 0 || nx >= G.cnx || ny >= G.cny || nz >= G.cnz) continue;
                uint32_t o2;
                if (!occupied(nx, ny, nz, &o2)) continue;
                for (uint32_t p = g->cell_start[o2]; p < g->cell_start[o2 + 1]; ++p) {
                    const float4 t = g->pts[p];
                    int lvl = 0;
                    if (oo != 0) {
                        const float mx = fmaxf(fmaxf(lox - t.x, t.x - hix), 0.0f), my = fmaxf(fmaxf(loy - t.y, t.y - hiy), 0.0f),
                                    mz = fmaxf(fmaxf(loz - t.z, t.z - hiz), 0.0f);
                        const float m2 = mx * mx + my * my + mz * mz;
                        if (m2 > dmax * dmax) continue;
                        lvl = PCR_SHELL_LEVELS - 1;
                        for (int l = PCR_SHELL_LEVELS - 2; l >= 1; --l) if (m2 <= (frac[l] * G.h) * (frac[l] * G.h)) lvl = l;
                    }
                    float w;
                    memcpy(&w, &p, 4);
                    lv[lvl].push_back(make_float4(t.x, t.y, t.z, w));
                }
            }
            g->shell_start.push_back((uint32_t)(g->shell_pts.size() / 4));      // offsets count groups of four entries
            std::vector<float4> flat;
            std::vector<float> bounds;
            for (int l = 0; l < PCR_SHELL_LEVELS; ++l) {
                for (size_t k = 0; k < lv[l].size(); ++k) {
                    if ((flat.size() & 3u) == 0u) {
                        const float lb = l >= 1 ? fmaxf(frac[l - 1] * G.h - slack_w, 0.0f) : 0.0f;
                        bounds.push_back(lb * lb);
                    }
                    flat.push_back(lv[l][k]);
                }
            }
            while (flat.size() & 3u) flat.push_back(make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f));
            for (size_t g4 = 0; g4 < flat.size(); g4 += 4) {      // groups of four, structure of arrays
                g->shell_pts.push_back(make_float4(flat[g4].x, flat[g4 + 1].x, flat[g4 + 2].x, flat[g4 + 3].x));
                g->shell_pts.push_back(make_float4(flat[g4].y, flat[g4 + 1].y, flat[g4 + 2].y, flat[g4 + 3].y));
                g->shell_pts.push_back(make_float4(flat[g4].z, flat[g4 + 1].z, flat[g4 + 2].z, flat[g4 + 3].z));
                g->shell_pts.push_back(make_float4(flat[g4].w, flat[g4 + 1].w, flat[g4 + 2].w, flat[g4 + 3].w));
            }
            for (float b2 : bounds) g->shell_margin2.push_back(b2);
            ++ord;
        }
    }
    g->shell_start.push_back((uint32_t)(g->shell_pts.size() / 4));
    const int64_t n_entries = (int64_t)g->shell_pts.size();
    for (int k = 0; k < 4; ++k) g->shell_pts.push_back(make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f));
    g->shell_margin2.push_back(3.0e38f);
    g->shell.bricks = g->shell_bricks.data(); g->shell.start = g->shell_start.data(); g->shell.pts = g->shell_pts.data();
    g->shell.margin2 = g->shell_margin2.data();
    const float cov = fmaxf(dmax - slack_w, 0.0f);
    g->shell.covered2 = cov * cov;
    g->shell.block_r = (double)dmax >= 1.7320508 * (double)G.h * 1.0001 ? 1 : 0;
    return n_entries;
}

// 1-NN through the shell lists (general search when the cell has no list), as the kernel does.
// used_list[i] = 1 when the list path answered.
void hs_shell_nn(void* gp, const float* q, int64_t m, double max_dist, int64_t* idx, float* dist, uint8_t* used_list, int pair) {
    HostGrid* g = (HostGrid*)gp;
    const GridView& G = g->v;
    const float md = (float)max_dist;
    for (int64_t i = 0; i < m; ++i) {
        float d2;
        int pos;
        bool ok;
        if (pair) {
            // the cursor API used step by step, as a kernel would interleave it with other work
            ShellCursor c;
            ok = shell_open(G, g->shell, q[3 * i], q[3 * i + 1], q[3 * i + 2], md * md, c);
            int st = 0;
            if (ok) {
                while (c.active) {
                    const float4* g4 = g->shell.pts + 4 * (size_t)c.k;
                    shell_eval_group(g4[0], g4[1], g4[2], q[3 * i], q[3 * i + 1], q[3 * i + 2], c.k, c.best, c.best_k, c.best_j);
                    shell_advance(c, g->shell.margin2[c.k + 1]);
                }
                st = shell_close(g->shell, c, d2, pos);
            }
            if (st == 2) shell_continue(G, g->shell, q[3 * i], q[3 * i + 1], q[3 * i + 2], d2, pos);
        } else {
            ok = shell_nn(G, g->shell, q[3 * i], q[3 * i + 1], q[3 * i + 2], md * md, d2, pos);
        }
        if (!ok) pos = grid_nn(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], md * md, d2);
        if (used_list) used_list[i] = ok ? 1 : 0;
        if (pos >= 0) {
            uint32_t j;
            memcpy(&j, &G.pts[pos].w, 4);
            idx[i] = j; dist[i] = sqrtf(d2);
        } else { idx[i] = -1; dist[i] = INFINITY; }
    }
}

// Development aid: status of the list scan per query (0 no list, 1 settled, 2 exhausted) and its distance.
void hs_shell_status(void* gp, const float* q, int64_t m, double max_dist, int32_t* status, float* dist) {
    HostGrid* g = (HostGrid*)gp;
    const float md = (float)max_dist;
    for (int64_t i = 0; i < m; ++i) {
        float d2; int pos;
        status[i] = shell_scan(g->v, g->shell, q[3 * i], q[3 * i + 1], q[3 * i + 2], md * md, d2, pos);
        dist[i] = pos >= 0 ? sqrtf(d2) : INFINITY;
    }
}

// Development aid: per warp row of 32 consecutive queries, groups of four evaluated by the
// shell-list loop: out[0] = sum over rows of the max over lanes (lock-step cost), out[1] = sum
// over all queries, out[2] = queries that fell back to the general search after the list,
// out[3] = queries without a list.
void hs_shell_study(void* gp, const float* q, int64_t m, double max_dist, double* out) {
    HostGrid* g = (HostGrid*)gp;
    const GridView& G = g->v;
    const ShellLists& S = g->shell;
    const float md2 = (float)max_dist * (float)max_dist;
    out[0] = out[1] = out[2] = out[3] = 0;
    for (int64_t r = 0; r < m; r += 32) {
        long long mx = 0;
        for (int64_t i = r; i < std::min(m, r + 32); ++i) {
            const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
            const float gx = (qx - G.ox) * G.inv_h, gy = (qy - G.oy) * G.inv_h, gz = (qz - G.oz) * G.inv_h;
            if (!(gx >= 0.0f && gy >= 0.0f && gz >= 0.0f && gx < (float)G.cnx && gy < (float)G.cny && gz < (float)G.cnz)) { out[3] += 1; continue; }
            const int cx = (int)gx, cy = (int)gy, cz = (int)gz;
            const uint4 rec = S.bricks[((size_t)(cz >> 2) * G.bny + (cy >> 2)) * G.bnx + (cx >> 2)];
            const unsigned long long band = ((unsigned long long)rec.y << 32) | rec.x;
            const int bit = brick_bit(cx, cy, cz);
            if (!((band >> bit) & 1ull)) { out[3] += 1; continue; }
            const uint32_t ord = rec.z + (uint32_t)popc64(band & ((1ull << bit) - 1ull));
            float best = md2;
            long long groups = 0;
            uint32_t k = S.start[ord];
            const uint32_t e = S.start[ord + 1];
            for (; k < e; k += 1) {
                if (S.margin2[k] >= best) break;
                ++groups;
                {
                    const float4 X = S.pts[4 * (size_t)k], Y = S.pts[4 * (size_t)k + 1], Z = S.pts[4 * (size_t)k + 2];
                    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
                    for (int u = 0; u < 4; ++u) {
                        const float d = dist2_rn(xs[u] - qx, ys[u] - qy, zs[u] - qz);
                        if (d < best) best = d;
                    }
                }
            }
            if (k >= e && !(best <= S.covered2)) out[2] += 1;
            out[1] += (double)groups;
            mx = std::max(mx, groups);
        }
        out[0] += (double)mx;
    }
}

// Per-point terms with given correspondences.  rec_in: per scan point the matched record
//   ICP: q(3) | PLANE/VPLANE: q(3), n(3) | NDT: mu(3), W6(6)   (float32, stride 9), ok flag.
void hs_linearize(int method, const double* T, const float* scan, int64_t n, const float* recs, const uint8_t* ok,
                  double* out29) {
    Pose32 P;
    pose32_from_T(T, P);
    double acc[PCR_NEQ_PAD] = {0};
    for (int64_t i = 0; i < n; ++i) {
        if (!ok[i]) continue;
        const float px = scan[3 * i], py = scan[3 * i + 1], pz = scan[3 * i + 2];
        float qx, qy, qz;
        transform32(P, px, py, pz, qx, qy, qz);
        const float* r = recs + 9 * i;
        if (method == PCR_METHOD_ICP) accum_icp(acc, P, px, py, pz, qx - r[0], qy - r[1], qz - r[2]);
        else if (method == PCR_METHOD_NDT) accum_ndt(acc, P, px, py, pz, qx - r[0], qy - r[1], qz - r[2], r + 3);
        else accum_plane(acc, P, px, py, pz, qx - r[0], qy - r[1], qz - r[2], r[3], r[4], r[5]);
    }
    if (method == PCR_METHOD_ICP) assemble_icp(acc, T, out29);
    else for (int i = 0; i < PCR_NEQ; ++i) out29[i] = acc[i];
}

int hs_gn_step(const double* rec, double tol, double* T, double* dx, double* dxn) { return gauss_newton_step(rec, tol, T, dx, dxn); }
void hs_so3_exp(const double* w, double* R) { so3_exp(w, R); }
int hs_solve6(const double* H, const double* g, double* x) { return solve6(H, g, x); }
void hs_eig3(const double* c6, int64_t n, double* v) {
    for (int64_t i = 0; i < n; ++i)
        smallest_eigvec_sym3(c6[6 * i], c6[6 * i + 1], c6[6 * i + 2], c6[6 * i + 3], c6[6 * i + 4], c6[6 * i + 5],
                             v[3 * i], v[3 * i + 1], v[3 * i + 2], nullptr);
}
void hs_icov(const double* cov, int64_t n, double* icov) {
    for (int64_t i = 0; i < n; ++i) icov_closed_form(cov + 9 * i, icov + 9 * i);
}

// ---- tile-stream search (pcr_tile.cuh): host replay of one warp row at a time -------------------
// Same per-lane functions as the kernel; the warp glue of tile_search_row (leader election, staged
// cell box, table pass, staging in pieces of `cap` points, settle test, next radius) is restated with
// plain loops over 32 lane states.
struct HostTile {
    TileGrid v{};
    std::vector<uint32_t> cs;
    std::vector<float4> pairs;
    std::vector<uint32_t> perm;
};

void* hs_tile_build(const float* xyz, int64_t n, double c) {
    HostTile* t = new HostTile();
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    double maxabs = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], xyz[3 * i + a]); hi[a] = std::max(hi[a], xyz[3 * i + a]);
            maxabs = std::max(maxabs, (double)fabsf(xyz[3 * i + a]));
        }
    TileGrid& V = t->v;
    V.c = (float)c; V.inv_c = (float)(1.0 / c); V.inv_c2 = V.inv_c * V.inv_c;
    V.ox = lo[0] - 1.5f * V.c; V.oy = lo[1] - 1.5f * V.c; V.oz = lo[2] - 1.5f * V.c;
    double d[3];
    for (int a = 0; a < 3; ++a) d[a] = floor(((double)hi[a] - (double)(lo[a] - 1.5f * V.c)) / c) + 3.0;
    V.nx = (int)d[0]; V.ny = (int)d[1]; V.nz = (int)d[2];
    V.slack = 1e-3f + 1e-6f * (float)(2.0 * maxabs / c + (double)std::max(V.nx, std::max(V.ny, V.nz)));
    V.n = (uint32_t)n;
    const size_t ncells = (size_t)V.nx * V.ny * V.nz;
    std::vector<uint32_t> key(n);
    t->cs.assign(ncells + 1, 0u);
    for (int64_t i = 0; i < n; ++i) {
        const float gx = (xyz[3 * i] - V.ox) * V.inv_c, gy = (xyz[3 * i + 1] - V.oy) * V.inv_c, gz = (xyz[3 * i + 2] - V.oz) * V.inv_c;
        const int cx = cell_of(gx, V.nx), cy = cell_of(gy, V.ny), cz = cell_of(gz, V.nz);
        key[i] = (uint32_t)(((size_t)cz * V.ny + cy) * V.nx + cx);
        t->cs[key[i] + 1]++;
    }
    for (size_t k = 0; k < ncells; ++k) t->cs[k + 1] += t->cs[k];
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    const int64_t n_even = (n + 1) / 2 * 2;
    t->pairs.resize(n_even); t->perm.resize(n);
    float* rec = reinterpret_cast<float*>(t->pairs.data());
    for (int64_t i = 0; i < n_even; ++i) {
        float x = 3.0e38f, y = 3.0e38f, z = 3.0e38f, w;
        uint32_t pi = 0xffffffffu;
        if (i < n) { const uint32_t j = order[i]; x = xyz[3 * j]; y = xyz[3 * j + 1]; z = xyz[3 * j + 2]; pi = (uint32_t)i; t->perm[i] = j; }
        memcpy(&w, &pi, 4);
        float* r = rec + (i >> 1) * 8 + (i & 1);
        r[0] = x; r[2] = y; r[4] = z; r[6] = w;
    }
    V.cs = t->cs.data(); V.pairs = t->pairs.data();
    return t;
}
void hs_tile_free(void* t) { delete (HostTile*)t; }

// q: (m,3) queries already posed, processed in rows of 32 in the given order.  idx: caller index of
// the match or -1; dist: its distance; stats[0] = passes, [1] = staged pieces, [2] = rows scanned in
// global memory, [3] = staged points, [5] = candidates evaluated
// ng: groups per warp row (the kernel's NG): groups are independent, so a group of L = 32 / ng lanes
// with its slice cap / ng of the stage buffer is replayed as a row of its own.
void hs_tile_nn(void* tp, const float* qs, int64_t m, double max_dist, int cap, int core_e, double hint, int ng,
                int64_t* idx, float* dist, int64_t* stats) {
    HostTile* t = (HostTile*)tp;
    const TileGrid& G = t->v;
    const float md = (float)max_dist, max_d2 = md * md;
    const float rmax = tile_rmax(G, max_d2);
    g_tile_evals = 0;
    std::vector<float4> spts(cap);
    const uint32_t capp = ((uint32_t)cap >> 1) / (uint32_t)ng;
    const int L = 32 / ng;
    for (int64_t row0 = 0; row0 < m; row0 += L) {
        TileQuery q[32]; TileBest b[32]; bool valid[32]; float rl[32];
        uint32_t todo = 0;
        for (int l = 0; l < L; ++l) {
            const int64_t i = row0 + l;
            valid[l] = false;
            b[l].d2 = max_d2; b[l].pos = kTileNone; b[l].x = b[l].y = b[l].z = 0.f;
            q[l] = TileQuery{};
            q[l].qx = q[l].qy = q[l].qz = NAN;
            if (i < m && qs[3 * i] == qs[3 * i])
                valid[l] = tile_make_query(G, qs[3 * i], qs[3 * i + 1], qs[3 * i + 2], q[l]) && !tile_query_far_outside(G, q[l], max_d2);
            rl[l] = fminf(fmaxf((float)hint, 0.03125f), rmax);
            if (valid[l]) todo |= 1u << l;
        }
        while (todo) {
            stats[0]++;
            const int leader = __builtin_ffs((int)todo) - 1;
            bool elig[32];
            float rho = 0.f;
            for (int l = 0; l < L; ++l) {
                elig[l] = ((todo >> l) & 1u) && abs(q[l].ix - q[leader].ix) <= core_e && abs(q[l].iy - q[leader].iy) <= core_e &&
                          abs(q[l].iz - q[leader].iz) <= core_e;
                if (elig[l]) rho = fmaxf(rho, rl[l]);
            }
            TileBox U{INT_MAX, INT_MIN, INT_MAX, INT_MIN, INT_MAX, INT_MIN};
            for (int l = 0; l < L; ++l) {
                if (!elig[l]) continue;
                U.x0 = std::min(U.x0, tile_cell_floor(q[l].gx - rho)); U.x1 = std::max(U.x1, tile_cell_floor(q[l].gx + rho));
                U.y0 = std::min(U.y0, tile_cell_floor(q[l].gy - rho)); U.y1 = std::max(U.y1, tile_cell_floor(q[l].gy + rho));
                U.z0 = std::min(U.z0, tile_cell_floor(q[l].gz - rho)); U.z1 = std::max(U.z1, tile_cell_floor(q[l].gz + rho));
            }
            TileBox R{std::max(U.x0, 0), std::min(U.x1, G.nx - 1), std::max(U.y0, 0), std::min(U.y1, G.ny - 1), std::max(U.z0, 0), std::min(U.z1, G.nz - 1)};
            if (R.x0 <= R.x1 && R.y0 <= R.y1 && R.z0 <= R.z1) {
                const int rnx = R.x1 - R.x0 + 1, rny = R.y1 - R.y0 + 1, rnz = R.z1 - R.z0 + 1, nrows = rny * rnz;
                int ra = 0;
                while (ra < nrows) {
                    uint32_t ps[32], lenp[32], incl[32];
                    uint32_t run = 0; int nfit = 0;
                    for (int l = 0; l < L; ++l) {
                        const int r = ra + l;
                        ps[l] = 0; lenp[l] = 0;
                        if (r < nrows) {
                            const int jz = R.z0 + r / rny, jy = R.y0 + r % rny;
                            const size_t base = ((size_t)jz * G.ny + jy) * G.nx + R.x0;
                            const uint32_t gs = G.cs[base], ge = G.cs[base + rnx];
                            if (ge > gs) { ps[l] = gs >> 1; lenp[l] = ((ge + 1u) >> 1) - ps[l]; }
                        }
                        run += lenp[l]; incl[l] = run;
                        if (r < nrows && incl[l] <= capp) nfit++;
                    }
                    if (nfit == 0) {
                        stats[2]++;
                        const float4* gp = G.pairs + 2 * (size_t)ps[0];
                        for (int l = 0; l < L; ++l) {            // every lane scans (same as the kernel): settled lanes cannot improve
                            float best = b[l].d2;
                            const uint32_t slot = tile_scan_pairs(gp, lenp[0], q[l], best);
                            if (slot != kTileNone) tile_take(gp, slot, lenp[0], q[l], best, b[l]);
                        }
                        ra += 1;
                        continue;
                    }
                    const uint32_t total = incl[nfit - 1];
                    if (total > 0) {
                        stats[1]++; stats[3] += 2 * total;
                        for (int l = 0; l < nfit; ++l)
                            for (uint32_t k = 0; k < 2 * lenp[l]; ++k) spts[2 * (incl[l] - lenp[l]) + k] = G.pairs[2 * (size_t)ps[l] + k];
                        for (int l = 0; l < L; ++l) {
                            float best = b[l].d2;
                            const uint32_t slot = tile_scan_pairs(spts.data(), total, q[l], best);
                            if (slot != kTileNone) tile_take(spts.data(), slot, total, q[l], best, b[l]);
                        }
                    }
                    ra += nfit;
                }
            }
            for (int l = 0; l < L; ++l) {
                if (!elig[l]) continue;
                const bool settled = rho >= rmax || tile_settled(G, q[l], b[l], U);
                if (settled) todo &= ~(1u << l);
                else rl[l] = fminf(b[l].pos != kTileNone ? tile_radius_for(G, b[l]) : 2.0f * rho, rmax);
            }
        }
        for (int l = 0; l < L; ++l) {
            const int64_t i = row0 + l;
            if (i >= m) break;
            if (b[l].pos != kTileNone) { idx[i] = t->perm[b[l].pos]; dist[i] = sqrtf(b[l].d2); }
            else { idx[i] = -1; dist[i] = INFINITY; }
        }
    }
    stats[5] = g_tile_evals;
}

uint64_t hs_box_mask(int x0, int x1, int y0, int y1, int z0, int z1) { return brick_box_mask(x0, x1, y0, y1, z0, z1); }

}  // extern "C"
