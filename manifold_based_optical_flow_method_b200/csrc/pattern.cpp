// Sparsity pattern of the P1 vector system over a triangle mesh (host side, once per
// mesh).  The reference never builds a pattern explicitly: it lets lil_matrix grow one
// from element updates at rows/cols  vertex + N*alpha
// (utils/compute_optical_flow.py:78-93 for a2, :127-141 for a1).  The union pattern is
// the vertex adjacency (plus diagonal) with a dense 2x2 block per pair, n_blocks = N+2E.
//
// Also here: the Cuthill-McKee renumbering that keeps the SpMV gather window small, and
// the per-block lists of contributing (face, local pair) entries in ascending face order
// -- the order in which the reference accumulates (:60, :113) -- so that the device
// assembly is a deterministic gather instead of an atomic scatter.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <queue>
#include <vector>

#include "mof_b200.h"
#include "mof_error.h"

struct mof_pattern {
    int64_t N = 0, F = 0;
    std::vector<int32_t> perm, rowptr, col, diag, cptr, centry, tri;
    std::vector<int32_t> color_tile_ptr;     // block-multicolor ordering only: tiles of colour c
    std::vector<int32_t> level_ptr;          // level-scheduled ordering only: rows of dependency level l
    int64_t max_row = 0, bandwidth = 0;
};

namespace {

// BFS from `start` over unvisited vertices of the (original-id) adjacency; appends the
// visit order; neighbours are visited by increasing degree (Cuthill-McKee).
void cm_bfs(int32_t start, const std::vector<int64_t>& aptr, const std::vector<int32_t>& adj,
            std::vector<uint8_t>& seen, std::vector<int32_t>& order, std::vector<int32_t>& scratch) {
    size_t head = order.size();
    order.push_back(start);
    seen[start] = 1;
    while (head < order.size()) {
        int32_t u = order[head++];
        scratch.clear();
        for (int64_t q = aptr[u]; q < aptr[u + 1]; ++q) {
            int32_t w = adj[q];
            if (!seen[w]) { seen[w] = 1; scratch.push_back(w); }
        }
        std::sort(scratch.begin(), scratch.end(), [&](int32_t a, int32_t b) {
            int64_t da = aptr[a + 1] - aptr[a], db = aptr[b + 1] - aptr[b];
            return da != db ? da < db : a < b;
        });
        for (int32_t w : scratch) order.push_back(w);
    }
}

// Last vertex reached by a plain BFS from `start` inside its component (a cheap
// pseudo-peripheral vertex finder, two sweeps are enough for mesh graphs).
int32_t bfs_far(int32_t start, const std::vector<int64_t>& aptr, const std::vector<int32_t>& adj,
                std::vector<int32_t>& mark, int32_t stamp) {
    std::vector<int32_t> q;
    q.push_back(start);
    mark[start] = stamp;
    size_t head = 0;
    while (head < q.size()) {
        int32_t u = q[head++];
        for (int64_t e = aptr[u]; e < aptr[u + 1]; ++e) {
            int32_t w = adj[e];
            if (mark[w] != stamp) { mark[w] = stamp; q.push_back(w); }
        }
    }
    return q.back();
}

// Recursive coordinate bisection of the vertex set into patches of exactly `tile` vertices
// (one remainder patch at most): split along the longest box axis at a multiple of `tile`.
void rcb(std::vector<int32_t>& ids, size_t lo, size_t hi, const double* xyz, size_t tile,
         std::vector<std::pair<size_t, size_t>>& leaves) {
    const size_t n = hi - lo;
    if (n <= tile) { leaves.emplace_back(lo, hi); return; }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (size_t q = lo; q < hi; ++q)
        for (int a = 0; a < 3; ++a) {
            double c = xyz[3 * size_t(ids[q]) + a];
            mn[a] = std::min(mn[a], c);
            mx[a] = std::max(mx[a], c);
        }
    int ax = 0;
    for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
    size_t half = ((n / 2 + tile - 1) / tile) * tile;
    if (half >= n) half = n - (n % tile ? n % tile : tile);
    std::nth_element(ids.begin() + lo, ids.begin() + lo + half, ids.begin() + hi, [&](int32_t a, int32_t b) {
        double ca = xyz[3 * size_t(a) + ax], cb = xyz[3 * size_t(b) + ax];
        return ca != cb ? ca < cb : a < b;
    });
    rcb(ids, lo, lo + half, xyz, tile, leaves);
    rcb(ids, lo + half, hi, xyz, tile, leaves);
}

// Block-multicolour ordering for the SSOR sweeps: RCB patches of MOF_TILE_ROWS vertices,
// greedy colouring of the patch graph, numbering colour-major; inside a patch vertices keep
// their Cuthill-McKee rank order.  cm_rank[v] = position of reference vertex v in the CM order.
void block_multicolor(int64_t N, const double* xyz, const std::vector<int64_t>& aptr, const std::vector<int32_t>& adj,
                      const std::vector<int32_t>& cm_rank, std::vector<int32_t>& perm,
                      std::vector<int32_t>& color_tile_ptr) {
    const size_t tile = MOF_TILE_ROWS;
    std::vector<int32_t> ids(N);
    for (int64_t v = 0; v < N; ++v) ids[v] = int32_t(v);
    std::vector<std::pair<size_t, size_t>> leaves;
    rcb(ids, 0, size_t(N), xyz, tile, leaves);
    const int32_t np = int32_t(leaves.size());
    std::vector<int32_t> pid(N);
    for (int32_t k = 0; k < np; ++k)
        for (size_t q = leaves[k].first; q < leaves[k].second; ++q) pid[ids[q]] = k;
    // balanced greedy colouring of the patch adjacency graph in RCB (space-filling) order: among
    // the colours no neighbour uses, take the one with the fewest patches so far (equal-sized
    // colour classes keep every sweep launch large); open a new colour only when forced to
    std::vector<int32_t> color(np, -1), class_size;
    std::vector<uint8_t> used;
    int32_t ncol = 0;
    for (int32_t k = 0; k < np; ++k) {
        used.assign(size_t(ncol) + 1, 0);
        for (size_t q = leaves[k].first; q < leaves[k].second; ++q) {
            int32_t v = ids[q];
            for (int64_t e = aptr[v]; e < aptr[v + 1]; ++e) {
                int32_t c = color[pid[adj[e]]];
                if (c >= 0 && pid[adj[e]] != k) used[c] = 1;
            }
        }
        int32_t best = -1;
        for (int32_t c = 0; c < ncol; ++c)
            if (!used[c] && (best < 0 || class_size[c] < class_size[best])) best = c;
        if (best < 0) { best = ncol++; class_size.push_back(0); }
        color[k] = best;
        class_size[best]++;
    }
    // the (single) short patch must be the very last tile: make its colour class the last one
    int32_t short_patch = -1;
    for (int32_t k = 0; k < np; ++k)
        if (leaves[k].second - leaves[k].first != tile) short_patch = k;
    std::vector<int32_t> class_order(ncol);
    for (int32_t c = 0; c < ncol; ++c) class_order[c] = c;
    if (short_patch >= 0) std::swap(class_order[color[short_patch]], class_order[ncol - 1]);
    // class_order[c] = position of colour c; invert to iterate positions
    std::vector<int32_t> at(ncol);
    for (int32_t c = 0; c < ncol; ++c) at[class_order[c]] = c;
    perm.clear();
    perm.reserve(N);
    color_tile_ptr.assign(1, 0);
    int32_t tiles = 0;
    for (int32_t pos = 0; pos < ncol; ++pos) {
        const int32_t c = at[pos];
        for (int pass = 0; pass < 2; ++pass)          // the short patch goes last inside its class
            for (int32_t k = 0; k < np; ++k) {
                if (color[k] != c || ((k == short_patch) != (pass == 1))) continue;
                size_t b = perm.size();
                for (size_t q = leaves[k].first; q < leaves[k].second; ++q) perm.push_back(ids[q]);
                std::sort(perm.begin() + b, perm.end(), [&](int32_t a, int32_t d) { return cm_rank[a] < cm_rank[d]; });
                ++tiles;
            }
        color_tile_ptr.push_back(tiles);
    }
}

// Level-scheduled natural ordering for the SSOR sweeps: level(v) = 1 + max level of the neighbours
// that precede v in the Cuthill-McKee order; numbering level-major, Cuthill-McKee rank inside a
// level.  Every edge keeps its orientation (an earlier neighbour sits in a strictly lower level),
// so the Gauss-Seidel splitting -- and with it the SSOR preconditioner -- is exactly the one of the
// Cuthill-McKee order, while the rows of one level are mutually independent and contiguous.
void level_schedule(int64_t N, const std::vector<int64_t>& aptr, const std::vector<int32_t>& adj,
                    const std::vector<int32_t>& cm, std::vector<int32_t>& perm, std::vector<int32_t>& level_ptr) {
    std::vector<int32_t> rank(N), lev(N, 0);
    for (int64_t q = 0; q < N; ++q) rank[cm[q]] = int32_t(q);
    int32_t nlev = 0;
    for (int64_t q = 0; q < N; ++q) {
        const int32_t v = cm[q];
        int32_t l = 0;
        for (int64_t e = aptr[v]; e < aptr[v + 1]; ++e) {
            const int32_t r = rank[adj[e]];
            if (r < q) l = std::max(l, lev[r] + 1);
        }
        lev[q] = l;
        nlev = std::max(nlev, l + 1);
    }
    level_ptr.assign(size_t(nlev) + 1, 0);
    for (int64_t q = 0; q < N; ++q) level_ptr[lev[q] + 1]++;
    for (int32_t l = 0; l < nlev; ++l) level_ptr[l + 1] += level_ptr[l];
    std::vector<int32_t> fill(level_ptr.begin(), level_ptr.end() - 1);
    perm.assign(N, 0);
    for (int64_t q = 0; q < N; ++q) perm[fill[lev[q]]++] = cm[q];
}

}  // namespace

extern "C" int mof_pattern_create(int64_t N, int64_t F, const int64_t* triangles, int reorder,
                                  const double* coords, mof_pattern** out) {
    if (!out) return mof_set_error(-1, "mof_pattern_create: out is NULL");
    *out = nullptr;
    if (N <= 0 || F < 0 || !triangles) return mof_set_error(-1, "mof_pattern_create: bad sizes");
    if (N >= (int64_t(1) << 30) || F >= (int64_t(1) << 27))
        return mof_set_error(-1, "mof_pattern_create: mesh too large for int32 indices");
    for (int64_t f = 0; f < F; ++f) {
        int64_t a = triangles[3 * f], b = triangles[3 * f + 1], c = triangles[3 * f + 2];
        if (a < 0 || b < 0 || c < 0 || a >= N || b >= N || c >= N)
            return mof_set_error(-2, "mof_pattern_create: face %lld has a vertex id out of range",
                                 (long long)f);
        if (a == b || b == c || a == c)
            return mof_set_error(-2, "mof_pattern_create: face %lld repeats a vertex", (long long)f);
    }
    mof_pattern* P = new (std::nothrow) mof_pattern;
    if (!P) return mof_set_error(-3, "mof_pattern_create: out of memory");
    try {
        P->N = N;
        P->F = F;
        // --- vertex adjacency in reference ids
        std::vector<uint64_t> edges;
        edges.reserve(size_t(F) * 6);
        for (int64_t f = 0; f < F; ++f) {
            const int64_t* t = triangles + 3 * f;
            for (int m = 0; m < 3; ++m) {
                uint64_t u = uint64_t(t[m]), w = uint64_t(t[(m + 1) % 3]);
                edges.push_back(u << 32 | w);
                edges.push_back(w << 32 | u);
            }
        }
        std::sort(edges.begin(), edges.end());
        edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
        std::vector<int64_t> aptr(N + 1, 0);
        std::vector<int32_t> adj(edges.size());
        for (size_t q = 0; q < edges.size(); ++q) {
            aptr[(edges[q] >> 32) + 1]++;
            adj[q] = int32_t(edges[q] & 0xffffffffu);
        }
        for (int64_t v = 0; v < N; ++v) aptr[v + 1] += aptr[v];
        std::vector<uint64_t>().swap(edges);

        // --- renumbering
        if (reorder == 2 && !coords) {
            delete P;
            return mof_set_error(-1, "mof_pattern_create: reorder = 2 (block multicolour) needs coordinates");
        }
        P->perm.resize(N);
        if (reorder) {
            std::vector<uint8_t> seen(N, 0);
            std::vector<int32_t> mark(N, -1), order, scratch;
            order.reserve(N);
            int32_t stamp = 0;
            for (int64_t s = 0; s < N; ++s) {
                if (seen[s]) continue;
                int32_t a = bfs_far(int32_t(s), aptr, adj, mark, stamp++);
                int32_t b = bfs_far(a, aptr, adj, mark, stamp++);
                cm_bfs(b, aptr, adj, seen, order, scratch);
            }
            P->perm.swap(order);
            if (reorder == 2) {
                std::vector<int32_t> cm_rank(N);
                for (int64_t v = 0; v < N; ++v) cm_rank[P->perm[v]] = int32_t(v);
                std::vector<int32_t> bm;
                block_multicolor(N, coords, aptr, adj, cm_rank, bm, P->color_tile_ptr);
                P->perm.swap(bm);
            } else if (reorder == 3) {
                std::vector<int32_t> lv;
                level_schedule(N, aptr, adj, P->perm, lv, P->level_ptr);
                P->perm.swap(lv);
            }
        } else {
            for (int64_t v = 0; v < N; ++v) P->perm[v] = int32_t(v);
        }
        std::vector<int32_t> iperm(N);
        for (int64_t v = 0; v < N; ++v) iperm[P->perm[v]] = int32_t(v);

        // --- block rows (internal ids), ascending columns
        P->rowptr.assign(N + 1, 0);
        for (int64_t v = 0; v < N; ++v) {
            int32_t o = P->perm[v];
            P->rowptr[v + 1] = P->rowptr[v] + int32_t(aptr[o + 1] - aptr[o]) + 1;
        }
        int64_t nb = P->rowptr[N];
        if (nb >= (int64_t(1) << 31)) throw std::bad_alloc();
        P->col.resize(nb);
        P->diag.resize(N);
        for (int64_t v = 0; v < N; ++v) {
            int32_t o = P->perm[v];
            int32_t* c = P->col.data() + P->rowptr[v];
            int64_t k = 0;
            c[k++] = int32_t(v);
            for (int64_t q = aptr[o]; q < aptr[o + 1]; ++q) c[k++] = iperm[adj[q]];
            std::sort(c, c + k);
            P->diag[v] = P->rowptr[v] + int32_t(std::lower_bound(c, c + k, int32_t(v)) - c);
            P->max_row = std::max<int64_t>(P->max_row, k);
            P->bandwidth = std::max<int64_t>(P->bandwidth, std::max<int64_t>(v - c[0], c[k - 1] - v));
        }

        // --- triangles in internal ids, contributor lists
        P->tri.resize(size_t(F) * 3);
        for (int64_t q = 0; q < 3 * F; ++q) P->tri[q] = iperm[triangles[q]];
        P->cptr.assign(nb + 1, 0);
        auto block_of = [&](int32_t r, int32_t c) -> int32_t {
            const int32_t* b = P->col.data() + P->rowptr[r];
            const int32_t* e = P->col.data() + P->rowptr[r + 1];
            return int32_t(std::lower_bound(b, e, c) - P->col.data());
        };
        for (int64_t f = 0; f < F; ++f)
            for (int m = 0; m < 3; ++m)
                for (int n = 0; n < 3; ++n) P->cptr[block_of(P->tri[3 * f + m], P->tri[3 * f + n]) + 1]++;
        for (int64_t b = 0; b < nb; ++b) P->cptr[b + 1] += P->cptr[b];
        P->centry.resize(size_t(F) * 9);
        std::vector<int32_t> fill(P->cptr.begin(), P->cptr.end() - 1);
        for (int64_t f = 0; f < F; ++f)
            for (int m = 0; m < 3; ++m)
                for (int n = 0; n < 3; ++n) {
                    int32_t b = block_of(P->tri[3 * f + m], P->tri[3 * f + n]);
                    P->centry[fill[b]++] = int32_t(f << 4 | m << 2 | n);
                }
    } catch (const std::bad_alloc&) {
        delete P;
        return mof_set_error(-3, "mof_pattern_create: out of memory");
    }
    *out = P;
    return 0;
}

extern "C" void mof_pattern_destroy(mof_pattern* p) { delete p; }
extern "C" int64_t mof_pattern_num_blocks(const mof_pattern* p) { return p ? int64_t(p->col.size()) : -1; }
extern "C" int64_t mof_pattern_num_contrib(const mof_pattern* p) { return p ? int64_t(p->centry.size()) : -1; }
extern "C" int64_t mof_pattern_max_row_blocks(const mof_pattern* p) { return p ? p->max_row : -1; }
extern "C" int64_t mof_pattern_bandwidth(const mof_pattern* p) { return p ? p->bandwidth : -1; }

extern "C" int mof_pattern_colors(const mof_pattern* p, int32_t* n_colors, int32_t* color_tile_ptr) {
    if (!p || !n_colors) return mof_set_error(-1, "mof_pattern_colors: NULL argument");
    const int32_t nc = p->color_tile_ptr.empty() ? 0 : int32_t(p->color_tile_ptr.size()) - 1;
    if (nc > MOF_MAX_COLORS) return mof_set_error(-2, "mof_pattern_colors: %d colours exceed MOF_MAX_COLORS", nc);
    *n_colors = nc;
    if (color_tile_ptr)
        for (int32_t c = 0; c <= nc && nc > 0; ++c) color_tile_ptr[c] = p->color_tile_ptr[c];
    return 0;
}

extern "C" int32_t mof_pattern_levels(const mof_pattern* p, int32_t* level_ptr) {
    if (!p) return mof_set_error(-1, "mof_pattern_levels: NULL pattern");
    const int32_t nl = p->level_ptr.empty() ? 0 : int32_t(p->level_ptr.size()) - 1;
    if (level_ptr && nl > 0) std::memcpy(level_ptr, p->level_ptr.data(), p->level_ptr.size() * sizeof(int32_t));
    return nl;
}

extern "C" int mof_pattern_export(const mof_pattern* p, int32_t* perm, int32_t* rowptr, int32_t* col,
                                  int32_t* diag, int32_t* cptr, int32_t* centry, int32_t* tri) {
    if (!p) return mof_set_error(-1, "mof_pattern_export: NULL pattern");
    auto cp = [](int32_t* dst, const std::vector<int32_t>& src) {
        if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(int32_t));
    };
    cp(perm, p->perm);
    cp(rowptr, p->rowptr);
    cp(col, p->col);
    cp(diag, p->diag);
    cp(cptr, p->cptr);
    cp(centry, p->centry);
    cp(tri, p->tri);
    return 0;
}
