// rmsb_sparse.h -- host-side symbolic analysis for the large-mechanism path (config C3, >128 state variables).
//
// Role in the reference: for sparse mechanisms the reference hands CVODE a GMRES linear solver preconditioned by an
// incomplete LU of W = I - gamma*J (src/Reactor.jl:342-382, third-party IncompleteLU).  Here W is factorised
// exactly: the sparsity pattern of the analytic Jacobian is fixed by the mechanism, so the ordering, the fill
// pattern and the dependency levels are computed ONCE per mechanism on the host and the device only does numeric
// work on a fixed pattern (no pivot search; the diagonal of I - c*J is the pivot, a vanishing pivot is reported as
// a failed factorisation and the integrator halves the step).
//
//   1. pattern of A = I - c*J (J's species block from the mechanism's Jacobian entry list), symmetrised
//   2. minimum-degree ordering on that graph (smallest current degree, lowest index on ties)
//   3. the elimination gives the filled pattern: struct(k) = neighbours of pivot k at elimination time (all > k)
//      row i of L = {k < i : i in struct(k)},  row i of U = {i} + struct(i)          (symmetric superset of the true fill)
//   4. CSR rows [L cols | diag | U cols], ascending
//   5. dependency levels of the rows (forward by the L rows, backward by the U rows) and the start of the dense core
//      (the first pivot from which the remaining graph is a clique): the numeric factorisation is up-looking and
//      level-parallel over the sparse rows, then a dense right-looking LU on the core
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace rmsb {

struct SparsePattern {
    int n = 0;
    std::vector<int> perm;       // new index -> old index
    std::vector<int> iperm;      // old index -> new index
    std::vector<int> row_ptr;    // [n+1]
    std::vector<int> col;        // [nnz] new indices, ascending within a row
    std::vector<int> diag;       // [n] position of the diagonal entry
    std::vector<int> lvl_l_ptr, lvl_l_rows;   // forward-substitution levels
    std::vector<int> lvl_u_ptr, lvl_u_rows;   // backward-substitution levels
    int core_start = 0;          // first pivot of the dense core: from here on the remaining graph is a clique
    int max_row = 0;             // longest row
    int64_t factor_flops = 0;    // multiply-adds of one numeric factorisation
};

// entries: (row, col) pairs of the Jacobian block (old indices, 0-based, n x n); duplicates allowed
inline SparsePattern analyse_pattern(int n, const std::vector<int>& erow, const std::vector<int>& ecol) {
    SparsePattern sp;
    sp.n = n;
    std::vector<std::vector<int>> adj(n);
    for (size_t e = 0; e < erow.size(); ++e) {
        int r = erow[e], c = ecol[e];
        if (r == c) continue;
        adj[r].push_back(c);
        adj[c].push_back(r);
    }
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    // ---- minimum degree elimination with explicit cliques (sorted adjacency vectors)
    std::vector<char> done(n, 0);
    std::vector<std::vector<int>> structs(n);     // by elimination step: neighbours (old indices) at elimination time
    sp.perm.resize(n); sp.iperm.resize(n);
    std::vector<int> merged;
    bool core_found = false;
    for (int step = 0; step < n; ++step) {
        int best = -1; size_t bd = SIZE_MAX;
        for (int v = 0; v < n; ++v) if (!done[v] && adj[v].size() < bd) { bd = adj[v].size(); best = v; }
        const int v = best;
        if (!core_found && (int)adj[v].size() == n - step - 1) { core_found = true; sp.core_start = step; }
        done[v] = 1;
        sp.perm[step] = v; sp.iperm[v] = step;
        structs[step] = adj[v];
        const std::vector<int>& nb = structs[step];
        for (int a : nb) {
            // adj[a] = (adj[a] U nb) \ {a, v}
            merged.clear();
            merged.reserve(adj[a].size() + nb.size());
            std::set_union(adj[a].begin(), adj[a].end(), nb.begin(), nb.end(), std::back_inserter(merged));
            merged.erase(std::remove_if(merged.begin(), merged.end(), [&](int x) { return x == a || x == v; }), merged.end());
            adj[a].swap(merged);
        }
        adj[v].clear(); adj[v].shrink_to_fit();
    }
    // ---- rows of the filled matrix in the new numbering
    std::vector<std::vector<int>> rows(n);
    for (int k = 0; k < n; ++k) {
        rows[k].push_back(k);
        for (int old : structs[k]) {
            int j = sp.iperm[old];          // j > k
            rows[k].push_back(j);           // U entry (k, j)
            rows[j].push_back(k);           // L entry (j, k)
        }
        sp.factor_flops += (int64_t)structs[k].size() * (int64_t)structs[k].size();
    }
    sp.row_ptr.assign(n + 1, 0);
    sp.diag.resize(n);
    for (int i = 0; i < n; ++i) {
        std::sort(rows[i].begin(), rows[i].end());
        sp.row_ptr[i + 1] = sp.row_ptr[i] + (int)rows[i].size();
        sp.max_row = std::max(sp.max_row, (int)rows[i].size());
    }
    sp.col.resize(sp.row_ptr[n]);
    for (int i = 0; i < n; ++i) {
        std::copy(rows[i].begin(), rows[i].end(), sp.col.begin() + sp.row_ptr[i]);
        sp.diag[i] = sp.row_ptr[i] + (int)(std::lower_bound(rows[i].begin(), rows[i].end(), i) - rows[i].begin());
    }
    // ---- levels
    auto bucket = [&](const std::vector<int>& level, std::vector<int>& ptr, std::vector<int>& lst) {
        int nl = 0;
        for (int v : level) nl = std::max(nl, v + 1);
        ptr.assign(nl + 1, 0);
        for (int v : level) ++ptr[v + 1];
        for (int l = 0; l < nl; ++l) ptr[l + 1] += ptr[l];
        lst.resize(n);
        std::vector<int> fill(ptr.begin(), ptr.end() - 1);
        for (int i = 0; i < n; ++i) lst[fill[level[i]]++] = i;
    };
    std::vector<int> lv(n, 0);
    for (int i = 0; i < n; ++i)
        for (int p = sp.row_ptr[i]; p < sp.diag[i]; ++p) lv[i] = std::max(lv[i], lv[sp.col[p]] + 1);
    bucket(lv, sp.lvl_l_ptr, sp.lvl_l_rows);
    std::vector<int> lu(n, 0);
    for (int i = n - 1; i >= 0; --i)
        for (int p = sp.diag[i] + 1; p < sp.row_ptr[i + 1]; ++p) lu[i] = std::max(lu[i], lu[sp.col[p]] + 1);
    bucket(lu, sp.lvl_u_ptr, sp.lvl_u_rows);
    return sp;
}

// position of entry (new row i, new col j) in the CSR arrays, -1 if absent
inline int find_entry(const SparsePattern& sp, int i, int j) {
    auto b = sp.col.begin() + sp.row_ptr[i], e = sp.col.begin() + sp.row_ptr[i + 1];
    auto it = std::lower_bound(b, e, j);
    return (it != e && *it == j) ? (int)(it - sp.col.begin()) : -1;
}

}  // namespace rmsb
