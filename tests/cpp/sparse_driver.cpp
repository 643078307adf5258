// Test driver for csrc/rmsb_sparse.h (host-only symbolic analysis of the large-mechanism path): reads "n m" and m (row col)
// pairs from stdin, prints the analysed pattern as plain integers.  Built and run by tests/test_sparse_host.py.
#include <cstdio>
#include "../../reactionmechanismsimulator.jl_b200/csrc/rmsb_sparse.h"

static void dump(const char* name, const std::vector<int>& v) {
    printf("%s %zu", name, v.size());
    for (int x : v) printf(" %d", x);
    printf("\n");
}

int main() {
    int n, m;
    if (scanf("%d %d", &n, &m) != 2) return 1;
    std::vector<int> r(m), c(m);
    for (int i = 0; i < m; ++i) if (scanf("%d %d", &r[i], &c[i]) != 2) return 1;
    rmsb::SparsePattern sp = rmsb::analyse_pattern(n, r, c);
    printf("n %d core %d max_row %d flops %lld\n", sp.n, sp.core_start, sp.max_row, (long long)sp.factor_flops);
    dump("perm", sp.perm); dump("iperm", sp.iperm); dump("row_ptr", sp.row_ptr); dump("col", sp.col); dump("diag", sp.diag);
    dump("lvl_l_ptr", sp.lvl_l_ptr); dump("lvl_l_rows", sp.lvl_l_rows); dump("lvl_u_ptr", sp.lvl_u_ptr); dump("lvl_u_rows", sp.lvl_u_rows);
    return 0;
}
