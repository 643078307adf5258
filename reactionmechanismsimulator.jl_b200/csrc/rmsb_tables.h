// rmsb_tables.h -- plain-data views of the per-mechanism device tables of the specialised paths (shared by the host-side
// builders and the kernels; no device code).
#pragma once
#include <stdint.h>

namespace rmsb {

// large-mechanism path (rmsb_large.cuh): fixed sparsity pattern of I - c*J analysed once per mechanism (rmsb_sparse.h)
struct LargeTables {
    int n, nnz, n_lvl_l, n_lvl_u, core;   // core: first row / column of the dense trailing block
    const int* perm;        // new -> old
    const int* row_ptr;
    const int* col;
    const int* diag;
    const int* ldiag;       // [nnz] for an entry in column k: position of the diagonal of row k
    const int* lnk;         // [nnz] ... and the number of U entries of row k right of its diagonal
    const int* lvl_l_ptr; const int* lvl_l_rows;
    const int* lvl_u_ptr; const int* lvl_u_rows;
    const int* amap;        // [n_je] Jacobian entry -> position in the factor's value array
    // ConstantTP (y = [ns; V], n = Ns + 1): the volume adds a dense last column and row to the species block
    int has_v, ns;
    const int* vcol_pos;    // [Ns]   position of entry (s, V)
    const int* vrow_pos;    // [Ns+1] position of entry (V, c), c = Ns is the diagonal
    const int* je_colptr;   // [Ns+1] Jacobian entries of column c are je_colptr[c] .. je_colptr[c+1] (the entry list is sorted by column)
};


// small-mechanism path (rmsb_small.cu): all-Arrhenius gas mechanisms with Ns <= 32, Nr <= 64.  One image of lane-major tables,
// staged once per block into shared memory; lane s of a reactor's lane group owns species s, lane j % G reaction j.
constexpr int SM_MAXNR = 64;   // reaction capacity = stride of the per-reaction tables
constexpr int SM_SP = 32;      // species capacity  = stride of the per-species tables
struct SmallTables {
    int ok;                    // 0: the mechanism does not qualify
    int Ns, Nr, NRP;           // NRP = Nr rounded up to a multiple of 32: layout of the per-group rate array [q | 0 | -q]
    int maxr, rows;            // NASA ranges per species (max), coefficients stored per range (7, or 9 if any species has 9)
    int max_nre, max_npr;      // most reactants / products of any reaction
    int max_kc;                // most distinct species in a row of the stoichiometric matrix
    int sr_words;              // 64-bit words (8 entries each) of the longest species -> reaction list
    int tab_doubles;           // size of the image
    int o_tmax, o_nasa, o_arr, o_rx, o_kc, o_kcc, o_sr;   // offsets (doubles) into the image
    const double* tab;         // the image in global memory
};

}  // namespace rmsb
