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
    // border: nb state variables after the species block, each a dense row and column of I - c*J
    //   ConstantTP: nb = 1 (V);  ConstantV: nb = 2 (T, P);  ConstantP: nb = 2 (T, V)   (has_v: the volume is a state, border bv)
    int has_v, ns, nb, bv;
    const int* bcol_pos;    // [nb][n]  position of entry (i, ns + b) for every row i (border-border entries included)
    const int* brow_pos;    // [nb][ns] position of entry (ns + b, c) for the species columns c
    const int* je_colptr;   // [Ns+1] Jacobian entries of column c are je_colptr[c] .. je_colptr[c+1] (the entry list is sorted by column)
    // third-body efficiency entries d(rate)/d n_s through Ceff (src/Domain.jl:2151-2167): distinct (row, col) species pairs
    int n_ee;
    const int* ee_pos;      // [n_ee] position in the factor's value array
    const int* ee_row;      // [n_ee]
    const int* ee_col;      // [n_ee]
    const int* ee_ptr;      // [n_effrxn+1] start of reaction q's block in ee_idx
    const int* ee_idx;      // per efficiency reaction: [n_listed_species][8] entry id of (species of slot k, listed species), -1 = empty slot
    const int* ee_colptr;   // [Ns+1] entries grouped by column ...
    const int* ee_bycol;    // [n_ee] ... entry ids in that order
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
