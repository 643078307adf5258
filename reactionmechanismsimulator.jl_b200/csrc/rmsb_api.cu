// rmsb_api.cu -- C ABI of librmsb200.so (see include/rmsb200.h): mechanism flattening into SoA device
// tables, kernel instantiations and the host/device-buffer entry points.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rmsb200.h"
#include "rmsb_device.cuh"
#include "rmsb_solver.cuh"
#include "rmsb_large.cuh"
#include "rmsb_sparse.h"

using namespace rmsb;

#include <thread>
#include "rmsb_internal.h"

thread_local std::string rmsb_g_err;
std::atomic<long long> rmsb_g_launches{0};
int rmsb_fail(int code, const std::string& msg) { rmsb_g_err = msg; return code; }

#define upload rmsb_upload

extern "C" int rmsb_version(void) { return RMSB_VERSION; }
extern "C" const char* rmsb_last_error(void) { return g_err.c_str(); }
extern "C" int rmsb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int64_t rmsb_kernel_launch_count(void) { return g_launches.load(); }

static int domain_ny(const rmsb_mech* m, int kind) {
    switch (kind) {
    case RMSB_DOMAIN_CONSTANT_V:
    case RMSB_DOMAIN_CONSTANT_P: return m->dm.Ns + 2;
    case RMSB_DOMAIN_CONSTANT_TP:
    case RMSB_DOMAIN_COUPLED_TP_TAPHI: return m->dm.Ns + 1;
    case RMSB_DOMAIN_CONSTANT_TV:
    case RMSB_DOMAIN_CONSTANT_TAPHI: return m->dm.Ns;
    }
    return -1;
}
static int domain_ncond(int kind) {
    switch (kind) {
    case RMSB_DOMAIN_CONSTANT_V:
    case RMSB_DOMAIN_CONSTANT_P: return 1;
    case RMSB_DOMAIN_CONSTANT_TP:
    case RMSB_DOMAIN_CONSTANT_TV:
    case RMSB_DOMAIN_CONSTANT_TAPHI: return 2;
    case RMSB_DOMAIN_COUPLED_TP_TAPHI: return 3;
    }
    return -1;
}
extern "C" int rmsb_mech_ny(const rmsb_mech* m, int kind) { return m ? domain_ny(m, kind) : -1; }
extern "C" int rmsb_mech_ncond(const rmsb_mech* m, int kind) { return m ? domain_ncond(kind) : -1; }
extern "C" int rmsb_mech_device(const rmsb_mech* m) { return m ? m->device : -1; }

extern "C" void rmsb_mech_destroy(rmsb_mech* h) {
    if (!h) return;
    DeviceGuard g(h->device);
    rmsb_spec_release(h);
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

static int build_mech(const rmsb_mech_desc* d, rmsb_mech* h, bool host_only = false) {
    const int Ns = d->n_species, Nr = d->n_reactions;
    DevMech& dm = h->dm;
    dm.Ns = Ns; dm.Nr = Nr;
    dm.phase_kind = d->phase_kind;
    dm.difflim = d->diffusion_limited ? 1 : 0;
    dm.sitedensity = d->sitedensity;
    dm.maxr = d->nasa_max_ranges;
    const int maxr = d->nasa_max_ranges;
    // ---- thermo, SoA over species
    std::vector<int> nr(Ns), nc((size_t)maxr * Ns, 7);
    std::vector<double> tmax((size_t)maxr * Ns, 0.0), coef((size_t)maxr * 9 * Ns, 0.0);
    for (int s = 0; s < Ns; ++s) {
        nr[s] = d->nasa_nranges[s];
        if (nr[s] < 1 || nr[s] > maxr) return fail(RMSB_ERR_ARG, "nasa_nranges out of range");
        for (int r = 0; r < nr[s]; ++r) {
            int ncf = d->nasa_ncoef[s * maxr + r];
            if (ncf != 7 && ncf != 9) return fail(RMSB_ERR_ARG, "NASA polynomial must have 7 or 9 coefficients");
            nc[(size_t)r * Ns + s] = ncf;
            tmax[(size_t)r * Ns + s] = d->nasa_tmax[s * maxr + r];
            for (int c = 0; c < ncf; ++c) coef[((size_t)r * 9 + c) * Ns + s] = d->nasa_coefs[((size_t)s * maxr + r) * 9 + c];
        }
    }
    // ---- reactions
    std::vector<int> rx((size_t)8 * Nr, -1), ktype(Nr), koff(Nr + 1), effptr(Nr + 1), effspc, effrxn;
    std::vector<double> effval, rev(Nr), fwd(Nr);
    for (int j = 0; j < Nr; ++j) {
        for (int k = 0; k < 8; ++k) {
            int v = d->rxnarray[j * 8 + k];
            if (v < 0 || v > Ns) return fail(RMSB_ERR_ARG, "rxnarray entry out of range");
            rx[(size_t)k * Nr + j] = v - 1;
        }
        if (rx[j] < 0 || rx[(size_t)4 * Nr + j] < 0) return fail(RMSB_ERR_ARG, "reaction without reactant or product");
        for (int k = 1; k < 4; ++k)
            if (rx[(size_t)k * Nr + j] >= 0 && rx[(size_t)(k - 1) * Nr + j] < 0) return fail(RMSB_ERR_ARG, "rxnarray slots must be packed");
        for (int k = 5; k < 8; ++k)
            if (rx[(size_t)k * Nr + j] >= 0 && rx[(size_t)(k - 1) * Nr + j] < 0) return fail(RMSB_ERR_ARG, "rxnarray slots must be packed");
        ktype[j] = d->kin_type[j];
        if (ktype[j] < 0 || ktype[j] > RMSB_KIN_CHEBYSHEV_VEC) return fail(RMSB_ERR_UNSUPPORTED, "unknown kinetics type");
        rev[j] = d->reversible[j] ? 1.0 : 0.0;
        fwd[j] = d->forwardable[j] ? 1.0 : 0.0;
    }
    if (maxr < 1) return fail(RMSB_ERR_ARG, "nasa_max_ranges must be >= 1");
    for (int j = 0; j <= Nr; ++j) { koff[j] = d->kin_offset[j]; effptr[j] = d->eff_ptr ? d->eff_ptr[j] : 0; }
    if (koff[0] < 0 || effptr[0] != 0) return fail(RMSB_ERR_ARG, "kin_offset[0] must be >= 0 and eff_ptr[0] must be 0");
    for (int j = 0; j < Nr; ++j) {
        if (koff[j + 1] < koff[j] || effptr[j + 1] < effptr[j]) return fail(RMSB_ERR_ARG, "kin_offset / eff_ptr must be non-decreasing");
        // parameter block sizes the kernels read for each kinetics type (rmsb200.h, RMSB_KIN_*)
        const int len = koff[j + 1] - koff[j];
        const double* a = d->kin_params + koff[j];
        int need = 3;
        switch (ktype[j]) {
        case RMSB_KIN_ARRHENIUS_VEC: case RMSB_KIN_ARRHENIUS: case RMSB_KIN_THIRDBODY: need = 3; break;
        case RMSB_KIN_LINDEMANN: need = 6; break;
        case RMSB_KIN_TROE: need = 10; break;
        case RMSB_KIN_MULTIARRHENIUS: need = len >= 1 ? 1 + 3 * (int)a[0] : 1; break;
        case RMSB_KIN_PLOG_VEC: need = len >= 1 ? 1 + 4 * (int)a[0] : 1; break;
        case RMSB_KIN_CHEBYSHEV_VEC: need = len >= 6 ? 6 + (int)a[4] * (int)a[5] : 6; break;
        default: need = 1; break;       // PLOG / MULTIPLOG: nested blocks, walked by their own counts
        }
        if (len < need) return fail(RMSB_ERR_ARG, "kin_params block shorter than its kinetics type needs");
    }
    std::vector<double> kpar(d->kin_params, d->kin_params + koff[Nr]);
    for (int k = 0; k < effptr[Nr]; ++k) {
        int s = d->eff_species[k];
        if (s < 1 || s > Ns) return fail(RMSB_ERR_ARG, "efficiency species out of range");
        effspc.push_back(s - 1);
        effval.push_back(d->eff_value[k]);
    }
    for (int j = 0; j < Nr; ++j) {
        bool falloff = ktype[j] == RMSB_KIN_THIRDBODY || ktype[j] == RMSB_KIN_LINDEMANN || ktype[j] == RMSB_KIN_TROE;
        if (effptr[j + 1] > effptr[j]) {
            if (!falloff) return fail(RMSB_ERR_ARG, "efficiencies on a non-falloff reaction");
            effrxn.push_back(j);
        }
    }
    // ---- stoichiometric matrix rows exactly as getstoichmatrix builds them (src/Phase.jl:203-222):
    //      slot index (1-based) > 3 is counted as a product.
    std::vector<int> kcptr(Nr + 1, 0), kcspc;
    std::vector<double> kccoef, nrp(Nr);
    for (int j = 0; j < Nr; ++j) {
        std::map<int, double> row;
        double n = 0.0;
        for (int k = 0; k < 8; ++k) {
            int s = rx[(size_t)k * Nr + j];
            if (s < 0) continue;
            if (k + 1 > 3) { row[s] -= 1.0; n += 1.0; } else { row[s] += 1.0; n -= 1.0; }
        }
        nrp[j] = n;
        for (auto& kv : row) { kcspc.push_back(kv.first); kccoef.push_back(kv.second); }
        kcptr[j + 1] = (int)kcspc.size();
    }
    // ---- species -> reaction slots, reaction order (src/Reactor.jl:484-505)
    std::vector<std::vector<int>> sr(Ns);
    for (int j = 0; j < Nr; ++j) {
        for (int k = 0; k < 4; ++k) { int s = rx[(size_t)k * Nr + j]; if (s >= 0) sr[s].push_back(j | (int)0x80000000); }
        for (int k = 4; k < 8; ++k) { int s = rx[(size_t)k * Nr + j]; if (s >= 0) sr[s].push_back(j); }
    }
    // species whose net stoichiometry is zero in every reaction: their dn/dt and Jacobian rows are exact zeros
    std::vector<int> inert(Ns, 1);
    for (int j = 0; j < Nr; ++j) {
        std::map<int, int> net;
        for (int k = 0; k < 8; ++k) { int s = rx[(size_t)k * Nr + j]; if (s >= 0) net[s] += (k < 4) ? -1 : 1; }
        for (auto& kv : net) if (kv.second != 0) inert[kv.first] = 0;
    }
    std::vector<int> srptr(Ns + 1, 0), srrxn;
    for (int s = 0; s < Ns; ++s) { for (int v : sr[s]) srrxn.push_back(v); srptr[s + 1] = (int)srrxn.size(); }
    // ---- Jacobian ns-block: derivative terms + per-entry gather lists in the accumulation order of
    //      _jacobianynswrtns! (src/Reactor.jl:899-1366): reaction ascending, forward then reverse.
    std::vector<int> dtrxn, dtmult, dtoth, dtspc;
    std::map<std::pair<int, int>, std::vector<int>> entries;  // (col,row) -> signed term refs
    for (int j = 0; j < Nr; ++j) {
        for (int dir = 0; dir < 2; ++dir) {
            std::vector<int> slots, oslots;
            for (int k = 0; k < 4; ++k) {
                int a = rx[(size_t)(dir ? k + 4 : k) * Nr + j], b = rx[(size_t)(dir ? k : k + 4) * Nr + j];
                if (a >= 0) slots.push_back(a);
                if (b >= 0) oslots.push_back(b);
            }
            std::vector<int> seen;
            for (int s : slots) {
                bool dup = false;
                for (int q : seen) if (q == s) dup = true;
                if (dup) continue;
                seen.push_back(s);
                int mcount = 0;
                for (int q : slots) if (q == s) ++mcount;
                int t = (int)dtrxn.size();
                dtrxn.push_back(j | (dir << 30));
                dtmult.push_back(mcount);
                dtspc.push_back(s);
                bool removed = false;
                int no = 0;
                int oth[3] = {-1, -1, -1};
                for (int q : slots) {
                    if (q == s && !removed) { removed = true; continue; }
                    oth[no++] = q;
                }
                for (int k = 0; k < 3; ++k) dtoth.push_back(oth[k]);
                for (int q : slots) entries[{s, q}].push_back(t | (int)0x80000000);
                for (int q : oslots) entries[{s, q}].push_back(t);
            }
        }
    }
    std::vector<int> jeptr{0}, jerow, jecol, jtidx;
    for (auto& kv : entries) {
        jecol.push_back(kv.first.first);
        jerow.push_back(kv.first.second);
        for (int v : kv.second) jtidx.push_back(v);
        jeptr.push_back((int)jtidx.size());
    }
    dm.heavy_sr = 0; dm.heavy_je = 0;
    for (int s2 = 0; s2 < Ns; ++s2) if (srptr[s2 + 1] - srptr[s2] > COOP_ROW) dm.heavy_sr = 1;
    for (size_t e = 0; e + 1 < jeptr.size(); ++e) if (jeptr[e + 1] - jeptr[e] > COOP_ROW) dm.heavy_je = 1;
    dm.n_dt = h->n_dt = (int)dtrxn.size();
    dm.n_je = h->n_je = (int)jerow.size();
    dm.n_effrxn = (int)effrxn.size();
    // ---- two-phase systems: phase of every species, domain of every reaction, C0 exponents per phase
    std::vector<int> rxdom(Nr, 0), srmid(Ns, 0);
    std::vector<double> nrp2(Nr, 0.0);
    dm.coupled = (d->species_phase && d->rxn_domain) ? 1 : 0;
    dm.n_bulk = Ns;
    if (dm.coupled) {
        int nb = 0;
        while (nb < Ns && d->species_phase[nb] == 0) ++nb;
        for (int s = nb; s < Ns; ++s) if (d->species_phase[s] != 1) return fail(RMSB_ERR_ARG, "species_phase must be 0...0 1...1");
        dm.n_bulk = nb;
        for (int j = 0; j < Nr; ++j) {
            rxdom[j] = d->rxn_domain[j];
            if (rxdom[j] < 0 || rxdom[j] > 2 || (j > 0 && rxdom[j] < rxdom[j - 1])) return fail(RMSB_ERR_ARG, "rxn_domain must be sorted 0..0 1..1 2..2");
            double nb_ = 0.0, ns_ = 0.0;
            for (int k = 0; k < 8; ++k) {
                int sp = rx[(size_t)k * Nr + j];
                if (sp < 0) continue;
                double sgn = (k < 4) ? -1.0 : 1.0;
                if (sp < nb) nb_ += sgn; else ns_ += sgn;
            }
            nrp[j] = nb_; nrp2[j] = ns_;
            if (effptr[j + 1] > effptr[j]) return fail(RMSB_ERR_UNSUPPORTED, "third-body efficiencies in a coupled gas-surface system");
        }
        for (int s = 0; s < Ns; ++s) {
            int mid = srptr[s + 1];
            for (int k = srptr[s]; k < srptr[s + 1]; ++k) if (rxdom[srrxn[k] & 0x7fffffff] == 2) { mid = k; break; }
            srmid[s] = mid;
        }
    }
    // ---- liquid phase data
    std::vector<double> radius, diffr;
    if (d->radius) radius.assign(d->radius, d->radius + Ns);
    if (d->diff_radius) diffr.assign(d->diff_radius, d->diff_radius + Ns);
    if (dm.difflim && (radius.empty() || diffr.empty()))
        return fail(RMSB_ERR_ARG, "diffusion-limited phase needs radius and diff_radius");
    dm.visc_kind = d->viscosity_kind;
    for (int k = 0; k < 5; ++k) dm.visc[k] = d->viscosity_params[k];

    int rc;
    if ((rc = rmsb_spec_capture(h, rx, srptr, srrxn, kcptr, kcspc, kccoef, nrp, ktype, koff, kpar, fwd, rev, nr, nc, tmax, coef, maxr,
                                dtrxn, dtmult, dtoth, jeptr, jerow, jecol, jtidx)) != RMSB_OK) return rc;
    if (host_only) return RMSB_OK;         // rmsb_spec_source_desc: tables only, nothing on a device
#define UP(vec, field) if ((rc = upload(h, vec, &dm.field)) != RMSB_OK) return rc
    UP(nr, nasa_nr); UP(nc, nasa_nc); UP(tmax, nasa_tmax); UP(coef, nasa_c);
    UP(rx, rx); UP(ktype, ktype); UP(koff, koff); UP(kpar, kpar);
    UP(effptr, eff_ptr); UP(effspc, eff_spc); UP(effval, eff_val); UP(effrxn, effrxn);
    UP(rev, rev); UP(fwd, fwd);
    UP(kcptr, kc_ptr); UP(kcspc, kc_spc); UP(kccoef, kc_coef); UP(nrp, nrp);
    UP(srptr, sr_ptr); UP(srrxn, sr_rxn);
    UP(dtrxn, dt_rxn); UP(dtmult, dt_mult); UP(dtoth, dt_oth); UP(dtspc, dt_spc);
    UP(jeptr, je_ptr); UP(jerow, je_row); UP(jecol, je_col); UP(jtidx, jt_idx);
    UP(rxdom, rx_dom); UP(nrp2, nrp2); UP(srmid, sr_mid); UP(inert, inert);
    if (Ns > 128 && !dm.coupled) {
        std::vector<int> colptr(Ns + 1, 0);                        // the entry list is sorted by (column, row)
        for (size_t e = 0; e < jecol.size(); ++e) ++colptr[jecol[e] + 1];
        for (int c = 0; c < Ns; ++c) colptr[c + 1] += colptr[c];
        // third-body efficiency entries (row = a species of the reaction, col = a species with a listed efficiency): distinct pairs,
        // per-reaction index blocks [listed species][8 slots] and the same entries grouped by column
        std::vector<int> eerow, eecol, eeptr(1, 0), eeidx;
        {
            std::map<std::pair<int, int>, int> ids;
            for (int j : effrxn) {
                for (int k = effptr[j]; k < effptr[j + 1]; ++k) {
                    const int col = effspc[k];
                    for (int q = 0; q < 8; ++q) {
                        const int sp = rx[(size_t)q * Nr + j];
                        if (sp < 0) { eeidx.push_back(-1); continue; }
                        auto it = ids.find({sp, col});
                        if (it == ids.end()) { it = ids.emplace(std::make_pair(sp, col), (int)eerow.size()).first; eerow.push_back(sp); eecol.push_back(col); }
                        eeidx.push_back(it->second);
                    }
                }
                eeptr.push_back((int)eeidx.size());
            }
        }
        const int n_ee = (int)eerow.size();
        std::vector<int> eecolptr(Ns + 1, 0), eebycol(n_ee);
        for (int e = 0; e < n_ee; ++e) ++eecolptr[eecol[e] + 1];
        for (int c = 0; c < Ns; ++c) eecolptr[c + 1] += eecolptr[c];
        {
            std::vector<int> fill(eecolptr.begin(), eecolptr.end() - 1);
            for (int e = 0; e < n_ee; ++e) eebycol[fill[eecol[e]]++] = e;
        }
        if (eeidx.empty()) eeidx.push_back(-1);
        for (int nb = 0; nb < 3; ++nb) {
            if (nb == 2 && dm.phase_kind != RMSB_PHASE_GAS) continue;   // ConstantV / ConstantP need an ideal gas
            const int n = Ns + nb;
            std::vector<int> er(jerow), ec(jecol);
            if (nb > 0) { er.insert(er.end(), eerow.begin(), eerow.end()); ec.insert(ec.end(), eecol.begin(), eecol.end()); }
            for (int b2 = 0; b2 < nb; ++b2)
                for (int i = 0; i < n; ++i) { er.push_back(i); ec.push_back(Ns + b2); er.push_back(Ns + b2); ec.push_back(i); }
            SparsePattern sp = analyse_pattern(n, er, ec);
            std::vector<int> ldiag(sp.col.size()), lnk(sp.col.size()), amap(jerow.size()), bcolpos, browpos, eepos;
            for (size_t p = 0; p < sp.col.size(); ++p) {
                const int k = sp.col[p];
                ldiag[p] = sp.diag[k];
                lnk[p] = sp.row_ptr[k + 1] - sp.diag[k] - 1;
            }
            for (size_t e = 0; e < jerow.size(); ++e) {
                amap[e] = find_entry(sp, sp.iperm[jerow[e]], sp.iperm[jecol[e]]);
                if (amap[e] < 0) return fail(RMSB_ERR_ARG, "internal: sparse pattern does not contain a Jacobian entry");
            }
            for (int b2 = 0; b2 < nb; ++b2) {
                for (int i = 0; i < n; ++i) bcolpos.push_back(find_entry(sp, sp.iperm[i], sp.iperm[Ns + b2]));
                for (int c = 0; c < Ns; ++c) browpos.push_back(find_entry(sp, sp.iperm[Ns + b2], sp.iperm[c]));
            }
            if (nb > 0) for (int e = 0; e < n_ee; ++e) eepos.push_back(find_entry(sp, sp.iperm[eerow[e]], sp.iperm[eecol[e]]));
            for (int v : bcolpos) if (v < 0) return fail(RMSB_ERR_ARG, "internal: sparse pattern does not contain a border entry");
            for (int v : eepos) if (v < 0) return fail(RMSB_ERR_ARG, "internal: sparse pattern does not contain an efficiency entry");
            LargeTables& lt = nb == 0 ? h->lt : (nb == 1 ? h->lt_tp : h->lt_vk);
            lt.n = n; lt.ns = Ns; lt.nb = nb; lt.has_v = (nb == 1) ? 1 : 0; lt.bv = 0; lt.nnz = (int)sp.col.size();
            lt.n_ee = nb > 0 ? n_ee : 0;
            lt.n_lvl_l = (int)sp.lvl_l_ptr.size() - 1; lt.n_lvl_u = (int)sp.lvl_u_ptr.size() - 1; lt.core = sp.core_start;
#define UPL(vec, field) if ((rc = upload(h, vec, &lt.field)) != RMSB_OK) return rc
            UPL(sp.perm, perm); UPL(sp.row_ptr, row_ptr); UPL(sp.col, col); UPL(sp.diag, diag); UPL(ldiag, ldiag); UPL(lnk, lnk);
            UPL(sp.lvl_l_ptr, lvl_l_ptr); UPL(sp.lvl_l_rows, lvl_l_rows); UPL(sp.lvl_u_ptr, lvl_u_ptr); UPL(sp.lvl_u_rows, lvl_u_rows);
            UPL(amap, amap); UPL(bcolpos, bcol_pos); UPL(browpos, brow_pos); UPL(colptr, je_colptr);
            UPL(eepos, ee_pos); UPL(eerow, ee_row); UPL(eecol, ee_col); UPL(eeptr, ee_ptr); UPL(eeidx, ee_idx);
            UPL(eecolptr, ee_colptr); UPL(eebycol, ee_bycol);
#undef UPL
            if (nb == 0) { h->large = true; h->large_max_row = sp.max_row; h->large_factor_flops = sp.factor_flops; }
        }
    }
#undef UP
    if ((rc = rmsb_small_build(h, rx, srptr, srrxn, kcptr, kcspc, kccoef, nrp, ktype, koff, kpar, fwd, rev, nr, nc, tmax, coef, maxr)) != RMSB_OK) return rc;
    if (!radius.empty()) { if ((rc = upload(h, radius, &dm.radius)) != RMSB_OK) return rc; } else dm.radius = nullptr;
    if (!diffr.empty()) { if ((rc = upload(h, diffr, &dm.diff_r)) != RMSB_OK) return rc; } else dm.diff_r = nullptr;
    return RMSB_OK;
}

extern "C" int rmsb_mech_create(const rmsb_mech_desc* desc, int device, rmsb_mech** out) {
    if (!desc || !out) return fail(RMSB_ERR_ARG, "null argument");
    *out = nullptr;
    if (desc->struct_size != (int32_t)sizeof(rmsb_mech_desc)) return fail(RMSB_ERR_ARG, "rmsb_mech_desc size mismatch (ABI)");
    if (desc->n_species < 1 || desc->n_reactions < 1) return fail(RMSB_ERR_ARG, "empty mechanism");
    int ndev = rmsb_device_count();
    if (ndev <= 0) return fail(RMSB_ERR_CUDA, "no CUDA device available (librmsb200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(RMSB_ERR_ARG, "device ordinal out of range");
    DeviceGuard g(device);
    if (!g.ok) return fail(RMSB_ERR_CUDA, "cudaSetDevice failed");
    rmsb_mech* h = new rmsb_mech();
    h->device = device;
    int rc = build_mech(desc, h);
    if (rc != RMSB_OK) { std::string keep = g_err; rmsb_mech_destroy(h); g_err = keep; return rc; }
    *out = h;
    return RMSB_OK;
}

// Generated CUDA source of the mechanism-specialised dydt / Jacobian kernels (rmsb_spec.cu), from the descriptor alone: needs no
// device.  Writes at most cap bytes (NUL-terminated) and the full length (without NUL) to *need.
extern "C" int rmsb_spec_source_desc(const rmsb_mech_desc* desc, int domain_kind, char* buf, size_t cap, size_t* need) {
    if (!desc) return fail(RMSB_ERR_ARG, "null argument");
    if (desc->struct_size != (int32_t)sizeof(rmsb_mech_desc)) return fail(RMSB_ERR_ARG, "rmsb_mech_desc size mismatch (ABI)");
    if (desc->n_species < 1 || desc->n_reactions < 1) return fail(RMSB_ERR_ARG, "empty mechanism");
    rmsb_mech tmp;
    int rc = build_mech(desc, &tmp, true);
    if (rc != RMSB_OK) return rc;
    std::string src;
    if ((rc = rmsb_spec_source(&tmp, domain_kind, &src)) != RMSB_OK) return rc;
    if (need) *need = src.size();
    if (buf && cap > 0) { size_t k = std::min(cap - 1, src.size()); std::memcpy(buf, src.data(), k); buf[k] = '\0'; }
    return RMSB_OK;
}

// 1 when large dydt / Jacobian batches of this (mechanism, domain) run on kernels generated for the mechanism (compiles them on
// first use), 0 when they run on the generic kernels (rmsb_last_error says why)
extern "C" int rmsb_mech_specialized(const rmsb_mech* h, int domain_kind) {
    if (!h) return 0;
    std::string why;
    int r = rmsb_spec_ready(h, domain_kind, &why);
    if (!r) g_err = "generic kernels: " + why;
    return r;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// Stage the mechanism descriptor at the start of the block's dynamic shared memory (see mech()).
__device__ __forceinline__ void stage_mech(const DevMech& mp) {
    const int* src = reinterpret_cast<const int*>(&mp);
    int* dst = reinterpret_cast<int*>(smem_ws);
    for (int i = threadIdx.x; i < (int)(sizeof(DevMech) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ Ws make_ws(int warp, int ws_per_warp, int Ny, int mode, int kind) {
    const DevMech& m = mech();
    Ws w;
    w.off = MECH_DOUBLES + warp * ws_per_warp; w.Ns = m.Ns; w.Nr = m.Nr; w.Ny = Ny; w.ndt = ws_ndt_reserved(m.n_dt, mode); w.LD = Ny | 1;
    w.tri = ws_tri(kind);
    w.jpo = ws_jp_offset(m.Ns, m.Nr, Ny, m.n_dt, mode, w.tri);
    return w;
}

// mode 0: dydt, 1: Jacobian, 2: kf/krev, 3: parameter Jacobian, 4: kf before diffusion limitation.  One warp per reactor, grid-stride over the batch.
template <int MODE>
__global__ void __launch_bounds__(128) eval_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ BatchArgs a, int ws_per_warp) {
    stage_mech(mp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const Ws w = make_ws(warp, ws_per_warp, a.Ny, MODE == 3 ? 3 : 0, a.domain_kind);
    const int Ny = a.Ny, Nr = w.Nr;
    const int64_t nw = (int64_t)gridDim.x * wpb;
    for (int64_t ir = (int64_t)blockIdx.x * wpb + warp; ir < a.n; ir += nw) {
        DomainArgs d;
        d.kind = a.domain_kind;
        d.p = a.p ? a.p + ir * a.p_stride : nullptr;
        d.constspc = a.constspc; d.nconst = a.nconst;
        d.cond = a.cond + ir * a.ncond;
        d.flow = a.flow ? a.flow + ir * a.flow_stride : nullptr;
        d.ptab = a.ptab ? a.ptab + ir * a.ptab_stride : nullptr;
        d.t = a.tval ? a.tval[ir] : 0.0;
        double* yb = w.ybuf();
        for (int i = lane; i < Ny; i += 32) yb[i] = a.y[ir * Ny + i];
        rhs_prepare(d, w, lane);
        setup_domain(d, w, lane);
        if (MODE == 0) {
            rhs_warp(d, w, lane, false);
            const double* fb = w.fbuf();
            for (int i = lane; i < Ny; i += 32) a.out[ir * Ny + i] = fb[i];
        } else if (MODE == 1) {
            jac_warp(d, w, lane);
            const int LD = w.LD;
            const double* Js = w.J();
            double* J = a.out + (size_t)ir * Ny * Ny;
            for (int idx = lane; idx < Ny * Ny; idx += 32) {
                int col = idx / Ny, row = idx - col * Ny;
                J[idx] = Js[row + col * LD];
            }
        } else if (MODE == 3) {
            const int Np = w.Ns + Nr;
            double* Jp = a.out + (size_t)ir * Ny * Np;      // column-major Ny x Np
            for (int c0 = 0; c0 < Np; c0 += JP_CHUNK) {
                jacp_warp(d, w, lane, c0, c0 == 0);
                const double* jc = w.jp();
                for (int idx = lane; idx < Ny * JP_CHUNK; idx += 32) {
                    int cl = idx / Ny, row = idx - cl * Ny;
                    if (c0 + cl < Np) Jp[(size_t)(c0 + cl) * Ny + row] = jc[row * JP_CHUNK + cl];
                }
                __syncwarp();
            }
        } else if (MODE == 4) {
            kfs_intrinsic(d, w, lane);
            for (int j = lane; j < Nr; j += 32) a.out[ir * Nr + j] = w.kf()[j];
        } else {
            rhs_warp(d, w, lane, true);
            for (int j = lane; j < Nr; j += 32) { a.out[ir * Nr + j] = w.kf()[j]; a.out2[ir * Nr + j] = w.kr()[j]; }
        }
        __syncwarp();
    }
}

// Resident blocks per SM the register allocation is sized for.  The kernel is latency bound (instruction fetch, dependent
// shared-memory chains): for one row per lane (Ny <= 32) 4 blocks of 4 warps at 128 registers run 26 % faster than 2 blocks at
// 255 (profiles/r01_configs.txt); the wider kernels are limited to one block by their shared-memory Jacobians anyway.
#ifndef RMSB_SOLVE_MINBLOCKS
#define RMSB_SOLVE_MINBLOCKS 4
#endif
template <int R>
__global__ void __launch_bounds__(128, (R == 1 ? RMSB_SOLVE_MINBLOCKS : 1)) solve_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ SolveArgs a, int ws_per_warp) {
    stage_mech(mp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Ws w = make_ws(warp, ws_per_warp, a.Ny, 1, a.domain_kind);
    for (;;) {
        unsigned long long ir = 0;
        if (lane == 0) ir = atomicAdd(a.counter, 1ULL);     // dynamic work queue: one reactor per warp at a time
        ir = __shfl_sync(FULL, ir, 0);
        if (ir >= (unsigned long long)a.n) break;
        bdf_solve_warp<R>(a, w, (int64_t)ir, lane);
    }
}

// forward sensitivities ride on the same integrator (bdf_solve_warp<R, true>)
#ifndef RMSB_SENS_MINBLOCKS
#define RMSB_SENS_MINBLOCKS 1
#endif
template <int R>
__global__ void __launch_bounds__(128, (R == 1 ? RMSB_SENS_MINBLOCKS : 1)) sens_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ SolveArgs a, int ws_per_warp) {
    stage_mech(mp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Ws w = make_ws(warp, ws_per_warp, a.Ny, 2, a.domain_kind);
    for (;;) {
        unsigned long long ir = 0;
        if (lane == 0) ir = atomicAdd(a.counter, 1ULL);
        ir = __shfl_sync(FULL, ir, 0);
        if (ir >= (unsigned long long)a.n) break;
        bdf_solve_warp<R, true>(a, w, (int64_t)ir, lane);
    }
}

// ---- large-mechanism path: one block per reactor (rmsb_large.cuh) -------------------------------------------------------
__device__ __forceinline__ LargeCtx make_large_ctx(const LargeTables& lt, double* scratch, size_t per_block, const DevMech& m) {
    LargeCtx c;
    c.t = lt;
    c.s.base = scratch + (size_t)blockIdx.x * per_block;
    c.s.n = lt.n; c.s.Ns = m.Ns; c.s.Nr = m.Nr; c.s.n_dt = m.n_dt; c.s.n_je = m.n_je; c.s.nnz = lt.nnz;
    c.s.nb = lt.nb; c.s.n_ee = lt.n_ee; c.s.vk = (lt.nb == 2) ? 1 : 0;
    c.red = smem_ws + MECH_DOUBLES;
    c.ru = c.red + 64;
    c.x = c.ru + 36;
    c.map = reinterpret_cast<int*>(c.x + lt.n) + (size_t)(threadIdx.x >> 5) * lt.n;
    return c;
}

#ifndef RMSB_LARGE_MINBLOCKS
#define RMSB_LARGE_MINBLOCKS 4     /* latency bound: 4 resident blocks of 256 threads (64 registers) give 2.1x the throughput of 1 */
#endif
__global__ void __launch_bounds__(LARGE_THREADS, RMSB_LARGE_MINBLOCKS) large_solve_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ SolveArgs a,
                                                                        const __grid_constant__ LargeTables lt, double* scratch, size_t per_block) {
    stage_mech(mp);
    LargeCtx c = make_large_ctx(lt, scratch, per_block, mech());
    __shared__ unsigned long long next;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) next = atomicAdd(a.counter, 1ULL);
        __syncthreads();
        const unsigned long long ir = next;
        if (ir >= (unsigned long long)a.n) break;
        c.d.kind = a.domain_kind;
        c.d.p = a.p ? a.p + ir * a.p_stride : nullptr;
        c.d.constspc = a.constspc; c.d.nconst = a.nconst;
        c.d.cond = a.cond + ir * a.ncond;
        c.T = c.d.cond[0]; c.V = c.P = (a.ncond > 1) ? c.d.cond[1] : 0.0;
        c.t.bv = (a.domain_kind == RMSB_DOMAIN_CONSTANT_P) ? 1 : 0; c.t.has_v = (a.domain_kind == RMSB_DOMAIN_CONSTANT_P || a.domain_kind == RMSB_DOMAIN_CONSTANT_TP) ? 1 : 0;
        large_bdf_solve(a, c, (int64_t)ir);
    }
}

// mode 0: dydt, 1: dense column-major Jacobian (zero-filled, entries scattered), 2: kf / krev
template <int MODE>
__global__ void __launch_bounds__(LARGE_THREADS, 1) large_eval_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ BatchArgs a,
                                                                       const __grid_constant__ LargeTables lt, double* scratch, size_t per_block) {
    stage_mech(mp);
    LargeCtx c = make_large_ctx(lt, scratch, per_block, mech());
    const DevMech& m = mech();
    const int n = lt.n, tid = threadIdx.x;
    for (int64_t ir = blockIdx.x; ir < a.n; ir += gridDim.x) {
        c.d.kind = a.domain_kind;
        c.d.p = a.p ? a.p + ir * a.p_stride : nullptr;
        c.d.constspc = a.constspc; c.d.nconst = a.nconst;
        c.d.cond = a.cond + ir * a.ncond;
        c.T = c.d.cond[0]; c.V = c.P = (a.ncond > 1) ? c.d.cond[1] : 0.0;
        c.t.bv = (a.domain_kind == RMSB_DOMAIN_CONSTANT_P) ? 1 : 0; c.t.has_v = (a.domain_kind == RMSB_DOMAIN_CONSTANT_P || a.domain_kind == RMSB_DOMAIN_CONSTANT_TP) ? 1 : 0;
        double* y = c.s.vec(8);
        for (int i = tid; i < n; i += LARGE_THREADS) y[i] = a.y[ir * n + i];
        __syncthreads();
        if (!c.vk()) large_setup(c, y);
        if (MODE == 0) {
            large_rhs_any(c, y, c.s.vec(14));
            for (int i = tid; i < n; i += LARGE_THREADS) a.out[ir * n + i] = c.s.vec(14)[i];
        } else if (MODE == 1) {
            large_jac(c, y);
            double* J = a.out + (size_t)ir * n * n;
            for (size_t i = tid; i < (size_t)n * n; i += LARGE_THREADS) J[i] = 0.0;
            __syncthreads();
            for (int e = tid; e < m.n_je; e += LARGE_THREADS) J[G(m.je_row, e) + (size_t)G(m.je_col, e) * n] = c.s.jv()[e];
            __syncthreads();
            for (int e = tid; e < lt.n_ee; e += LARGE_THREADS) J[G(lt.ee_row, e) + (size_t)G(lt.ee_col, e) * n] += c.s.effv()[e];
            for (int b = 0; b < lt.nb; ++b) {
                for (int i = tid; i < n; i += LARGE_THREADS) J[i + (size_t)(lt.ns + b) * n] = c.s.bcolv()[(size_t)b * n + i];
                for (int col = tid; col < lt.ns; col += LARGE_THREADS) J[lt.ns + b + (size_t)col * n] = c.s.browv()[(size_t)b * lt.ns + col];
            }
        } else if (MODE == 2 && c.vk()) {
            // kf / krev as calcthermo returns them: the value pass of the variable-k RHS keeps them
            if (a.domain_kind == RMSB_DOMAIN_CONSTANT_P) large_rhs_v<double, true>(c, y, c.s.fd(), true);
            else large_rhs_v<double, false>(c, y, c.s.fd(), true);
            for (int j = tid; j < m.Nr; j += LARGE_THREADS) { a.out[ir * m.Nr + j] = c.s.kf()[j]; a.out2[ir * m.Nr + j] = c.s.kr()[j]; }
        } else if (MODE == 4) {
            // getkfs (src/PhaseState.jl:165-171): forward rate coefficients BEFORE diffusion limitation (kfs_intrinsic, block-strided)
            double acc = 0.0;
            for (int s = tid; s < m.Ns; s += LARGE_THREADS) acc += y[s];
            const double N = block_sum(acc, c.red);
            const bool vk = c.vk(), isp = a.domain_kind == RMSB_DOMAIN_CONSTANT_P;
            const double T = vk ? y[m.Ns] : c.T, V = vk ? (isp ? y[m.Ns + 1] : c.d.cond[0]) : c.vol(y);
            const double P = vk ? (isp ? c.d.cond[0] : y[m.Ns + 1])
                                : ((a.domain_kind == RMSB_DOMAIN_CONSTANT_TV) ? 1.0e8 : (a.domain_kind == RMSB_DOMAIN_CONSTANT_TP ? c.P : 0.0));
            const double C = (vk || a.domain_kind == RMSB_DOMAIN_CONSTANT_TP) ? P / (RMSB_R * T) : (a.domain_kind == RMSB_DOMAIN_CONSTANT_TAPHI ? 0.0 : N / V);
            KinEnv<double> e;
            e.T = T; e.logT = log(T); e.RT = RMSB_R * T; e.invRT = 1.0 / (RMSB_R * T); e.P = P;
            for (int j = tid; j < m.Nr; j += LARGE_THREADS) a.out[ir * m.Nr + j] = eval_kin<double>(j, &e, ceff_of(j, C, y, V), false);
        } else {
            for (int j = tid; j < m.Nr; j += LARGE_THREADS) { a.out[ir * m.Nr + j] = c.s.kf()[j]; a.out2[ir * m.Nr + j] = c.s.kr()[j]; }
        }
        __syncthreads();
    }
}

// ---- dydt with ONE THREAD PER REACTOR for small variable-k mechanisms (ConstantV / ConstantP, Ns <= 32, Nr <= 64) ----------------
// The warp-per-reactor evaluation spends ~1000 warp instructions per state on a 12-species mechanism with 14 of 32 lanes busy; here
// every lane evaluates its own state with the same scalar routines (thermo, rate coefficients, Kc, mass action) in the reference's
// own summation order, ~100 warp instructions per state.  States are staged through shared memory so that global loads / stores
// stay coalesced; the per-thread species / reaction arrays live in (lane-interleaved) local memory.
constexpr int TPR_MAXNS = 32, TPR_MAXNR = 64, TPR_THREADS = 128;
constexpr int TPR_J_MAXNY = 16, TPR_J_MAXNR = 32, TPR_MAXDT = 96;   // Jacobian variant: the dense Ny x Ny block lives in local memory

template <int MS, int MR>
struct TprKeep {        // what the Jacobian needs from the value pass
    double kf[MR], kr[MR], us[MS], cp[MS], cs[MS];
    double T, P, V, Nt, C, Cv, dotU;
};

// dy/dt of one reactor by one thread.  y: [ns; T; P] (ConstantV, cond = V) or [ns; T; V] (ConstantP, cond = P).
// N = Dual2 seeds d/dT (and d/dP for ConstantV) exactly as rhs_v<Dual2> does.
template <class N, bool ISP, int MS, int MR>
__device__ __forceinline__ void rhs_thread(const double* y, double cond, const double* p, const int* constspc, int nconst, N* f, TprKeep<MS, MR>* keep) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    N g[MS], u[MS], q[MR];
    double cs[MS];
    const double V = ISP ? y[Ns + 1] : cond;
    const double invV = 1.0 / V;
    N T = seed<N>(y[Ns], 0), P = ISP ? N(cond) : seed<N>(y[Ns + 1], 1);
    TPow<N> tp(T);
    N RT = RMSB_R * T;
    N invRT = tp.T4 * RMSB_INV_R;
    N Nt = P * V * invRT, C = P * invRT;
    N cpsum = 0.0;
    for (int s = 0; s < Ns; ++s) {
        N cp, h, sr;
        nasa_hscp(s, tp, val(T), cp, h, sr);
        if (p) h = h + p[s] * invRT;
        u[s] = ISP ? h * RT : (h - 1.0) * RT;
        g[s] = h - sr;
        cs[s] = y[s] * invV;
        cpsum = cpsum + cp * y[s];
        if (keep) { keep->us[s] = val(u[s]); keep->cp[s] = val(cp); keep->cs[s] = cs[s]; }
    }
    KinEnv<N> e;
    e.T = T; e.logT = tp.T6; e.RT = RT; e.invRT = invRT; e.P = P;
    N logC0 = RMSB_LOG_1E5_OVER_R - tp.T6;
    for (int j = 0; j < Nr; ++j) {
        N kf;
        if (G(m.ktype, j) == RMSB_KIN_ARRHENIUS_VEC) kf = arr_vec(m.kpar + G(m.koff, j), e);
        else kf = eval_kin<N>(j, &e, ceff_of(j, C, y, V), false);
        N ex = 0.0;
        for (int k = G(m.kc_ptr, j); k < G(m.kc_ptr, j + 1); ++k) ex = ex + G(m.kc_coef, k) * g[G(m.kc_spc, k)];
        ex = ex + G(m.nrp, j) * logC0;
        N kr = kf * dexp(-ex);
        kf = kf * G(m.fwd, j); kr = kr * G(m.rev, j);
        if (p) { kf = kf * p[Ns + j]; kr = kr * p[Ns + j]; }
        if (keep) { keep->kf[j] = val(kf); keep->kr[j] = val(kr); }
        q[j] = net_rate<N>(j, kf, kr, cs);
    }
    N usum = 0.0, nsum = 0.0;
    for (int s = 0; s < Ns; ++s) {
        N acc = 0.0;
        for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
            int jr = G(m.sr_rxn, k);
            N qv = q[jr & 0x7fffffff];
            acc = (jr < 0) ? acc - qv : acc + qv;
        }
        N dn = acc * V;
        f[s] = dn;
        usum = usum + u[s] * dn;
        nsum = nsum + dn;
    }
    N dT = ISP ? -usum / (RMSB_R * cpsum) : -usum / (RMSB_R * (cpsum - Nt));
    f[Ns] = dT;
    f[Ns + 1] = ISP ? nsum * RT / P + dT * V * tp.T4 : nsum * RT * invV + P * tp.T4 * dT;
    for (int k = 0; k < nconst; ++k) f[constspc[k] - 1] = N(0.0);
    if (keep) {
        keep->T = val(T); keep->P = val(P); keep->V = V; keep->Nt = val(Nt); keep->C = val(C);
        keep->Cv = val(cpsum) * (RMSB_R / val(Nt)) - RMSB_R; keep->dotU = val(usum);
    }
}

template <bool ISP>
__global__ void __launch_bounds__(TPR_THREADS) dydt_tpr_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ BatchArgs a) {
    stage_mech(mp);
    const int Ny = a.Ny, LDt = Ny | 1, tid = threadIdx.x;
    double* tile = smem_ws + MECH_DOUBLES;                       // [TPR_THREADS][LDt]
    for (int64_t base = (int64_t)blockIdx.x * TPR_THREADS; base < a.n; base += (int64_t)gridDim.x * TPR_THREADS) {
        const int cnt = (int)((a.n - base < TPR_THREADS) ? a.n - base : TPR_THREADS);
        for (int idx = tid; idx < cnt * Ny; idx += TPR_THREADS) { int r = idx / Ny; tile[r * LDt + (idx - r * Ny)] = a.y[base * Ny + idx]; }
        __syncthreads();
        if (tid < cnt) {
            const int64_t ir = base + tid;
            double* row = tile + tid * LDt;
            double y[TPR_MAXNS + 2], f[TPR_MAXNS + 2];
            for (int i = 0; i < Ny; ++i) y[i] = row[i];
            rhs_thread<double, ISP, TPR_MAXNS, TPR_MAXNR>(y, a.cond[ir], a.p ? a.p + ir * a.p_stride : nullptr, a.constspc, a.nconst, f, nullptr);
            for (int i = 0; i < Ny; ++i) row[i] = f[i];
        }
        __syncthreads();
        for (int idx = tid; idx < cnt * Ny; idx += TPR_THREADS) { int r = idx / Ny; a.out[base * Ny + idx] = tile[r * LDt + (idx - r * Ny)]; }
        __syncthreads();
    }
}

// jacobiany! of one reactor by one thread (Ny <= 16): the steps of jac_warp in the same order, dense block in local memory.
template <bool ISP>
__global__ void __launch_bounds__(TPR_THREADS) jac_tpr_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ BatchArgs a) {
    stage_mech(mp);
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, Ny = a.Ny, LDt = Ny | 1, tid = threadIdx.x;
    double* tile = smem_ws + MECH_DOUBLES;
    for (int64_t base = (int64_t)blockIdx.x * TPR_THREADS; base < a.n; base += (int64_t)gridDim.x * TPR_THREADS) {
        const int cnt = (int)((a.n - base < TPR_THREADS) ? a.n - base : TPR_THREADS);
        for (int idx = tid; idx < cnt * Ny; idx += TPR_THREADS) { int r = idx / Ny; tile[r * LDt + (idx - r * Ny)] = a.y[base * Ny + idx]; }
        __syncthreads();
        if (tid < cnt) {
            const int64_t ir = base + tid;
            const double* row = tile + tid * LDt;
            const double* p = a.p ? a.p + ir * a.p_stride : nullptr;
            double y[TPR_J_MAXNY], f[TPR_J_MAXNY], J[TPR_J_MAXNY * TPR_J_MAXNY], dterm[TPR_MAXDT];
            TprKeep<TPR_J_MAXNY, TPR_J_MAXNR> kp;
            for (int i = 0; i < Ny; ++i) y[i] = row[i];
            rhs_thread<double, ISP, TPR_J_MAXNY, TPR_J_MAXNR>(y, a.cond[ir], p, a.constspc, a.nconst, f, &kp);
            for (int i = 0; i < Ny * Ny; ++i) J[i] = 0.0;
            // (1) derivative terms, (2) entries in the reference's accumulation order (src/Reactor.jl:899-1366)
            for (int t = 0; t < m.n_dt; ++t) {
                int jr = G(m.dt_rxn, t);
                int j = jr & 0x3fffffff;
                double v = (double)G(m.dt_mult, t) * ((jr >> 30) ? kp.kr[j] : kp.kf[j]);
                for (int k = 0; k < 3; ++k) { int s = G(m.dt_oth, 3 * t + k); if (s < 0) break; v *= kp.cs[s]; }
                dterm[t] = v;
            }
            for (int e = 0; e < m.n_je; ++e) {
                double acc = 0.0;
                for (int k = G(m.je_ptr, e); k < G(m.je_ptr, e + 1); ++k) {
                    int ti = G(m.jt_idx, k);
                    double v = dterm[ti & 0x7fffffff];
                    acc = (ti < 0) ? acc - v : acc + v;
                }
                J[G(m.je_row, e) + G(m.je_col, e) * Ny] = acc;
            }
            const int Vcol = Ns + 1;
            if (ISP) {      // (3) V column of the species block (src/Reactor.jl:1368-1413)
                for (int s = 0; s < Ns; ++s) {
                    double acc = 0.0;
                    for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
                        int jr = G(m.sr_rxn, k);
                        int j = jr & 0x7fffffff;
                        double df = 0.0, dr = 0.0;
                        int nre = 0, npr = 0;
                        double pf = kp.kf[j], pr = kp.kr[j];
                        for (int qq = 0; qq < 4; ++qq) { int sp = G(m.rx, qq * Nr + j); if (sp < 0) break; pf *= kp.cs[sp]; ++nre; }
                        for (int qq = 4; qq < 8; ++qq) { int sp = G(m.rx, qq * Nr + j); if (sp < 0) break; pr *= kp.cs[sp]; ++npr; }
                        if (nre > 1) df = -(double)(nre - 1) * pf;
                        if (npr > 1) dr = -(double)(npr - 1) * pr;
                        if (jr < 0) { acc -= df; acc += dr; } else { acc += df; acc -= dr; }
                    }
                    J[s + Vcol * Ny] = acc;
                }
            }
            if (m.n_effrxn > 0) {   // (4) efficiency reactions (src/Domain.jl:2151-2167)
                KinEnv<double> e;
                e.T = kp.T; e.logT = log(kp.T); e.RT = RMSB_R * kp.T; e.invRT = 1.0 / (RMSB_R * kp.T); e.P = kp.P;
                for (int qn = 0; qn < m.n_effrxn; ++qn) {
                    int j = G(m.effrxn, qn);
                    double Kc = kp.kf[j] / kp.kr[j];
                    double acc = 0.0;
                    for (int k = G(m.eff_ptr, j); k < G(m.eff_ptr, j + 1); ++k) acc += kp.cs[G(m.eff_spc, k)] * G(m.eff_val, k);
                    double Ceff = kp.C + acc;
                    double dkdC = dk_dCeff(j, &e, Ceff);
                    double pa = 1.0, pb = 1.0;
                    for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; pa *= kp.cs[sp]; }
                    for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; pb *= kp.cs[sp]; }
                    double pf = pa * kp.V, pr = pb * kp.V;
                    int ne = G(m.eff_ptr, j + 1) - G(m.eff_ptr, j);
                    int ncol = ne + (ISP ? 1 : 0);
                    for (int c = 0; c < ncol; ++c) {
                        int col; double dkf;
                        if (c < ne) { col = G(m.eff_spc, G(m.eff_ptr, j) + c); dkf = dkdC * (G(m.eff_val, G(m.eff_ptr, j) + c) / kp.V); }
                        else { col = Vcol; dkf = dkdC * (kp.C - Ceff) / kp.V; }
                        double dkr = dkf / Kc;
                        double dvf = dkf * pf, dvr = dkr * pr;
                        for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * Ny] -= dvf; }
                        for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * Ny] += dvf; }
                        for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * Ny] -= dvr; }
                        for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * Ny] += dvr; }
                    }
                }
            }
            // (5) thermodynamic-variable rows
            if (!ISP) {
                for (int c = 0; c < Ns; ++c) {          // Domain.jl:2312-2316
                    double dCv = kp.cp[c] * RMSB_R / kp.Nt, sa = 0.0, sm = 0.0;
                    for (int i = 0; i < Ns; ++i) { double v = J[i + c * Ny]; sa += kp.us[i] * v; sm += v; }
                    double jt = -sa / (kp.Nt * kp.Cv) + kp.dotU / (kp.Nt * kp.Cv * kp.Cv) * dCv;
                    J[Ns + c * Ny] = jt;
                    J[Ns + 1 + c * Ny] = sm * RMSB_R * kp.T / kp.V + kp.P / kp.T * jt;
                }
            } else {
                const double aCp = kp.Cv + RMSB_R, dTdt = -kp.dotU / (kp.Nt * aCp);      // Domain.jl:2480-2490
                for (int c = 0; c < Ns + 2; ++c) {
                    if (c == Ns) continue;
                    double sa = 0.0, sm = 0.0;
                    for (int i = 0; i < Ns; ++i) { double v = J[i + c * Ny]; sa += kp.us[i] * v; sm += v; }
                    double jt = -sa / (kp.Nt * aCp);
                    if (c < Ns) jt -= dTdt * ((kp.cp[c] * RMSB_R / kp.Nt) / aCp);
                    J[Ns + c * Ny] = jt;
                    double jv = sm * RMSB_R * kp.T / kp.P + kp.V / kp.T * jt;
                    if (c == Ns + 1) jv += dTdt / kp.T;
                    J[Ns + 1 + c * Ny] = jv;
                }
            }
            // (6) constant species rows
            for (int k = 0; k < a.nconst; ++k) { int s = a.constspc[k] - 1; for (int c = 0; c < Ny; ++c) J[s + c * Ny] = 0.0; }
            // (7) T (and P) columns: derivative of the complete RHS
            {
                Dual2 fd[TPR_J_MAXNY];
                rhs_thread<Dual2, ISP, TPR_J_MAXNY, TPR_J_MAXNR>(y, a.cond[ir], p, a.constspc, a.nconst, fd, nullptr);
                for (int i = 0; i < Ny; ++i) { J[i + Ns * Ny] = fd[i].a; if (!ISP) J[i + (Ns + 1) * Ny] = fd[i].b; }
            }
            double* out = a.out + (size_t)ir * Ny * Ny;
            for (int i = 0; i < Ny * Ny; ++i) out[i] = J[i];
        }
        __syncthreads();
    }
}

__global__ void gibbs_kernel(const __grid_constant__ DevMech mp, int64_t n, const double* T, int scalar_path, double* G) {
    stage_mech(mp);
    const int Ns = mech().Ns;
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t ir = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ir < n; ir += nw) {
        double Tv = T[ir];
        TPow<double> tp(Tv);
        for (int s = lane; s < Ns; s += 32) {
            double g;
            if (scalar_path) g = gibbs_scalar(s, Tv);
            else { double cp, h, sr; nasa_hscp(s, tp, Tv, cp, h, sr); g = (h - sr) * (RMSB_R * Tv); }
            G[ir * Ns + s] = g;
        }
    }
}

__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double b = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
static int rows_per_lane(int Ny) { return (Ny + 31) / 32; }

// cudaGetDeviceProperties takes milliseconds: query each device once
cudaError_t rmsb_cached_props(int device, cudaDeviceProp* out) {
    static std::mutex mu;
    static std::map<int, cudaDeviceProp> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(device);
    if (it == cache.end()) {
        cudaDeviceProp p;
        cudaError_t e = cudaGetDeviceProperties(&p, device);
        if (e != cudaSuccess) return e;
        it = cache.emplace(device, p).first;
    }
    *out = it->second;
    return cudaSuccess;
}

struct LaunchCfg { int wpb; int ws; size_t smem; int blocks_per_sm; int sms; };

template <class K>
static int config_launch(K kernel, const rmsb_mech* h, int kind, int Ny, int mode, int want_wpb, LaunchCfg* cfg) {
    cudaDeviceProp prop;
    CU(rmsb_cached_props(h->device, &prop));
    int ws = ws_doubles(h->dm.Ns, h->dm.Nr, Ny, h->n_dt, mode, ws_tri(kind));
    size_t per_warp = (size_t)ws * sizeof(double);
    const size_t mech_bytes = (size_t)MECH_DOUBLES * sizeof(double);
    size_t max_smem = prop.sharedMemPerBlockOptin - mech_bytes;
    if (per_warp > max_smem)
        return fail(RMSB_ERR_UNSUPPORTED, "mechanism too large for the warp-per-reactor dense kernels (per-warp workspace exceeds shared memory)");
    int wpb = want_wpb > 0 ? want_wpb : 4;
    if (wpb > 4) wpb = 4;
    while (wpb > 1 && per_warp * wpb > max_smem) --wpb;
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb + mech_bytes)));
    int bps = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, wpb * 32, per_warp * wpb + mech_bytes));
    if (bps < 1) bps = 1;
    if (want_wpb <= 0) {
        // the per-warp workspace of a 33..128-variable mechanism is tens of KB: fewer warps per block can pack more warps
        // into the SM's shared memory (e.g. 2 x 3 warps instead of 1 x 4), and these kernels are latency bound
        for (int cand = wpb - 1; cand >= 1; --cand) {
            int b = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, cand * 32, per_warp * cand + mech_bytes));
            if (b * cand > bps * wpb) { bps = b; wpb = cand; }
        }
    }
    size_t smem = per_warp * wpb + mech_bytes;
    cfg->wpb = wpb; cfg->ws = ws; cfg->smem = smem; cfg->blocks_per_sm = bps; cfg->sms = prop.multiProcessorCount;
    return RMSB_OK;
}

static const LargeTables& large_tables(const rmsb_mech* h, int kind) {
    if (kind == RMSB_DOMAIN_CONSTANT_TP) return h->lt_tp;
    if (kind == RMSB_DOMAIN_CONSTANT_V || kind == RMSB_DOMAIN_CONSTANT_P) return h->lt_vk;
    return h->lt;
}
struct LargeCfg { int grid; size_t per_block; size_t smem; };
template <class K>
static int large_config(K kernel, const rmsb_mech* h, int kind, int64_t n, LargeCfg* cfg) {
    if (!h->large) return fail(RMSB_ERR_UNSUPPORTED, "more than 128 state variables need the large-mechanism path, which this mechanism was not analysed for");
    if (kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) return fail(RMSB_ERR_UNSUPPORTED, "large-mechanism path (Ny > 128): coupled gas-surface systems are not built");
    const LargeTables& lt = large_tables(h, kind);
    if (lt.n == 0) return fail(RMSB_ERR_UNSUPPORTED, "large-mechanism path: this domain kind was not analysed for this phase");
    cudaDeviceProp prop;
    CU(rmsb_cached_props(h->device, &prop));
    cfg->smem = large_smem_bytes(lt.n);
    if (cfg->smem > prop.sharedMemPerBlockOptin) return fail(RMSB_ERR_UNSUPPORTED, "mechanism too large for the block-per-reactor kernels (shared memory)");
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
    int bps = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, LARGE_THREADS, cfg->smem));
    if (bps < 1) bps = 1;
    if (bps > 16) bps = 16;
    int64_t resident = (int64_t)prop.multiProcessorCount * bps;
    cfg->grid = (int)(n < resident ? (n < 1 ? 1 : n) : resident);
    cfg->per_block = (large_scratch_doubles(lt.n, h->dm.Ns, h->dm.Nr, h->n_dt, h->n_je, lt.nnz, lt.nb, lt.n_ee, lt.nb == 2 ? 1 : 0) + 1) & ~(size_t)1;
    return RMSB_OK;
}

template <int MODE>
static int launch_eval_large(const rmsb_mech* h, const BatchArgs& a, cudaStream_t st) {
    LargeCfg cfg;
    int rc = large_config(large_eval_kernel<MODE>, h, a.domain_kind, a.n, &cfg);
    if (rc != RMSB_OK) return rc;
    double* scratch = nullptr;
    CU(cudaMallocAsync((void**)&scratch, sizeof(double) * cfg.per_block * cfg.grid, st));
    large_eval_kernel<MODE><<<cfg.grid, LARGE_THREADS, cfg.smem, st>>>(h->dm, a, large_tables(h, a.domain_kind), scratch, cfg.per_block);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return fail(RMSB_ERR_CUDA, std::string("large_eval_kernel: ") + cudaGetErrorString(e));
    return RMSB_OK;
}

static int launch_solve_large(const rmsb_mech* h, const SolveArgs& a, cudaStream_t st, size_t* work_bytes) {
    LargeCfg cfg;
    int rc = large_config(large_solve_kernel, h, a.domain_kind, a.n, &cfg);
    if (rc != RMSB_OK) return rc;
    if (work_bytes) { *work_bytes = 256 + sizeof(double) * cfg.per_block * cfg.grid; return RMSB_OK; }
    CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st));
    double* scratch = reinterpret_cast<double*>(reinterpret_cast<char*>(a.counter) + 256);
    large_solve_kernel<<<cfg.grid, LARGE_THREADS, cfg.smem, st>>>(h->dm, a, large_tables(h, a.domain_kind), scratch, cfg.per_block);
    ++g_launches;
    CU(cudaGetLastError());
    return RMSB_OK;
}

// RMSB_NO_TPR=1 in the environment routes large batches of small mechanisms to the lane-group kernels instead of the
// thread-per-reactor ones (A/B measurements, scripts/prof_evals.py)
static bool tpr_disabled() { const char* v = std::getenv("RMSB_NO_TPR"); return v && v[0] == '1'; }

template <int MODE>
static int launch_eval(const rmsb_mech* h, const BatchArgs& a, cudaStream_t st) {
    const bool extras = a.flow || a.ptab;
    if (extras) {
        // boundary interfaces / time profiles are evaluated by the warp-per-reactor routines (rhs_v / rhs_ck / jac_warp)
        if (a.Ny > 128) return fail(RMSB_ERR_UNSUPPORTED, "boundary interfaces / time profiles on the large-mechanism path (Ny > 128) are not built");
        if (MODE != 0 && MODE != 1 && MODE != 3) return fail(RMSB_ERR_UNSUPPORTED, "boundary interfaces / time profiles: dydt, Jacobian, parameter Jacobian and solve only");
        if (a.domain_kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) return fail(RMSB_ERR_UNSUPPORTED, "boundary interfaces on the coupled gas-surface system are not built");
    }
    if (a.Ny > 128) {
        if (MODE <= 2 || MODE == 4) return launch_eval_large<((MODE <= 2 || MODE == 4) ? MODE : 0)>(h, a, st);
        return fail(RMSB_ERR_UNSUPPORTED, "Ny > 128: this evaluation is not built for the large-mechanism path");
    }
    if ((MODE == 0 || MODE == 1) && a.n >= 4096 && !extras && !tpr_disabled() && rmsb_spec_ready(h, a.domain_kind, nullptr))
        return rmsb_spec_launch_eval(h, MODE, a, st);          // straight-line kernels generated for this mechanism (rmsb_spec.cu)
    if (MODE == 0 && (a.domain_kind == RMSB_DOMAIN_CONSTANT_V || a.domain_kind == RMSB_DOMAIN_CONSTANT_P) &&
        h->dm.Ns <= TPR_MAXNS && h->dm.Nr <= TPR_MAXNR && a.n >= 4096 && !tpr_disabled() && !extras) {
        // small variable-k mechanisms, large batches: one thread per reactor
        cudaDeviceProp prop;
        CU(rmsb_cached_props(h->device, &prop));
        const size_t smem = ((size_t)MECH_DOUBLES + (size_t)TPR_THREADS * (a.Ny | 1)) * sizeof(double);
        int64_t blocks = (a.n + TPR_THREADS - 1) / TPR_THREADS, cap = (int64_t)prop.multiProcessorCount * 16;
        const int grid = (int)(blocks < cap ? blocks : cap);
        if (a.domain_kind == RMSB_DOMAIN_CONSTANT_P) dydt_tpr_kernel<true><<<grid, TPR_THREADS, smem, st>>>(h->dm, a);
        else dydt_tpr_kernel<false><<<grid, TPR_THREADS, smem, st>>>(h->dm, a);
        ++g_launches;
        CU(cudaGetLastError());
        return RMSB_OK;
    }
    if (MODE == 1 && (a.domain_kind == RMSB_DOMAIN_CONSTANT_V || a.domain_kind == RMSB_DOMAIN_CONSTANT_P) && a.Ny <= TPR_J_MAXNY &&
        h->dm.Nr <= TPR_J_MAXNR && h->n_dt <= TPR_MAXDT && a.n >= 4096 && !tpr_disabled() && !extras) {
        cudaDeviceProp prop;
        CU(rmsb_cached_props(h->device, &prop));
        const size_t smem = ((size_t)MECH_DOUBLES + (size_t)TPR_THREADS * (a.Ny | 1)) * sizeof(double);
        int64_t blocks = (a.n + TPR_THREADS - 1) / TPR_THREADS, cap = (int64_t)prop.multiProcessorCount * 16;
        const int grid = (int)(blocks < cap ? blocks : cap);
        if (a.domain_kind == RMSB_DOMAIN_CONSTANT_P) jac_tpr_kernel<true><<<grid, TPR_THREADS, smem, st>>>(h->dm, a);
        else jac_tpr_kernel<false><<<grid, TPR_THREADS, smem, st>>>(h->dm, a);
        ++g_launches;
        CU(cudaGetLastError());
        return RMSB_OK;
    }
    if ((MODE == 0 || MODE == 1) && !extras && rmsb_small_eligible(h, a.domain_kind, a.Ny, 0)) return rmsb_small_launch_eval(h, MODE, a, st);
    LaunchCfg cfg;
    int rc = config_launch(eval_kernel<MODE>, h, a.domain_kind, a.Ny, MODE == 3 ? 3 : 0, 4, &cfg);
    if (rc != RMSB_OK) return rc;
    int64_t blocks_needed = (a.n + cfg.wpb - 1) / cfg.wpb;
    int64_t max_blocks = (int64_t)cfg.sms * cfg.blocks_per_sm * 8;   // grid-stride beyond 8 waves
    int grid = (int)(blocks_needed < max_blocks ? blocks_needed : max_blocks);
    if (grid < 1) grid = 1;
    eval_kernel<MODE><<<grid, cfg.wpb * 32, cfg.smem, st>>>(h->dm, a, cfg.ws);
    ++g_launches;
    CU(cudaGetLastError());
    return RMSB_OK;
}

template <int R>
static int launch_solve_R(const rmsb_mech* h, const SolveArgs& a, int want_wpb, cudaStream_t st) {
    LaunchCfg cfg;
    int rc = config_launch(solve_kernel<R>, h, a.domain_kind, a.Ny, 1, want_wpb, &cfg);
    if (rc != RMSB_OK) return rc;
    int64_t resident = (int64_t)cfg.sms * cfg.blocks_per_sm;
    int64_t blocks_needed = (a.n + cfg.wpb - 1) / cfg.wpb;
    int grid = (int)(blocks_needed < resident ? blocks_needed : resident);
    if (grid < 1) grid = 1;
    CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st));
    solve_kernel<R><<<grid, cfg.wpb * 32, cfg.smem, st>>>(h->dm, a, cfg.ws);
    ++g_launches;
    CU(cudaGetLastError());
    return RMSB_OK;
}

constexpr size_t SENS_WORK_HEADER = 256;
static int sens_npp(const rmsb_mech* h) { return (h->dm.Ns + h->dm.Nr + 31) & ~31; }

// resident warps of the sensitivity kernel and the scratch they need; launches when `a` is given
template <int R>
static int sens_R(const rmsb_mech* h, int kind, int Ny, int64_t n, int want_wpb, size_t* work_bytes, SolveArgs* a, cudaStream_t st) {
    LaunchCfg cfg;
    int rc = config_launch(sens_kernel<R>, h, kind, Ny, 2, want_wpb, &cfg);
    if (rc != RMSB_OK) return rc;
    int64_t resident = (int64_t)cfg.sms * cfg.blocks_per_sm;
    int64_t blocks_needed = (n + cfg.wpb - 1) / cfg.wpb;
    int grid = (int)(blocks_needed < resident ? blocks_needed : resident);
    if (grid < 1) grid = 1;
    size_t per_warp = sizeof(double) * (size_t)(BDF_MAX_ORDER + 3) * Ny * sens_npp(h);
    if (work_bytes) *work_bytes = SENS_WORK_HEADER + per_warp * (size_t)grid * cfg.wpb;
    if (!a) return RMSB_OK;
    CU(cudaMemsetAsync(a->counter, 0, sizeof(unsigned long long), st));
    sens_kernel<R><<<grid, cfg.wpb * 32, cfg.smem, st>>>(h->dm, *a, cfg.ws);
    ++g_launches;
    CU(cudaGetLastError());
    return RMSB_OK;
}

static int sens_dispatch(const rmsb_mech* h, int kind, int Ny, int64_t n, int want_wpb, size_t* work_bytes, SolveArgs* a, cudaStream_t st) {
    switch (rows_per_lane(Ny)) {
    case 1: return sens_R<1>(h, kind, Ny, n, want_wpb, work_bytes, a, st);
    case 2: return sens_R<2>(h, kind, Ny, n, want_wpb, work_bytes, a, st);
    case 3: return sens_R<3>(h, kind, Ny, n, want_wpb, work_bytes, a, st);
    case 4: return sens_R<4>(h, kind, Ny, n, want_wpb, work_bytes, a, st);
    }
    return fail(RMSB_ERR_UNSUPPORTED, "Ny > 128: needs the sparse CTA-per-reactor path (not built yet, see DESIGN.md)");
}

static int launch_solve(const rmsb_mech* h, const SolveArgs& a, int want_wpb, cudaStream_t st) {
    if (a.flow || a.ptab) {      // boundary interfaces / time profiles: the warp-per-reactor integrator (rhs_v / rhs_ck / jac_warp carry the terms)
        if (a.Ny > 128) return fail(RMSB_ERR_UNSUPPORTED, "boundary interfaces / time profiles on the large-mechanism path (Ny > 128) are not built");
        if (a.domain_kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) return fail(RMSB_ERR_UNSUPPORTED, "boundary interfaces on the coupled gas-surface system are not built");
    } else if (rmsb_small_eligible(h, a.domain_kind, a.Ny, a.flags)) return rmsb_small_launch_solve(h, a, st);   // lane group per reactor (rmsb_small.cu)
    switch (rows_per_lane(a.Ny)) {
    case 1: return launch_solve_R<1>(h, a, want_wpb, st);
    case 2: return launch_solve_R<2>(h, a, want_wpb, st);
    case 3: return launch_solve_R<3>(h, a, want_wpb, st);
    case 4: return launch_solve_R<4>(h, a, want_wpb, st);
    }
    return launch_solve_large(h, a, st, nullptr);
}

static int check_common(const rmsb_mech* h, int kind, int64_t n, const void* y, const void* cond, int64_t p_stride, const void* p) {
    if (!h) return fail(RMSB_ERR_ARG, "null mechanism");
    if (domain_ny(h, kind) < 0) return fail(RMSB_ERR_ARG, "unknown domain kind");
    if (n < 0) return fail(RMSB_ERR_ARG, "negative batch size");
    if (n > 0 && (!y || !cond)) return fail(RMSB_ERR_ARG, "null state/condition buffer");
    if (p && p_stride != 0 && p_stride != h->dm.Ns + h->dm.Nr) return fail(RMSB_ERR_ARG, "p_stride must be 0 or Ns+Nr");
    if ((kind == RMSB_DOMAIN_CONSTANT_V || kind == RMSB_DOMAIN_CONSTANT_P) && h->dm.phase_kind != RMSB_PHASE_GAS)
        return fail(RMSB_ERR_ARG, "ConstantV / ConstantP domains need an IdealGas phase");
    if (kind == RMSB_DOMAIN_CONSTANT_TAPHI && h->dm.phase_kind != RMSB_PHASE_SURFACE) return fail(RMSB_ERR_ARG, "ConstantTAPhi domain needs an IdealSurface phase");
    if ((kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) != (h->dm.coupled != 0)) return fail(RMSB_ERR_ARG, "the coupled gas-surface domain and two-phase mechanisms go together");
    return RMSB_OK;
}

// const_species of the host-buffer entry points: 1-based species indices (the *_dev variants take a device array and document the
// requirement instead)
static int check_const_species(const rmsb_mech* h, const int32_t* cs, int32_t ncs) {
    if (!cs || ncs <= 0) return RMSB_OK;
    for (int k = 0; k < ncs; ++k)
        if (cs[k] < 1 || cs[k] > h->dm.Ns) return fail(RMSB_ERR_ARG, "const_species entries are 1-based species indices in 1..Ns");
    return RMSB_OK;
}

// ------------------------------------------------------------------------------------------------
// device-pointer entry points
// ------------------------------------------------------------------------------------------------
static BatchArgs make_batch(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                            int64_t p_stride, const int32_t* cs, int32_t ncs, double* out, double* out2) {
    BatchArgs a;
    a.n = n; a.y = y; a.cond = cond; a.ncond = domain_ncond(kind); a.p = p; a.p_stride = p ? p_stride : 0;
    a.constspc = cs; a.nconst = cs ? ncs : 0; a.out = out; a.out2 = out2; a.domain_kind = kind; a.Ny = domain_ny(h, kind);
    return a;
}

// validated copy of the caller's rmsb_extras (NULL / short structs give the defaults)
struct Extras { const double* ifs = nullptr; int64_t if_stride = 0; const double* prof = nullptr; int64_t prof_stride = 0; const double* t = nullptr; int64_t prof_len = 0; };
static int check_extras(const rmsb_mech* h, int kind, const rmsb_extras* x, Extras* e) {
    *e = Extras();
    if (!x) return RMSB_OK;
    if (x->struct_size != (int32_t)sizeof(rmsb_extras)) return fail(RMSB_ERR_ARG, "rmsb_extras size mismatch (ABI)");
    e->ifs = x->interfaces; e->if_stride = x->interfaces ? x->if_stride : 0;
    e->prof = x->profile; e->prof_stride = x->profile ? x->profile_stride : 0; e->t = x->t;
    if (e->ifs && e->if_stride != 0 && e->if_stride != 7 + 2 * (int64_t)h->dm.Ns) return fail(RMSB_ERR_ARG, "if_stride must be 0 or 7 + 2*Ns");
    if (e->prof) {
        if (kind != RMSB_DOMAIN_CONSTANT_V && kind != RMSB_DOMAIN_CONSTANT_P && kind != RMSB_DOMAIN_CONSTANT_TP && kind != RMSB_DOMAIN_CONSTANT_TV)
            return fail(RMSB_ERR_UNSUPPORTED, "time profiles: ConstantV (V(t)), ConstantP (P(t)), ConstantTP (T(t), P(t)) and ConstantTV (T(t)) layouts only");
        const int64_t nt = (int64_t)e->prof[0];
        if (nt < 2 || (double)nt != e->prof[0]) return fail(RMSB_ERR_ARG, "profile: the first entry is the number of knots (>= 2)");
        e->prof_len = 1 + 3 * nt;
        for (int64_t k = 1; k < nt; ++k) if (!(e->prof[1 + k] > e->prof[k])) return fail(RMSB_ERR_ARG, "profile: knot times must increase");
        if (kind == RMSB_DOMAIN_CONSTANT_TP) {                 // ParametrizedTP: the T table is followed by the P table
            const int64_t nt2 = (int64_t)e->prof[e->prof_len];
            if (nt2 < 2 || (double)nt2 != e->prof[e->prof_len]) return fail(RMSB_ERR_ARG, "profile: ParametrizedTP needs a T table followed by a P table");
            e->prof_len += 1 + 3 * nt2;
        }
        if (e->prof_stride != 0 && e->prof_stride != e->prof_len) return fail(RMSB_ERR_ARG, "profile_stride must be 0 or the length of one reactor's table(s)");
    }
    return RMSB_OK;
}

extern "C" int rmsb_dydt_batch_dev(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond,
                                   const double* p, int64_t p_stride, const int32_t* cs, int32_t ncs, double* dydt, void* stream) {
    int rc = check_common(h, kind, n, y, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    DeviceGuard g(h->device);
    return launch_eval<0>(h, make_batch(h, kind, n, y, cond, p, p_stride, cs, ncs, dydt, nullptr), (cudaStream_t)stream);
}

extern "C" int rmsb_jac_batch_dev(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond,
                                  const double* p, int64_t p_stride, const int32_t* cs, int32_t ncs, double* J, void* stream) {
    int rc = check_common(h, kind, n, y, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    DeviceGuard g(h->device);
    return launch_eval<1>(h, make_batch(h, kind, n, y, cond, p, p_stride, cs, ncs, J, nullptr), (cudaStream_t)stream);
}

static int check_jacp(const rmsb_mech* h, int kind, const double* p, bool profile = false) {
    if (kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) return fail(RMSB_ERR_UNSUPPORTED, "parameter Jacobian of the coupled gas-surface system is not built");
    if (kind == RMSB_DOMAIN_CONSTANT_TV && !p && !profile) return fail(RMSB_ERR_ARG, "ConstantTV parameter Jacobian needs the explicit parameter vector p (chain rule through the diffusion limit)");
    (void)h;
    return RMSB_OK;
}

extern "C" int rmsb_jacp_batch_dev(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond,
                                   const double* p, int64_t p_stride, const int32_t* cs, int32_t ncs, double* Jp, void* stream) {
    int rc = check_common(h, kind, n, y, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    if ((rc = check_jacp(h, kind, p)) != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    DeviceGuard g(h->device);
    return launch_eval<3>(h, make_batch(h, kind, n, y, cond, p, p_stride, cs, ncs, Jp, nullptr), (cudaStream_t)stream);
}

extern "C" int64_t rmsb_solve_work_bytes(const rmsb_mech* h, int kind, int64_t n) {
    if (!h || domain_ny(h, kind) < 0) { fail(RMSB_ERR_ARG, "bad mechanism / domain kind"); return -1; }
    if (domain_ny(h, kind) <= 128) return 256 + (int64_t)rmsb_small_work_bytes(h, kind, domain_ny(h, kind), n);
    DeviceGuard g(h->device);
    SolveArgs a{};
    a.domain_kind = kind; a.n = n > 0 ? n : 1;
    size_t bytes = 0;
    if (launch_solve_large(h, a, nullptr, &bytes) != RMSB_OK) return -1;
    return (int64_t)bytes;
}

constexpr int32_t RMSB_SOLVE_OPTS_V1_SIZE = 32;    // rmsb_solve_opts before dysave was added
static void default_opts(rmsb_solve_opts* o, const rmsb_solve_opts* in) {
    o->struct_size = sizeof(rmsb_solve_opts);
    o->max_steps = 500000; o->max_order = 5; o->idt_index = -1; o->first_step = 0.0; o->warps_per_block = 0; o->reserved = 0;
    o->dysave = nullptr;
    if (in && (in->struct_size == (int32_t)sizeof(rmsb_solve_opts) || in->struct_size == RMSB_SOLVE_OPTS_V1_SIZE)) {
        if (in->struct_size == (int32_t)sizeof(rmsb_solve_opts)) o->dysave = in->dysave;
        if (in->max_steps > 0) o->max_steps = in->max_steps;
        if (in->max_order >= 1 && in->max_order <= 5) o->max_order = in->max_order;
        o->idt_index = in->idt_index;
        if (in->first_step > 0.0) o->first_step = in->first_step;
        o->warps_per_block = in->warps_per_block;
        o->reserved = in->reserved;
    }
}

static int solve_dev_impl(bool sens, const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond,
                          const double* p, int64_t p_stride, const int32_t* cs, int32_t ncs,
                          const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave,
                          double* ysave, double* ssave, double* yend, double* send, double* idt, int32_t* status, rmsb_stats* stats,
                          const rmsb_solve_opts* opts, void* work, void* stream, const double* flow = nullptr, int64_t flow_stride = 0,
                          const double* ptab = nullptr, int64_t ptab_stride = 0) {
    int rc = check_common(h, kind, n, y0, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    if (sens && (rc = check_jacp(h, kind, p, ptab != nullptr)) != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    if (!t_end || !status || !work) return fail(RMSB_ERR_ARG, "t_end, status and work buffers are required");
    if (!(rtol > 0.0) || !(atol >= 0.0)) return fail(RMSB_ERR_ARG, "rtol must be > 0 and atol >= 0");
    if (nsave < 0 || (nsave > 0 && !tsave)) return fail(RMSB_ERR_ARG, "bad tsave");
    rmsb_solve_opts o;
    default_opts(&o, opts);
    SolveArgs a;
    a.n = n; a.y0 = y0; a.cond = cond; a.ncond = domain_ncond(kind); a.p = p; a.p_stride = p ? p_stride : 0;
    a.constspc = cs; a.nconst = cs ? ncs : 0; a.t_end = t_end; a.rtol = rtol; a.atol = atol;
    a.nsave = nsave; a.tsave = tsave; a.ysave = ysave; a.yend = yend; a.idt = idt; a.status = status; a.stats = stats;
    a.max_steps = o.max_steps; a.max_order = o.max_order; a.first_step = o.first_step;
    a.domain_kind = kind; a.Ny = domain_ny(h, kind); a.flags = o.reserved;
    a.idt_index = o.idt_index >= 0 ? o.idt_index : ((kind == RMSB_DOMAIN_CONSTANT_V || kind == RMSB_DOMAIN_CONSTANT_P) ? h->dm.Ns : -1);
    if (a.idt_index >= a.Ny) return fail(RMSB_ERR_ARG, "idt_index out of range");
    a.counter = reinterpret_cast<unsigned long long*>(work);
    a.work_bytes = sens ? 256 : (size_t)std::max<int64_t>(256, rmsb_solve_work_bytes(h, kind, n));     // the caller's contract (rmsb200.h)
    a.ssave = ssave; a.send = send; a.sens_hist = nullptr; a.Npp = sens_npp(h);
    a.flow = flow; a.flow_stride = flow ? flow_stride : 0;
    a.ptab = ptab; a.ptab_stride = ptab ? ptab_stride : 0;
    DeviceGuard g(h->device);
    if (!sens) return launch_solve(h, a, o.warps_per_block, (cudaStream_t)stream);
    a.idt_index = -1; a.idt = nullptr;
    a.sens_hist = reinterpret_cast<double*>(static_cast<char*>(work) + SENS_WORK_HEADER);
    return sens_dispatch(h, a.domain_kind, a.Ny, n, o.warps_per_block, nullptr, &a, (cudaStream_t)stream);
}

extern "C" int rmsb_solve_batch_dev(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond,
                                    const double* p, int64_t p_stride, const int32_t* cs, int32_t ncs,
                                    const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave,
                                    double* ysave, double* yend, double* idt, int32_t* status, rmsb_stats* stats,
                                    const rmsb_solve_opts* opts, void* work, void* stream) {
    return solve_dev_impl(false, h, kind, n, y0, cond, p, p_stride, cs, ncs, t_end, rtol, atol, nsave, tsave, ysave, nullptr,
                          yend, nullptr, idt, status, stats, opts, work, stream);
}

extern "C" int64_t rmsb_sens_work_bytes(const rmsb_mech* h, int kind, int64_t n) {
    if (!h || domain_ny(h, kind) < 0) { fail(RMSB_ERR_ARG, "bad mechanism / domain kind"); return -1; }
    DeviceGuard g(h->device);
    // the history buffer is indexed by resident warp (blockIdx.x * warps_per_block + warp): size it for the launch geometry that
    // needs the most warps over every warps_per_block the caller may pass in rmsb_solve_opts (0 = automatic, 1..4 fixed)
    size_t bytes = 0;
    for (int wpb = 0; wpb <= 4; ++wpb) {
        size_t b = 0;
        if (sens_dispatch(h, kind, domain_ny(h, kind), n > 0 ? n : 1, wpb, &b, nullptr, nullptr) != RMSB_OK) return -1;
        if (b > bytes) bytes = b;
    }
    return (int64_t)bytes;
}

extern "C" int rmsb_sens_batch_dev(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond,
                                   const double* p, int64_t p_stride, const int32_t* cs, int32_t ncs,
                                   const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave,
                                   double* ysave, double* ssave, double* yend, double* send, int32_t* status, rmsb_stats* stats,
                                   const rmsb_solve_opts* opts, void* work, void* stream) {
    return solve_dev_impl(true, h, kind, n, y0, cond, p, p_stride, cs, ncs, t_end, rtol, atol, nsave, tsave, ysave, ssave,
                          yend, send, nullptr, status, stats, opts, work, stream);
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points: copy in, run, copy out (stream-ordered pool allocations)
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 8, st); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
struct StreamGuard {
    cudaStream_t st = nullptr;
    ~StreamGuard() { if (st) cudaStreamDestroy(st); }
};
static void tune_pool(int device) {
    static std::atomic<unsigned> done{0};
    unsigned bit = 1u << (device & 31);
    if (done.load() & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ULL;   // keep freed blocks cached: repeated batch calls re-use them
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done.fetch_or(bit);
}

// device time of the kernels of the last host-buffer solve of this thread (rmsb_last_kernel_ms)
static thread_local double g_last_kernel_ms = -1.0;
extern "C" double rmsb_last_kernel_ms(void) { return g_last_kernel_ms; }
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    cudaError_t start(cudaStream_t st) {
        cudaError_t e = cudaEventCreate(&a);
        if (e == cudaSuccess) e = cudaEventCreate(&b);
        if (e == cudaSuccess) e = cudaEventRecord(a, st);
        return e;
    }
    cudaError_t stop(cudaStream_t st) { return cudaEventRecord(b, st); }
    void read() { float ms = -1.f; if (a && b && cudaEventElapsedTime(&ms, a, b) == cudaSuccess) g_last_kernel_ms = ms; }
};

// out[(r*rep + k)][:] = in[r][:]: per-reactor constants repeated for each of a reactor's saved states
__global__ void expand_rows_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t n, int rep, int width) {
    const int64_t tot = n * rep * (int64_t)width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / width;
        out[i] = in[(row / rep) * width + (i - row * width)];
    }
}

#define H2D(dst, src, bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sg.st))
#define D2H(dst, src, bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, sg.st))

static int eval_host(int mode, const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                     int64_t p_stride, const int32_t* cs, int32_t ncs, double* out, double* out2,
                     const rmsb_extras* xtr = nullptr) {
    int rc = check_common(h, kind, n, y, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    Extras ex;
    if ((rc = check_extras(h, kind, xtr, &ex)) != RMSB_OK) return rc;
    if (mode == 3 && (rc = check_jacp(h, kind, p, ex.prof != nullptr)) != RMSB_OK) return rc;
    if ((rc = check_const_species(h, cs, ncs)) != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    if (!out || (mode == 2 && !out2)) return fail(RMSB_ERR_ARG, "null output buffer");
    DeviceGuard g(h->device);
    if (!g.ok) return fail(RMSB_ERR_CUDA, "cudaSetDevice failed");
    tune_pool(h->device);
    StreamGuard sg;
    CU(cudaStreamCreateWithFlags(&sg.st, cudaStreamNonBlocking));
    const int Ny = domain_ny(h, kind), nc = domain_ncond(kind), Np = h->dm.Ns + h->dm.Nr;
    size_t out_per = mode == 0 ? (size_t)Ny : (mode == 1 ? (size_t)Ny * Ny : (mode == 3 ? (size_t)Ny * Np : (size_t)h->dm.Nr));
    DevBuf dy(sg.st), dc(sg.st), dp(sg.st), dcs(sg.st), dout(sg.st), dout2(sg.st);
    CU(dy.alloc(sizeof(double) * n * Ny)); CU(dc.alloc(sizeof(double) * n * nc)); CU(dout.alloc(sizeof(double) * n * out_per));
    H2D(dy.p, y, sizeof(double) * n * Ny);
    H2D(dc.p, cond, sizeof(double) * n * nc);
    if (p) { size_t np = p_stride ? (size_t)n * Np : (size_t)Np; CU(dp.alloc(sizeof(double) * np)); H2D(dp.p, p, sizeof(double) * np); }
    if (cs && ncs > 0) { CU(dcs.alloc(sizeof(int32_t) * ncs)); H2D(dcs.p, cs, sizeof(int32_t) * ncs); }
    if (mode == 2) CU(dout2.alloc(sizeof(double) * n * out_per));
    BatchArgs a = make_batch(h, kind, n, dy.as<double>(), dc.as<double>(), p ? dp.as<double>() : nullptr, p_stride,
                             (cs && ncs > 0) ? dcs.as<int32_t>() : nullptr, ncs, dout.as<double>(), dout2.as<double>());
    DevBuf dif(sg.st), dpr(sg.st), dtv(sg.st);
    if (ex.ifs) {
        const size_t nif = (size_t)(ex.if_stride ? n : 1) * (7 + 2 * (size_t)h->dm.Ns);
        CU(dif.alloc(sizeof(double) * nif)); H2D(dif.p, ex.ifs, sizeof(double) * nif);
        a.flow = dif.as<double>(); a.flow_stride = ex.if_stride;
    }
    if (ex.prof) {
        const size_t np = (size_t)(ex.prof_stride ? n : 1) * (size_t)ex.prof_len;
        CU(dpr.alloc(sizeof(double) * np)); H2D(dpr.p, ex.prof, sizeof(double) * np);
        a.ptab = dpr.as<double>(); a.ptab_stride = ex.prof_stride;
    }
    if (ex.t) { CU(dtv.alloc(sizeof(double) * n)); H2D(dtv.p, ex.t, sizeof(double) * n); a.tval = dtv.as<double>(); }
    if (mode == 0) rc = launch_eval<0>(h, a, sg.st);
    else if (mode == 1) rc = launch_eval<1>(h, a, sg.st);
    else if (mode == 3) rc = launch_eval<3>(h, a, sg.st);
    else if (mode == 4) rc = launch_eval<4>(h, a, sg.st);
    else rc = launch_eval<2>(h, a, sg.st);
    if (rc != RMSB_OK) return rc;
    D2H(out, dout.p, sizeof(double) * n * out_per);
    if (mode == 2) D2H(out2, dout2.p, sizeof(double) * n * out_per);
    CU(cudaStreamSynchronize(sg.st));
    return RMSB_OK;
}

extern "C" int rmsb_dydt_batch(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                               int64_t p_stride, const int32_t* cs, int32_t ncs, double* dydt) {
    return eval_host(0, h, kind, n, y, cond, p, p_stride, cs, ncs, dydt, nullptr);
}
extern "C" int rmsb_jac_batch(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                              int64_t p_stride, const int32_t* cs, int32_t ncs, double* J) {
    return eval_host(1, h, kind, n, y, cond, p, p_stride, cs, ncs, J, nullptr);
}
extern "C" int rmsb_jacp_batch(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                               int64_t p_stride, const int32_t* cs, int32_t ncs, double* Jp) {
    return eval_host(3, h, kind, n, y, cond, p, p_stride, cs, ncs, Jp, nullptr);
}
static rmsb_extras extras_of_interfaces(const double* ifs, int64_t if_stride) {
    rmsb_extras x;
    std::memset(&x, 0, sizeof(x));
    x.struct_size = sizeof(x); x.interfaces = ifs; x.if_stride = if_stride;
    return x;
}
extern "C" int rmsb_dydt_batch_ex(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                                  int64_t p_stride, const int32_t* cs, int32_t ncs, const rmsb_extras* extras, double* dydt) {
    return eval_host(0, h, kind, n, y, cond, p, p_stride, cs, ncs, dydt, nullptr, extras);
}
extern "C" int rmsb_jac_batch_ex(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                                 int64_t p_stride, const int32_t* cs, int32_t ncs, const rmsb_extras* extras, double* J) {
    return eval_host(1, h, kind, n, y, cond, p, p_stride, cs, ncs, J, nullptr, extras);
}
extern "C" int rmsb_jacp_batch_ex(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                                  int64_t p_stride, const int32_t* cs, int32_t ncs, const rmsb_extras* extras, double* Jp) {
    return eval_host(3, h, kind, n, y, cond, p, p_stride, cs, ncs, Jp, nullptr, extras);
}
extern "C" int rmsb_dydt_batch_if(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                                  int64_t p_stride, const int32_t* cs, int32_t ncs, const double* ifs, int64_t if_stride, double* dydt) {
    rmsb_extras x = extras_of_interfaces(ifs, if_stride);
    return eval_host(0, h, kind, n, y, cond, p, p_stride, cs, ncs, dydt, nullptr, &x);
}
extern "C" int rmsb_jac_batch_if(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                                 int64_t p_stride, const int32_t* cs, int32_t ncs, const double* ifs, int64_t if_stride, double* J) {
    rmsb_extras x = extras_of_interfaces(ifs, if_stride);
    return eval_host(1, h, kind, n, y, cond, p, p_stride, cs, ncs, J, nullptr, &x);
}
extern "C" int rmsb_jacp_batch_if(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                                  int64_t p_stride, const int32_t* cs, int32_t ncs, const double* ifs, int64_t if_stride, double* Jp) {
    rmsb_extras x = extras_of_interfaces(ifs, if_stride);
    return eval_host(3, h, kind, n, y, cond, p, p_stride, cs, ncs, Jp, nullptr, &x);
}
extern "C" int rmsb_rates_batch(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, const double* p,
                                int64_t p_stride, double* kf, double* krev) {
    return eval_host(2, h, kind, n, y, cond, p, p_stride, nullptr, 0, kf, krev);
}

extern "C" int rmsb_kfs_batch(const rmsb_mech* h, int kind, int64_t n, const double* y, const double* cond, double* kf) {
    if (kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) return fail(RMSB_ERR_UNSUPPORTED, "rmsb_kfs_batch: single-phase domains only");
    return eval_host(4, h, kind, n, y, cond, nullptr, 0, nullptr, 0, kf, nullptr);
}

extern "C" int rmsb_gibbs_batch(const rmsb_mech* h, int64_t n, const double* T, int scalar_path, double* G) {
    if (!h || n < 0 || (n > 0 && (!T || !G))) return fail(RMSB_ERR_ARG, "bad argument");
    if (n == 0) return RMSB_OK;
    DeviceGuard g(h->device);
    if (!g.ok) return fail(RMSB_ERR_CUDA, "cudaSetDevice failed");
    StreamGuard sg;
    CU(cudaStreamCreateWithFlags(&sg.st, cudaStreamNonBlocking));
    DevBuf dT(sg.st), dG(sg.st);
    CU(dT.alloc(sizeof(double) * n)); CU(dG.alloc(sizeof(double) * n * h->dm.Ns));
    H2D(dT.p, T, sizeof(double) * n);
    int64_t blocks = (n + 3) / 4;
    if (blocks > 65535) blocks = 65535;
    gibbs_kernel<<<(int)blocks, 128, (size_t)MECH_DOUBLES * sizeof(double), sg.st>>>(h->dm, n, dT.as<double>(), scalar_path, dG.as<double>());
    ++g_launches;
    CU(cudaGetLastError());
    D2H(G, dG.p, sizeof(double) * n * h->dm.Ns);
    CU(cudaStreamSynchronize(sg.st));
    return RMSB_OK;
}

static int solve_host_impl(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                           int64_t p_stride, const int32_t* cs, int32_t ncs, const rmsb_extras* xtr,
                           const double* t_end, double rtol, double atol,
                           int32_t nsave, const double* tsave, double* ysave, double* yend, double* idt, int32_t* status,
                           rmsb_stats* stats, const rmsb_solve_opts* opts) {
    int rc = check_common(h, kind, n, y0, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    Extras ex;
    if ((rc = check_extras(h, kind, xtr, &ex)) != RMSB_OK) return rc;
    const double* ifs = ex.ifs;
    const int64_t if_stride = ex.if_stride;
    if ((rc = check_const_species(h, cs, ncs)) != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    if (!t_end || !status) return fail(RMSB_ERR_ARG, "t_end and status are required");
    if (nsave > 0 && (!tsave || !ysave)) return fail(RMSB_ERR_ARG, "tsave/ysave required when nsave > 0");
    DeviceGuard g(h->device);
    if (!g.ok) return fail(RMSB_ERR_CUDA, "cudaSetDevice failed");
    tune_pool(h->device);
    StreamGuard sg;
    CU(cudaStreamCreateWithFlags(&sg.st, cudaStreamNonBlocking));
    const int Ny = domain_ny(h, kind), nc = domain_ncond(kind), Np = h->dm.Ns + h->dm.Nr;
    DevBuf dy(sg.st), dc(sg.st), dp(sg.st), dcs(sg.st), dte(sg.st), dts(sg.st), dys(sg.st), dye(sg.st), didt(sg.st), dst(sg.st),
        dstats(sg.st), dwork(sg.st);
    CU(dy.alloc(sizeof(double) * n * Ny)); H2D(dy.p, y0, sizeof(double) * n * Ny);
    CU(dc.alloc(sizeof(double) * n * nc)); H2D(dc.p, cond, sizeof(double) * n * nc);
    if (p) { size_t np = p_stride ? (size_t)n * Np : (size_t)Np; CU(dp.alloc(sizeof(double) * np)); H2D(dp.p, p, sizeof(double) * np); }
    if (cs && ncs > 0) { CU(dcs.alloc(sizeof(int32_t) * ncs)); H2D(dcs.p, cs, sizeof(int32_t) * ncs); }
    CU(dte.alloc(sizeof(double) * n)); H2D(dte.p, t_end, sizeof(double) * n);
    if (nsave > 0) {
        CU(dts.alloc(sizeof(double) * nsave)); H2D(dts.p, tsave, sizeof(double) * nsave);
        CU(dys.alloc(sizeof(double) * n * nsave * Ny));
    }
    if (yend) CU(dye.alloc(sizeof(double) * n * Ny));
    if (idt) CU(didt.alloc(sizeof(double) * n));
    CU(dst.alloc(sizeof(int32_t) * n));
    if (stats) CU(dstats.alloc(sizeof(rmsb_stats) * n));
    int64_t wbytes = rmsb_solve_work_bytes(h, kind, n);
    if (wbytes < 0) return RMSB_ERR_UNSUPPORTED;
    CU(dwork.alloc((size_t)wbytes));
    DevBuf dif(sg.st), dpr(sg.st);
    if (ifs) {
        const size_t nif = (size_t)(if_stride ? n : 1) * (7 + 2 * (size_t)h->dm.Ns);
        CU(dif.alloc(sizeof(double) * nif)); H2D(dif.p, ifs, sizeof(double) * nif);
    }
    if (ex.prof) {
        const size_t np = (size_t)(ex.prof_stride ? n : 1) * (size_t)ex.prof_len;
        CU(dpr.alloc(sizeof(double) * np)); H2D(dpr.p, ex.prof, sizeof(double) * np);
    }
    EventPair ev;
    CU(ev.start(sg.st));
    rc = solve_dev_impl(false, h, kind, n, dy.as<double>(), dc.as<double>(), p ? dp.as<double>() : nullptr, p_stride,
                        (cs && ncs > 0) ? dcs.as<int32_t>() : nullptr, ncs, dte.as<double>(), rtol, atol, nsave,
                        nsave > 0 ? dts.as<double>() : nullptr, nsave > 0 ? dys.as<double>() : nullptr, nullptr,
                        yend ? dye.as<double>() : nullptr, nullptr, idt ? didt.as<double>() : nullptr, dst.as<int32_t>(),
                        stats ? dstats.as<rmsb_stats>() : nullptr, opts, dwork.p, sg.st, ifs ? dif.as<double>() : nullptr, if_stride,
                        ex.prof ? dpr.as<double>() : nullptr, ex.prof_stride);
    if (rc != RMSB_OK) return rc;
    // du at the saved states (opts->dysave): one dydt evaluation of the n*nsave saved states, each with its reactor's constants
    double* dysave = (opts && opts->struct_size == (int32_t)sizeof(rmsb_solve_opts)) ? opts->dysave : nullptr;
    if (dysave && ((ifs && if_stride != 0) || ex.prof)) return fail(RMSB_ERR_UNSUPPORTED, "dysave with per-reactor boundary interfaces or time profiles is not built");
    DevBuf dce(sg.st), dpe(sg.st), ddy(sg.st);
    if (dysave && nsave > 0) {
        const int64_t ne = n * nsave;
        CU(dce.alloc(sizeof(double) * ne * nc)); CU(ddy.alloc(sizeof(double) * ne * Ny));
        expand_rows_kernel<<<(int)std::min<int64_t>((ne * nc + 255) / 256, 65535), 256, 0, sg.st>>>(dc.as<double>(), dce.as<double>(), n, nsave, nc);
        ++g_launches;
        const double* pe = p ? dp.as<double>() : nullptr;
        if (p && p_stride) {
            CU(dpe.alloc(sizeof(double) * ne * Np));
            expand_rows_kernel<<<(int)std::min<int64_t>((ne * Np + 255) / 256, 65535), 256, 0, sg.st>>>(dp.as<double>(), dpe.as<double>(), n, nsave, Np);
            ++g_launches;
            pe = dpe.as<double>();
        }
        CU(cudaGetLastError());
        BatchArgs be = make_batch(h, kind, ne, dys.as<double>(), dce.as<double>(), pe, p_stride, (cs && ncs > 0) ? dcs.as<int32_t>() : nullptr, ncs,
                                  ddy.as<double>(), nullptr);
        if (ifs) { be.flow = dif.as<double>(); be.flow_stride = 0; }
        rc = launch_eval<0>(h, be, sg.st);
        if (rc != RMSB_OK) return rc;
    }
    CU(ev.stop(sg.st));
    if (dysave && nsave > 0) D2H(dysave, ddy.p, sizeof(double) * n * nsave * Ny);
    if (nsave > 0) D2H(ysave, dys.p, sizeof(double) * n * nsave * Ny);
    if (yend) D2H(yend, dye.p, sizeof(double) * n * Ny);
    if (idt) D2H(idt, didt.p, sizeof(double) * n);
    D2H(status, dst.p, sizeof(int32_t) * n);
    if (stats) D2H(stats, dstats.p, sizeof(rmsb_stats) * n);
    CU(cudaStreamSynchronize(sg.st));
    ev.read();
    return RMSB_OK;
}

extern "C" int rmsb_solve_batch(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                                int64_t p_stride, const int32_t* cs, int32_t ncs, const double* t_end, double rtol, double atol,
                                int32_t nsave, const double* tsave, double* ysave, double* yend, double* idt, int32_t* status,
                                rmsb_stats* stats, const rmsb_solve_opts* opts) {
    return solve_host_impl(h, kind, n, y0, cond, p, p_stride, cs, ncs, nullptr, t_end, rtol, atol, nsave, tsave, ysave, yend, idt, status, stats, opts);
}

extern "C" int rmsb_solve_batch_if(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                                   int64_t p_stride, const int32_t* cs, int32_t ncs, const double* ifs, int64_t if_stride,
                                   const double* t_end, double rtol, double atol,
                                   int32_t nsave, const double* tsave, double* ysave, double* yend, double* idt, int32_t* status,
                                   rmsb_stats* stats, const rmsb_solve_opts* opts) {
    rmsb_extras x = extras_of_interfaces(ifs, if_stride);
    return solve_host_impl(h, kind, n, y0, cond, p, p_stride, cs, ncs, &x, t_end, rtol, atol, nsave, tsave, ysave, yend, idt, status, stats, opts);
}

extern "C" int rmsb_solve_batch_ex(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                                   int64_t p_stride, const int32_t* cs, int32_t ncs, const rmsb_extras* extras,
                                   const double* t_end, double rtol, double atol,
                                   int32_t nsave, const double* tsave, double* ysave, double* yend, double* idt, int32_t* status,
                                   rmsb_stats* stats, const rmsb_solve_opts* opts) {
    return solve_host_impl(h, kind, n, y0, cond, p, p_stride, cs, ncs, extras, t_end, rtol, atol, nsave, tsave, ysave, yend, idt, status, stats, opts);
}

extern "C" int rmsb_solve_batch_multi(const rmsb_mech* const* mechs, int32_t n_mech, int kind, int64_t n, const double* y0,
                                      const double* cond, const double* p, int64_t p_stride, const int32_t* cs, int32_t ncs,
                                      const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave, double* ysave,
                                      double* yend, double* idt, int32_t* status, rmsb_stats* stats, const rmsb_solve_opts* opts) {
    if (!mechs || n_mech < 1) return fail(RMSB_ERR_ARG, "rmsb_solve_batch_multi: at least one mechanism handle");
    for (int g = 0; g < n_mech; ++g) {
        if (!mechs[g]) return fail(RMSB_ERR_ARG, "rmsb_solve_batch_multi: null mechanism handle");
        if (mechs[g]->dm.Ns != mechs[0]->dm.Ns || mechs[g]->dm.Nr != mechs[0]->dm.Nr) return fail(RMSB_ERR_ARG, "rmsb_solve_batch_multi: handles of different mechanisms");
    }
    int rc = check_common(mechs[0], kind, n, y0, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    if (!t_end || !status) return fail(RMSB_ERR_ARG, "t_end and status are required");
    const int Ny = domain_ny(mechs[0], kind), nc = domain_ncond(kind), Np = mechs[0]->dm.Ns + mechs[0]->dm.Nr;
    std::vector<int> rcs(n_mech, RMSB_OK);
    std::vector<std::string> errs(n_mech);
    std::vector<double> kms(n_mech, -1.0);
    std::vector<std::thread> th;
    for (int g = 0; g < n_mech; ++g) {
        const int64_t lo = n * g / n_mech, hi = n * (g + 1) / n_mech, m = hi - lo;
        if (m == 0) continue;
        th.emplace_back([=, &rcs, &errs, &kms]() {
            rmsb_solve_opts o;
            if (opts) { std::memset(&o, 0, sizeof(o)); std::memcpy(&o, opts, std::min<size_t>(sizeof(o), (size_t)std::max(0, opts->struct_size))); }
            if (opts && opts->struct_size == (int32_t)sizeof(rmsb_solve_opts) && o.dysave) o.dysave += (size_t)lo * nsave * Ny;
            rcs[g] = rmsb_solve_batch(mechs[g], kind, m, y0 + lo * Ny, cond + lo * nc, (p && p_stride) ? p + lo * Np : p, p_stride, cs, ncs,
                                      t_end + lo, rtol, atol, nsave, tsave, ysave ? ysave + (size_t)lo * nsave * Ny : nullptr,
                                      yend ? yend + lo * Ny : nullptr, idt ? idt + lo : nullptr, status + lo, stats ? stats + lo : nullptr,
                                      opts ? &o : nullptr);
            if (rcs[g] != RMSB_OK) errs[g] = rmsb_g_err;     // thread-local message of the worker
            kms[g] = g_last_kernel_ms;
        });
    }
    for (auto& t : th) t.join();
    double worst = -1.0;
    for (int g = 0; g < n_mech; ++g) worst = std::max(worst, kms[g]);
    g_last_kernel_ms = worst;                                 // the slowest device bounds the call
    for (int g = 0; g < n_mech; ++g) if (rcs[g] != RMSB_OK) return fail(rcs[g], "device block " + std::to_string(g) + ": " + errs[g]);
    return RMSB_OK;
}

static int sens_host_impl(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                          int64_t p_stride, const int32_t* cs, int32_t ncs, const double* t_end, double rtol, double atol,
                          int32_t nsave, const double* tsave, double* ysave, double* ssave, double* yend, double* send,
                          int32_t* status, rmsb_stats* stats, const rmsb_solve_opts* opts, const rmsb_extras* xtr) {
    int rc = check_common(h, kind, n, y0, cond, p_stride, p);
    if (rc != RMSB_OK) return rc;
    if (n == 0) return RMSB_OK;
    if (!t_end || !status) return fail(RMSB_ERR_ARG, "t_end and status are required");
    if (nsave > 0 && !tsave) return fail(RMSB_ERR_ARG, "tsave required when nsave > 0");
    DeviceGuard g(h->device);
    if (!g.ok) return fail(RMSB_ERR_CUDA, "cudaSetDevice failed");
    tune_pool(h->device);
    StreamGuard sg;
    CU(cudaStreamCreateWithFlags(&sg.st, cudaStreamNonBlocking));
    const int Ny = domain_ny(h, kind), nc = domain_ncond(kind), Np = h->dm.Ns + h->dm.Nr;
    const size_t sper = (size_t)Ny * Np;
    int64_t wb = rmsb_sens_work_bytes(h, kind, n);
    if (wb < 0) return RMSB_ERR_UNSUPPORTED;
    DevBuf dy(sg.st), dc(sg.st), dp(sg.st), dcs(sg.st), dte(sg.st), dts(sg.st), dys(sg.st), dss(sg.st), dye(sg.st), dse(sg.st),
        dst(sg.st), dstats(sg.st), dwork(sg.st);
    CU(dy.alloc(sizeof(double) * n * Ny)); H2D(dy.p, y0, sizeof(double) * n * Ny);
    CU(dc.alloc(sizeof(double) * n * nc)); H2D(dc.p, cond, sizeof(double) * n * nc);
    if (p) { size_t np = p_stride ? (size_t)n * Np : (size_t)Np; CU(dp.alloc(sizeof(double) * np)); H2D(dp.p, p, sizeof(double) * np); }
    if (cs && ncs > 0) { CU(dcs.alloc(sizeof(int32_t) * ncs)); H2D(dcs.p, cs, sizeof(int32_t) * ncs); }
    CU(dte.alloc(sizeof(double) * n)); H2D(dte.p, t_end, sizeof(double) * n);
    if (nsave > 0) {
        CU(dts.alloc(sizeof(double) * nsave)); H2D(dts.p, tsave, sizeof(double) * nsave);
        if (ysave) CU(dys.alloc(sizeof(double) * n * nsave * Ny));
        if (ssave) CU(dss.alloc(sizeof(double) * n * nsave * sper));
    }
    if (yend) CU(dye.alloc(sizeof(double) * n * Ny));
    if (send) CU(dse.alloc(sizeof(double) * n * sper));
    CU(dst.alloc(sizeof(int32_t) * n));
    if (stats) CU(dstats.alloc(sizeof(rmsb_stats) * n));
    CU(dwork.alloc((size_t)wb));
    EventPair ev;
    CU(ev.start(sg.st));
    Extras ex;
    if ((rc = check_extras(h, kind, xtr, &ex)) != RMSB_OK) return rc;
    DevBuf dif(sg.st), dpr(sg.st);
    if (ex.ifs) {
        const size_t nif = (size_t)(ex.if_stride ? n : 1) * (7 + 2 * (size_t)h->dm.Ns);
        CU(dif.alloc(sizeof(double) * nif)); H2D(dif.p, ex.ifs, sizeof(double) * nif);
    }
    if (ex.prof) {
        const size_t np = (size_t)(ex.prof_stride ? n : 1) * (size_t)ex.prof_len;
        CU(dpr.alloc(sizeof(double) * np)); H2D(dpr.p, ex.prof, sizeof(double) * np);
    }
    rc = solve_dev_impl(true, h, kind, n, dy.as<double>(), dc.as<double>(), p ? dp.as<double>() : nullptr, p_stride,
                        (cs && ncs > 0) ? dcs.as<int32_t>() : nullptr, ncs, dte.as<double>(), rtol, atol, nsave,
                        nsave > 0 ? dts.as<double>() : nullptr, (nsave > 0 && ysave) ? dys.as<double>() : nullptr,
                        (nsave > 0 && ssave) ? dss.as<double>() : nullptr, yend ? dye.as<double>() : nullptr,
                        send ? dse.as<double>() : nullptr, nullptr, dst.as<int32_t>(), stats ? dstats.as<rmsb_stats>() : nullptr, opts,
                        dwork.p, sg.st, ex.ifs ? dif.as<double>() : nullptr, ex.if_stride, ex.prof ? dpr.as<double>() : nullptr, ex.prof_stride);
    if (rc != RMSB_OK) return rc;
    CU(ev.stop(sg.st));
    if (nsave > 0 && ysave) D2H(ysave, dys.p, sizeof(double) * n * nsave * Ny);
    if (nsave > 0 && ssave) D2H(ssave, dss.p, sizeof(double) * n * nsave * sper);
    if (yend) D2H(yend, dye.p, sizeof(double) * n * Ny);
    if (send) D2H(send, dse.p, sizeof(double) * n * sper);
    D2H(status, dst.p, sizeof(int32_t) * n);
    if (stats) D2H(stats, dstats.p, sizeof(rmsb_stats) * n);
    CU(cudaStreamSynchronize(sg.st));
    ev.read();
    return RMSB_OK;
}

extern "C" int rmsb_sens_batch(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                               int64_t p_stride, const int32_t* cs, int32_t ncs, const double* t_end, double rtol, double atol,
                               int32_t nsave, const double* tsave, double* ysave, double* ssave, double* yend, double* send,
                               int32_t* status, rmsb_stats* stats, const rmsb_solve_opts* opts) {
    return sens_host_impl(h, kind, n, y0, cond, p, p_stride, cs, ncs, t_end, rtol, atol, nsave, tsave, ysave, ssave, yend, send, status, stats,
                          opts, nullptr);
}
extern "C" int rmsb_sens_batch_ex(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                                  int64_t p_stride, const int32_t* cs, int32_t ncs, const rmsb_extras* extras,
                                  const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave, double* ysave,
                                  double* ssave, double* yend, double* send, int32_t* status, rmsb_stats* stats, const rmsb_solve_opts* opts) {
    return sens_host_impl(h, kind, n, y0, cond, p, p_stride, cs, ncs, t_end, rtol, atol, nsave, tsave, ysave, ssave, yend, send, status, stats,
                          opts, extras);
}
extern "C" int rmsb_sens_batch_if(const rmsb_mech* h, int kind, int64_t n, const double* y0, const double* cond, const double* p,
                                  int64_t p_stride, const int32_t* cs, int32_t ncs, const double* ifs, int64_t if_stride,
                                  const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave, double* ysave,
                                  double* ssave, double* yend, double* send, int32_t* status, rmsb_stats* stats, const rmsb_solve_opts* opts) {
    rmsb_extras x = extras_of_interfaces(ifs, if_stride);
    return sens_host_impl(h, kind, n, y0, cond, p, p_stride, cs, ncs, t_end, rtol, atol, nsave, tsave, ysave, ssave, yend, send, status, stats,
                          opts, &x);
}

extern "C" int rmsb_large_info(const rmsb_mech* h, int64_t* out) {
    if (!h || !out) return fail(RMSB_ERR_ARG, "null argument");
    if (!h->large) { for (int k = 0; k < 8; ++k) out[k] = 0; return RMSB_OK; }
    out[0] = h->lt.n; out[1] = h->lt.nnz; out[2] = h->lt.n_lvl_l; out[3] = h->lt.n_lvl_u; out[4] = h->large_max_row;
    out[5] = h->large_factor_flops; out[6] = h->lt.core; out[7] = (int64_t)large_scratch_doubles(h->lt.n, h->dm.Ns, h->dm.Nr, h->n_dt, h->n_je, h->lt.nnz, 0, 0, 0);
    return RMSB_OK;
}

extern "C" int rmsb_large_profile(const rmsb_mech* h, int reset, double* out) {
    if (!h || !out) return fail(RMSB_ERR_ARG, "null argument");
    DeviceGuard g(h->device);
    unsigned long long v[8];
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(v, g_large_prof, sizeof(v)));
    for (int k = 0; k < 8; ++k) out[k] = (double)v[k];
    if (reset) { unsigned long long z[8] = {0}; CU(cudaMemcpyToSymbol(g_large_prof, z, sizeof(z))); }
    return RMSB_OK;
}

extern "C" double rmsb_fp64_peak_tflops(int device, int iters) {
    if (rmsb_device_count() <= 0) { fail(RMSB_ERR_CUDA, "no CUDA device"); return -1.0; }
    DeviceGuard g(device);
    if (!g.ok) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    if (iters <= 0) iters = 1 << 16;
    int threads = 256, blocks = prop.multiProcessorCount * 8;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * threads * blocks) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads>>>(out, iters / 16);   // warm-up
    ++g_launches;
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(out, iters);
        ++g_launches;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    return best;
}
