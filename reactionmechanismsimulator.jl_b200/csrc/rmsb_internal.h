// rmsb_internal.h -- declarations shared by the translation units of librmsb200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "../../include/rmsb200.h"
#include "rmsb_device.cuh"
#include "rmsb_tables.h"

// ------------------------------------------------------------------------------------------------
// error plumbing (defined in rmsb_api.cu)
// ------------------------------------------------------------------------------------------------
extern thread_local std::string rmsb_g_err;
extern std::atomic<long long> rmsb_g_launches;
int rmsb_fail(int code, const std::string& msg);
#define g_err rmsb_g_err
#define g_launches rmsb_g_launches
#define fail rmsb_fail

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(RMSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));        \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// cudaGetDeviceProperties takes milliseconds: query each device once (rmsb_api.cu)
cudaError_t rmsb_cached_props(int device, cudaDeviceProp* out);

// ------------------------------------------------------------------------------------------------
// mechanism handle
// ------------------------------------------------------------------------------------------------
struct rmsb_spec_state;            // rmsb_spec.cu: host tables + the run-time compiled kernels of a small all-Arrhenius mechanism
struct rmsb_mech {
    int device = 0;
    rmsb::DevMech dm{};
    std::vector<void*> allocs;
    int n_dt = 0, n_je = 0;
    // large-mechanism path (one block per reactor, sparse LU on a fixed pattern): built when Ns > 128
    bool large = false;
    rmsb::LargeTables lt{};        // y = ns            (ConstantTV, ConstantTAPhi)
    rmsb::LargeTables lt_tp{};     // y = [ns; V]       (ConstantTP): the same block bordered by a dense row and column
    rmsb::LargeTables lt_vk{};     // y = [ns; T; P|V]  (ConstantV, ConstantP): two dense border rows and columns
    int large_max_row = 0;
    long long large_factor_flops = 0;
    // small-mechanism path (group of 16 / 32 lanes per reactor, tables staged in shared memory): rmsb_small.cu
    rmsb::SmallTables st{};
    // mechanism-specialised evaluation kernels (generated source, NVRTC): rmsb_spec.cu; null when the mechanism does not qualify
    std::shared_ptr<rmsb_spec_state> spec;
};

template <class T>
static int rmsb_upload(rmsb_mech* h, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    size_t bytes = (v.empty() ? 1 : v.size()) * sizeof(T);
    CU(cudaMalloc(&p, bytes));
    h->allocs.push_back(p);
    if (!v.empty()) CU(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T*>(p);
    return RMSB_OK;
}

namespace rmsb { struct SolveArgs; struct BatchArgs; }

// rmsb_small.cu ------------------------------------------------------------------------------------
// Build the shared-memory table image of a small all-Arrhenius mechanism (st.ok = 0 when the mechanism does not qualify).
int rmsb_small_build(rmsb_mech* h, const std::vector<int>& rx /* [8][Nr] 0-based, -1 none */, const std::vector<int>& srptr,
                     const std::vector<int>& srrxn, const std::vector<int>& kcptr, const std::vector<int>& kcspc,
                     const std::vector<double>& kccoef, const std::vector<double>& nrp, const std::vector<int>& ktype,
                     const std::vector<int>& koff, const std::vector<double>& kpar, const std::vector<double>& fwd,
                     const std::vector<double>& rev, const std::vector<int>& nasa_nr, const std::vector<int>& nasa_nc,
                     const std::vector<double>& nasa_tmax, const std::vector<double>& nasa_coef, int maxr);
// true when (mechanism, domain, options) go through the group-per-reactor kernels
bool rmsb_small_eligible(const rmsb_mech* h, int kind, int Ny, int flags);
int rmsb_small_launch_solve(const rmsb_mech* h, const rmsb::SolveArgs& a, cudaStream_t st);
// extra bytes of work buffer (behind the 256-byte header) the small path uses to order the work queue; 0 when it does not
size_t rmsb_small_work_bytes(const rmsb_mech* h, int kind, int Ny, int64_t n);
// mode 0: dydt, 1: Jacobian (dense column-major Ny x Ny)
int rmsb_small_launch_eval(const rmsb_mech* h, int mode, const rmsb::BatchArgs& a, cudaStream_t st);

// rmsb_spec.cu -------------------------------------------------------------------------------------
// keep the host tables the source generator needs (no-op when the mechanism does not qualify)
int rmsb_spec_capture(rmsb_mech* h, const std::vector<int>& rx, const std::vector<int>& srptr, const std::vector<int>& srrxn, const std::vector<int>& kcptr,
                      const std::vector<int>& kcspc, const std::vector<double>& kccoef, const std::vector<double>& nrp, const std::vector<int>& ktype,
                      const std::vector<int>& koff, const std::vector<double>& kpar, const std::vector<double>& fwd, const std::vector<double>& rev,
                      const std::vector<int>& nasa_nr, const std::vector<int>& nasa_nc, const std::vector<double>& nasa_tmax, const std::vector<double>& nasa_coef,
                      int maxr, const std::vector<int>& dtrxn, const std::vector<int>& dtmult, const std::vector<int>& dtoth, const std::vector<int>& jeptr,
                      const std::vector<int>& jerow, const std::vector<int>& jecol, const std::vector<int>& jtidx);
int rmsb_spec_source(const rmsb_mech* h, int kind, std::string* out);
int rmsb_spec_ready(const rmsb_mech* h, int kind, std::string* why);
int rmsb_spec_launch_eval(const rmsb_mech* h, int mode, const rmsb::BatchArgs& a, cudaStream_t st);
void rmsb_spec_release(rmsb_mech* h);
