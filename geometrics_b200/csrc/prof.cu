// prof.cu -- optional per-kernel device timing (thread-local): brackets every kernel launch the
// calling thread makes through the C ABI with CUDA events on the launch stream, so that a harness
// can report each kernel's average duration over its own timed region (bench.py "roofline").
// Off by default; when off a launch costs one thread-local load.
#include <vector>

#include "common.cuh"

namespace geom {
namespace prof {

struct Rec {
    cudaEvent_t a, b;
    int id;
};
struct State {
    bool on = false;
    std::vector<Rec> pool;  // events are created once and reused across begin/end rounds
    size_t used = 0;
};
static thread_local State tls;

static const char *const kNames[kKernelCount] = {
    "geom::chamfer::twoway_kernel", "geom::chamfer::refine_kernel", "geom::chamfer::oneway_kernel",
    "geom::tri::tri_kernel",        "geom::chamfer::loss_kernel",   "geom::chamfer::loss_bwd_kernel",
    "geom::sample::sample_fwd_kernel", "geom::sample::sample_bwd_kernel", "geom::sample::area_kernel",
    "geom::gcn::gemm_kernel",       "geom::gcn::spmm_kernel",       "geom::gcn::bwd_kernel",
    "geom::tri::closest_kernel",    "geom::tri::closest_bwd_kernel", "geom::misc", "geom::misc2",
    "geom::chamfer_tc::scan_kernel", "geom::chamfer_tc::pack_kernel", "geom::pool::fwd_kernel", "geom::pool::bwd_kernel",
    "geom::tri_tc::scan_kernel", "geom::tri_tc::prep_kernel", "geom::gcn::bwd_gemm_kernel", "geom::bn::fwd_kernel",
    "geom::bn::bwd_kernel", "geom::spare"};

Scope::Scope(cudaStream_t st, int id) : st_(st), rec_(-1)
{
    State &s = tls;
    if (!s.on) return;
    if (s.used == s.pool.size()) {
        Rec r;
        r.id = id;
        if (cudaEventCreate(&r.a) != cudaSuccess) return;
        if (cudaEventCreate(&r.b) != cudaSuccess) {
            cudaEventDestroy(r.a);
            return;
        }
        s.pool.push_back(r);
    }
    rec_ = (int)s.used++;
    s.pool[rec_].id = id;
    cudaEventRecord(s.pool[rec_].a, st_);
}

Scope::~Scope()
{
    if (rec_ >= 0) cudaEventRecord(tls.pool[rec_].b, st_);
}

}  // namespace prof
}  // namespace geom

using namespace geom::prof;

extern "C" int geom_prof_begin(void)
{
    tls.on = true;
    tls.used = 0;
    return GEOM_OK;
}

extern "C" int geom_prof_end(float *ms_per_kernel, int *launches_per_kernel)
{
    State &s = tls;
    s.on = false;
    for (int k = 0; k < GEOM_PROF_KERNELS; ++k) {
        if (ms_per_kernel) ms_per_kernel[k] = 0.f;
        if (launches_per_kernel) launches_per_kernel[k] = 0;
    }
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < s.used; ++i) {
        const Rec &r = s.pool[i];
        if ((e = cudaEventSynchronize(r.b)) != cudaSuccess) break;
        float ms = 0.f;
        if ((e = cudaEventElapsedTime(&ms, r.a, r.b)) != cudaSuccess) break;
        if (ms_per_kernel) ms_per_kernel[r.id] += ms;
        if (launches_per_kernel) launches_per_kernel[r.id] += 1;
    }
    s.used = 0;
    return geom::status_from_cuda(e);
}

extern "C" const char *geom_prof_kernel_name(int id)
{
    return (id >= 0 && id < kKernelCount) ? kNames[id] : "";
}
