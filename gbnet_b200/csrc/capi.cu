// C ABI of the B200-native OR-NOR sampler (include/gbnet_b200.h): the model layer.
//
// Replaces gbn::ModelBase / gbn::ModelORNOR (reference libgbnet/src/ModelBase.cpp:37-215,
// ModelORNOR.cpp:13-29): owns the compiled topology, every chain's state and counters in HBM,
// the sample / stop loop, Gelman-Rubin and the posterior summaries.  There is no CPU path: every
// compute entry point launches kernels or fails with GBN_ERR_CUDA.
#include "../../include/gbnet_b200.h"

#include <atomic>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "topology.hpp"
#include "kernels.cuh"
#include "device_math.cuh"

using namespace gbn;

// ------------------------------------------------------------------------------ errors, signal
static thread_local std::string g_error;

static int fail(gbn_status s, const std::string &msg)
{
    g_error = msg;
    return (int) s;
}

extern "C" const char *gbn_last_error(void) { return g_error.c_str(); }
extern "C" int gbn_abi_version(void) { return GBN_ABI_VERSION; }

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(GBN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

// ModelBase.cpp:3-33: a process-global flag polled between batches of sweeps
static volatile std::sig_atomic_t g_signal = 0;
static volatile std::sig_atomic_t g_caught = 0;        // a real SIGINT arrived (as opposed to gbn_set_signal)
static void sigint_handler(int signum) { g_signal = signum; g_caught = 1; }
extern "C" void gbn_set_signal(int signum) { g_signal = signum; }
extern "C" void gbn_install_sigint_handler(void) { std::signal(SIGINT, sigint_handler); }

// The reference installs its SIGINT handler in the ModelBase ctor (ModelBase.cpp:40) and keeps it for the
// life of the process, which takes KeyboardInterrupt away from Python.  Here the handler is installed for
// the duration of a sampling call only: Ctrl-C stops the loop after the current batch (the reference's poll
// point, ModelBase.cpp:121), the previous handler is put back, and the signal is raised again for it - so
// the caller's own handler (Python: KeyboardInterrupt) still runs once the call has returned.
namespace {
struct SigintScope {
    void (*prev)(int);
    SigintScope() : prev(std::signal(SIGINT, sigint_handler)) {}
    ~SigintScope()
    {
        std::signal(SIGINT, prev == SIG_ERR ? SIG_DFL : prev);
        const bool again = g_caught != 0 && prev != SIG_ERR && prev != SIG_IGN && prev != SIG_DFL && prev != sigint_handler;
        g_caught = 0;
        g_signal = 0;                  // clearSignal(), ModelBase.cpp:126
        if (again) std::raise(SIGINT);
    }
};
}  // namespace

extern "C" int gbn_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------ NCCL (dlopen)
// Only the moment sums cross ranks; NCCL is resolved at run time so the library
// loads on machines without it.
namespace {
struct NcclApi {
    struct Uid { char b[128]; };         // ncclUniqueId (passed by value)
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Uid, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string &why)
    {
        if (lib) return true;
        const char *names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { why = "cannot dlopen libnccl.so.2"; return false; }
        GetUniqueId = (decltype(GetUniqueId)) dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank)) dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce)) dlsym(lib, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy)) dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString)) dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { why = "NCCL symbols missing"; lib = nullptr; return false; }
        return true;
    }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_UINT64 = 5, NCCL_SUM = 0;
}  // namespace

// ------------------------------------------------------------------------------ model
struct gbn_model {
    gbn_params p{};
    Topology tp;
    int device = 0;
    int kernel = GBN_KERNEL_STRICT;
    uint32_t n_datasets = 1, n_graphs = 0, n_graphs_local = 0, chain_offset = 0, n_chains = 0;
    DevNet net{};
    DevChains ch{};
    std::vector<void *> owned;          // every cudaMalloc of this model
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
    uint64_t sweeps_total = 0;          // Philox counter word 1
    uint32_t stat_n = 0;                // sweeps since burn_stats = RVNode::chain_length
    uint64_t launches = 0;
    double last_ms = 0.;
    // device buffers
    uint8_t *d_y = nullptr, *d_uf = nullptr; double *d_yprob = nullptr, *d_stab = nullptr;
    double *d_inj_flat = nullptr, *d_inj_beta = nullptr, *d_inj_exp = nullptr;
    double *d_sums = nullptr, *d_psums = nullptr, *d_dev = nullptr, *d_pdev = nullptr, *d_gr = nullptr, *d_max = nullptr;
    double *d_tsum = nullptr, *d_tdev = nullptr;      // T nodes: the two small buffers that cross the ranks
    unsigned long long *d_isums = nullptr;
    double *d_scratch = nullptr; size_t scratch_bytes = 0;
    bool moments_valid = false;
    bool sim = false;                   // simulation mode (empty evidence)
    // fast kernel: the sweep works on packed views of S and counts into per-batch 8-bit counters
    uint32_t *h_flags = nullptr;        // host-mapped flag word the fast kernels raise (DevChains::flags)
    bool small = false;                 // the persistent shared-memory kernel (sweep_small.cu) sweeps this model
    bool s_bytes_valid = true;          // ch.S (bytes) agrees with ch.Sw (2-bit words)
    uint32_t pending = 0;               // sweeps counted in ch.dC and not yet folded into ch.cS
    bool cs_zero = false;               // fast path: the S totals (ch.cS) are logically zero and their memory is not written yet:
                                        // the first fold overwrites instead of adding, the R-hat check does not read them
    // per-kernel timing (gbn_model_set_phase_timing)
    bool phase_timing = false;
    std::vector<cudaEvent_t> phase_ev;  // [group][3 n + 1]
    uint32_t phase_pending = 0;         // sweeps whose events are recorded but not yet read
    uint32_t phase_groups = 1, phase_stride = 0;
    // chain groups on separate streams (fast path): one group's X kernel overlaps another's S kernel
    std::vector<cudaStream_t> gstream, xstream;     // per group: T/S stream, high-priority X stream
    std::vector<cudaEvent_t> gdone, gevx, gevs;
    cudaEvent_t gfork = nullptr;
    uint32_t stream_groups = 0;         // 0 = default (GBN_STREAM_GROUPS or 4)
    double phase_ms[3] = { 0., 0., 0. };
    uint64_t phase_launches = 0;        // sweeps accumulated in phase_ms
    // comm
    void *comm = nullptr; int rank = 0, world = 1;
    // evidence kept for re-initialisation
    std::vector<uint8_t> h_y; std::vector<double> h_yprob;
};

namespace {

template <class T>
int dev_alloc(gbn_model *m, T **out, size_t n)
{
    void *p = nullptr;
    size_t bytes = sizeof(T) * (n ? n : 1);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(GBN_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + " B): " + cudaGetErrorString(e));
    m->owned.push_back(p);
    *out = (T *) p;
    return 0;
}

template <class T>
int dev_upload(gbn_model *m, T **out, const std::vector<T> &v)
{
    if (int rc = dev_alloc(m, out, v.size())) return rc;
    if (!v.empty()) CUDA_TRY(cudaMemcpy(*out, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return 0;
}

// initial chain state (GraphORNOR.cpp:57-73, 223-224, 246): X drawn from its prior by the
// Multinomial ctor (Multinomial.cpp:42-49: inside the base-class ctor the blanket is prob[v]),
// T at the prior mean, S at mor + 1.  The two cumulative likelihoods exp(log(prob[v])) are evaluated
// here with the host libm so that the initial state is the oracle's bit for bit; the draw itself
// (keyed Philox, kind INIT) happens on the device.
// fast_views: the caller fills the packed views of the fast path from the network's initial words next
// (refresh_fast(m, true)), so the byte view of S is not written here (it is unpacked on demand)
int init_chains(gbn_model *m, bool fast_views = false)
{
    const double xprob[2][2] = { { 0.99, 0.01 }, { 0.10, 0.90 } };
    double cum[2][2];
    for (int a = 0; a < 2; a++) {
        cum[a][0] = std::exp(std::log(xprob[a][0]));
        cum[a][1] = cum[a][0] + std::exp(std::log(xprob[a][1]));
    }
    launch_init_chains(m->net, m->ch, cum, m->stream, fast_views);
    m->launches += 1;
    if (m->sim) {       // H starts at 1 ("no deg"), Y copies it (GraphORNOR.cpp:94-96, YDataNode.cpp:20-25)
        CUDA_TRY(cudaMemsetAsync(m->ch.H, 1, (size_t) m->n_chains * m->tp.nH, m->stream));
        CUDA_TRY(cudaMemsetAsync(m->ch.Y, 1, (size_t) m->n_chains * m->tp.nH, m->stream));
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int zero_stats(gbn_model *m)
{
    const size_t C = m->n_chains, nX = m->tp.nX, E = m->tp.Ep;
    CUDA_TRY(cudaMemsetAsync(m->ch.cX1, 0, sizeof(uint32_t) * C * nX, m->stream));
    CUDA_TRY(cudaMemsetAsync(m->ch.tMean, 0, sizeof(double) * C * nX, m->stream));
    CUDA_TRY(cudaMemsetAsync(m->ch.tM2, 0, sizeof(double) * C * nX, m->stream));
    // (the fast path leaves the 2.5 MB of S totals per chain untouched until a fold or a reader needs them)
    if (m->kernel == GBN_KERNEL_FAST && !m->sim && m->ch.dC) m->cs_zero = true;
    else CUDA_TRY(cudaMemsetAsync(m->ch.cS, 0, sizeof(uint2) * C * E, m->stream));
    if (m->ch.dC) CUDA_TRY(cudaMemsetAsync(m->ch.dC, 0, sizeof(uint2) * C * m->tp.nD, m->stream));
    m->pending = 0;
    if (m->sim) {
        CUDA_TRY(cudaMemsetAsync(m->ch.cH, 0, sizeof(uint2) * C * m->tp.nH, m->stream));
        CUDA_TRY(cudaMemsetAsync(m->ch.cY, 0, sizeof(uint2) * C * m->tp.nH, m->stream));
    }
    m->stat_n = 0;
    m->moments_valid = false;
    return 0;
}

// the fast kernel's packed views and caches from X, T and the byte view of S
// initial: the chains were just initialised (init_chains(m, true)): S words and the by-TF copy are broadcast from the
// network's initial words instead of being packed / gathered from the byte view, which is left stale
int refresh_fast(gbn_model *m, bool initial = false)
{
    if (m->kernel != GBN_KERNEL_FAST) return 0;
    const char *err = nullptr;
    int n = launch_fast_refresh(m->net, m->ch, m->stream, &err, initial);
    if (n < 0) return fail(GBN_ERR_CUDA, std::string("fast refresh: ") + err);
    m->launches += (uint64_t) n;
    m->s_bytes_valid = !initial;
    return 0;
}

// before anything reads ch.S: the fast sweep keeps S as 2-bit words
int ensure_s_bytes(gbn_model *m)
{
    if (m->kernel != GBN_KERNEL_FAST || m->s_bytes_valid) return 0;
    const char *err = nullptr;
    int n = launch_fast_unpack(m->net, m->ch, m->stream, &err);
    if (n < 0) return fail(GBN_ERR_CUDA, std::string("fast unpack: ") + err);
    m->launches += (uint64_t) n;
    m->s_bytes_valid = true;
    return 0;
}

// before anything reads ch.cS: the fast sweep counts S outcomes into per-batch 8-bit counters
int fold_counts(gbn_model *m)
{
    if (m->kernel != GBN_KERNEL_FAST) return 0;
    if (m->pending == 0) {
        if (m->cs_zero) {                                   // a reader of totals that were never written
            CUDA_TRY(cudaMemsetAsync(m->ch.cS, 0, sizeof(uint2) * (size_t) m->n_chains * m->tp.Ep, m->stream));
            m->cs_zero = false;
        }
        return 0;
    }
    const char *err = nullptr;
    int n = launch_fast_fold(m->net, m->ch, m->stream, &err, m->cs_zero);
    if (n < 0) return fail(GBN_ERR_CUDA, std::string("fast fold: ") + err);
    m->launches += (uint64_t) n;
    m->pending = 0;
    m->cs_zero = false;
    return 0;
}

int upload_evidence(gbn_model *m, const gbn_evidence *ev)
{
    const Topology &tp = m->tp;
    const bool empty = !ev || ev->n_evid == 0 || ev->n_datasets == 0;
    if (empty != m->sim) return fail(GBN_ERR_INVALID, "evidence cannot switch a model between inference and simulation mode");
    if (!empty && ev->n_datasets != m->n_datasets) return fail(GBN_ERR_INVALID, "n_datasets differs from the model's");
    m->h_y.assign((size_t) m->n_datasets * tp.nH, 1);
    m->h_yprob.assign((size_t) m->n_datasets * 3, 0.);
    if (empty) { m->h_yprob[0] = 0.05; m->h_yprob[1] = 0.90; m->h_yprob[2] = 0.05; }   // GraphORNOR.h:45
    try {
        for (uint32_t d = 0; !empty && d < m->n_datasets; d++)
            compile_evidence(tp, ev->n_evid, ev->evid_trg, ev->evid_val + (size_t) d * ev->n_evid, m->p.comp_yprob != 0,
                             &m->h_y[(size_t) d * tp.nH], &m->h_yprob[(size_t) d * 3]);
    } catch (const std::exception &ex) {
        return fail(GBN_ERR_INVALID, ex.what());
    }
    {
        // tables of the fast S phase (sweep_fast.cu): per dataset, over n = number of parents with S != 1
        const uint32_t D = m->net.stab_D;
        const double ynoise[3][3] = { { 0.945, 0.050, 0.005 }, { 0.050, 0.900, 0.050 }, { 0.005, 0.050, 0.945 } };
        std::vector<double> tab((size_t) m->n_datasets * STAB_ROWS * D);
        for (uint32_t d = 0; d < m->n_datasets; d++) {
            double *t = &tab[(size_t) d * STAB_ROWS * D];
            const double *yp = &m->h_yprob[(size_t) d * 3];
            for (int cls = 0; cls < 2; cls++) {
                double z = cls ? m->net.z_n : m->net.z_y, v = 1.;
                for (uint32_t n = 0; n < D; n++) { t[(5 + cls) * D + n] = v; t[(3 + cls) * D + n] = 1. - v; v *= (1. - z); }
                v = 1.;
                for (uint32_t n = 0; n < D; n++) { t[(13 + cls) * D + n] = v; v *= z; }     // z^n: the factor product of n inactive parents
            }
            for (int y = 0; y < 3; y++) {
                const int cls = (y != 1) ? 0 : 1;
                const double K = yp[0] * ynoise[0][y] + yp[1] * ynoise[1][y] + yp[2] * ynoise[2][y];
                for (uint32_t n = 0; n < D; n++) {
                    const double omz = t[(3 + cls) * D + n];
                    t[y * D + n] = omz * ynoise[0][y] + t[(5 + cls) * D + n] * K;
                    t[(7 + y) * D + n] = omz * (ynoise[2][y] - ynoise[0][y]);
                    t[(10 + y) * D + n] = omz * (ynoise[1][y] - ynoise[2][y]);
                }
            }
        }
        CUDA_TRY(cudaMemcpy(m->d_stab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
        // k_fast_s2 drops the all-zero fallback of Multinomial.cpp:78-85: prior x likelihood must be positive for some
        // outcome.  That holds when every prior row has an entry that is not tiny and the likelihood tables are
        // positive: L = omz (a convex combination of YNOISE entries) + zc K with K = sum_h YPROB[h] YNOISE[h][y], so
        // YPROB >= 0 with a positive sum suffices (comp_yprob can make an entry negative, GraphORNOR.cpp:113-125,
        // when evidence outside the network is counted).  Anything else runs on k_fast_s, which keeps the fallback.
        uint32_t safe = 1;
        for (int r = 0; r < 3; r++) {
            bool pos = false;
            for (int v = 0; v < 3; v++) {
                const double q = m->net.sprob[r][v];
                if (!(q == 0. || (q >= 1e-200 && q <= 1e200))) safe = 0;         // negative, NaN, inf, or tiny
                if (q >= 1e-200) pos = true;
            }
            if (!pos) safe = 0;
        }
        for (uint32_t d = 0; d < m->n_datasets; d++) {
            const double *yp = &m->h_yprob[(size_t) d * 3];
            if (!(yp[0] >= 0. && yp[1] >= 0. && yp[2] >= 0. && yp[0] + yp[1] + yp[2] >= 1e-100 && yp[0] + yp[1] + yp[2] <= 1e100)) safe = 0;
        }
        m->net.sprior_safe = safe;
    }
    {
        // (dataset, TF) pairs whose children's likelihood product could underflow a double: a target likelihood is a
        // mixture sum_h P(h) YNOISE[h][y] >= min_h YNOISE[h][y]; only for these does the X kernel carry prod L
        // (the prior fallback of Multinomial.cpp:78-85 needs its magnitude, sweep_fast.cu k_fast_x)
        const double ynoise[3][3] = { { 0.945, 0.050, 0.005 }, { 0.050, 0.900, 0.050 }, { 0.005, 0.050, 0.945 } };
        double lmin[3];
        for (int y = 0; y < 3; y++) lmin[y] = std::log(std::min(ynoise[0][y], std::min(ynoise[1][y], ynoise[2][y])));
        // counted on the device (one warp per (dataset, TF) over the by-TF child list): on the host this loop over every
        // edge was a third of a set_evidence call
        CUDA_TRY(cudaMemcpy(m->d_y, m->h_y.data(), m->h_y.size(), cudaMemcpyHostToDevice));
        launch_tf_underflow_flags(m->net, m->d_uf, lmin, -600., m->stream);       // -600 in natural logs ~ 2^-865
        m->launches += 1;
    }
    CUDA_TRY(cudaMemcpy(m->d_yprob, m->h_yprob.data(), sizeof(double) * m->h_yprob.size(), cudaMemcpyHostToDevice));
    return 0;
}

int ensure_scratch(gbn_model *m, size_t bytes)
{
    if (m->scratch_bytes >= bytes) return 0;
    if (m->d_scratch) {
        cudaFree(m->d_scratch);
        for (auto &p : m->owned) if (p == m->d_scratch) p = nullptr;
    }
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    m->owned.push_back(p);
    m->d_scratch = (double *) p;
    m->scratch_bytes = bytes;
    return 0;
}

int use_device(const gbn_model *m)
{
    CUDA_TRY(cudaSetDevice(m->device));
    return 0;
}

int allreduce(gbn_model *m, void *buf, size_t n, int op, int dtype = NCCL_FLOAT64)
{
    if (!m->comm || m->world <= 1) return 0;
    int rc = g_nccl.AllReduce(buf, buf, n, dtype, op, m->comm, m->stream);
    if (rc != 0) return fail(GBN_ERR_COMM, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return 0;
}

// after the stream has been synchronised: did a fast kernel meet a state it cannot represent?
// T = 1.0 exactly with X = 1 makes the edge factor (1 - t x) z zero; the fast path divides it out of the
// cached products (leave-one-out), so its reciprocal must exist.  The reference's arithmetic has no such
// limit: the strict kernel (kernel = GBN_KERNEL_STRICT) follows it.  Reaching the value needs prev + dx to
// round to 1.0 exactly and a Beta prior with b <= 1 (t_hmargin < 1), else its density there is 0 and the
// proposal is rejected (Beta.cpp:66-72).
int check_flags(gbn_model *m)
{
    if (m->h_flags && (*m->h_flags & 1u))
        return fail(GBN_ERR_UNSUPPORTED, "a T node reached exactly 1.0: the fast kernels cannot represent the zero edge factor; "
                                         "rebuild the model with kernel = GBN_KERNEL_STRICT");
    return 0;
}

// passes 1 and 2 of moments.cu with the all-reduce between them; results stay on the device
int compute_moments(gbn_model *m)
{
    if (m->moments_valid) return 0;
    const double N = (double) m->stat_n;
    // counts of the discrete nodes overflow the exact integer path only beyond M * N^2 >= 2^63
    if ((double) m->n_graphs * N * N >= 9.0e18) return fail(GBN_ERR_UNSUPPORTED, "chain length too large for the exact moment sums");
    // the fast path's per-batch byte counters are read beside the totals, not folded into them: the fold (a rewrite
    // of every total) waits until a byte could overflow or something reads the totals themselves
    const bool pend = m->kernel == GBN_KERNEL_FAST && m->pending > 0;
    // one u64 all-reduce of the stored counts' sums, two small fp64 ones for the T nodes (3 nX, then 2 nX)
    const size_t nt = (size_t) m->n_datasets * n_t_of(m->net);
    launch_count_sums(m->net, m->ch, m->d_isums, m->stream, pend, m->cs_zero);
    launch_moment_sums(m->net, m->ch, N, m->d_tsum, m->stream);
    if (int rc = allreduce(m, m->d_isums, (size_t) m->n_datasets * moment_isum_words(m->net), NCCL_SUM, NCCL_UINT64)) return rc;
    if (nt) if (int rc = allreduce(m, m->d_tsum, 3 * nt, NCCL_SUM)) return rc;
    launch_moment_devs(m->net, m->ch, N, m->d_tsum, m->d_tdev, m->stream);
    if (nt) if (int rc = allreduce(m, m->d_tdev, 2 * nt, NCCL_SUM)) return rc;
    launch_expand_counts(m->net, N, m->d_isums, m->d_tsum, m->d_tdev, m->d_sums, m->d_psums, m->d_dev, m->d_pdev, m->stream);
    m->launches += m->net.const_params ? 2 : 4;
    CUDA_TRY(cudaGetLastError());
    m->moments_valid = true;
    return 0;
}

int max_gelman_rubin(gbn_model *m, double *out)
{
    if (m->stat_n == 0) { *out = std::numeric_limits<double>::infinity(); return 0; }   // ModelBase.cpp:96-101
    if (int rc = compute_moments(m)) return rc;
    CUDA_TRY(cudaMemsetAsync(m->d_max, 0, sizeof(double) * m->n_datasets, m->stream));
    launch_gelman_rubin(m->net, (double) m->stat_n, m->d_sums, m->d_dev, UINT32_MAX, nullptr, m->d_max, m->stream);
    m->launches += 1;
    std::vector<double> h(m->n_datasets);
    CUDA_TRY(cudaMemcpyAsync(h.data(), m->d_max, sizeof(double) * m->n_datasets, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (int rc = check_flags(m)) return rc;
    double mx = 0.;
    for (double v : h) mx = std::fmax(mx, v);
    *out = mx;
    return 0;
}

// n sweeps of every chain on the fast path (n <= FAST_MAX_PENDING - pending): the chains are independent, so
// they are split into groups, one stream pair each, joined at the end.
// Measured on B200 (profiles/r1_summary.md): 3 groups for >= 192 chains (2: -4 %, 4: -1 %; from 5
// groups on the streams exceed the default 8 hardware queues and serialise); fewer chains are
// latency-bound and only pay for the extra launches
int fast_batch(gbn_model *m, uint32_t n, uint32_t sweep0, uint32_t stat0, const char **err)
{
    static const uint32_t env_groups = env_u32("GBN_STREAM_GROUPS", 0);
    uint32_t G = m->stream_groups ? m->stream_groups
               : (env_groups ? env_groups : (m->n_chains >= 192 ? 3 : (m->n_chains >= 96 ? 2 : 1)));
    // ... and a sweep of little work (C2 network x 512 chains: 0.19 ms) is bound by the host's launch rate when
    // three groups multiply the launches: one group below 1.5e7 edge-updates per sweep
    if (!m->stream_groups && !env_groups && (uint64_t) m->n_chains * m->tp.E < 15000000ull) G = 1;
    const uint32_t xq = fast_x_bundle();
    G = std::min(G, (m->n_chains + xq - 1) / xq);       // no empty groups (boundaries fall on whole bundles)
    if (G < 1) G = 1;
    static const uint32_t prio_x = env_u32("GBN_PRIO_X", 1);        // 2 = off
    while (m->gstream.size() < G) {
        cudaStream_t s, sx; cudaEvent_t e, e1, e2;
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, lo));
        CUDA_TRY(cudaStreamCreateWithPriority(&sx, cudaStreamNonBlocking, hi));
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        m->gstream.push_back(s); m->xstream.push_back(sx); m->gdone.push_back(e);
        m->gevx.push_back(e1); m->gevs.push_back(e2);
    }
    if (!m->gfork) CUDA_TRY(cudaEventCreateWithFlags(&m->gfork, cudaEventDisableTiming));
    const uint32_t stride = 3 * n + 1;
    if (m->phase_timing) {
        while (m->phase_ev.size() < (size_t) G * stride) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreate(&e));
            m->phase_ev.push_back(e);
        }
        m->phase_pending = n; m->phase_groups = G; m->phase_stride = stride;
    }
    CUDA_TRY(cudaEventRecord(m->gfork, m->stream));
    int launched = 0;
    for (uint32_t g = 0; g < G; g++) {
        // group boundaries at whole bundles: the X kernel sweeps fast_x_bundle() chains per CTA
        const uint32_t nb = (m->n_chains + xq - 1) / xq;                 // bundles
        const uint32_t c0 = (uint32_t) ((uint64_t) nb * g / G) * xq;
        const uint32_t c1 = g + 1 == G ? m->n_chains : (uint32_t) ((uint64_t) nb * (g + 1) / G) * xq;
        cudaStream_t st = G == 1 ? m->stream : m->gstream[g];
        // the X kernels get their own (high-priority) stream: they overlap the other groups' S kernels and,
        // when the T nodes do not listen to their children, this group's own T draws
        const bool px = (G > 1 || fast_uses_x_stream(m->net)) && prio_x != 2 && !m->phase_timing;
        if (G > 1 || px) CUDA_TRY(cudaStreamWaitEvent(st, m->gfork, 0));
        if (px) CUDA_TRY(cudaStreamWaitEvent(m->xstream[g], m->gfork, 0));
        int l = launch_sweep_fast(m->net, m->ch, n, sweep0, stat0, st, err,
                                  m->phase_timing ? m->phase_ev.data() + (size_t) g * stride : nullptr, c0, c1 - c0,
                                  px ? m->xstream[g] : nullptr, px ? m->gevx[g] : nullptr, px ? m->gevs[g] : nullptr);
        if (l < 0) return -1;
        launched += l;
        if (G > 1) {
            CUDA_TRY(cudaEventRecord(m->gdone[g], st));
            CUDA_TRY(cudaStreamWaitEvent(m->stream, m->gdone[g], 0));
        }
    }
    if (fast_t_beside_x(m->net) && (n & 1)) std::swap(m->ch.T, m->ch.T2);
    m->pending += n;
    m->s_bytes_valid = false;
    return launched;
}

int sample_n(gbn_model *m, uint32_t n)
{
    if (n == 0) return 0;
    if ((uint64_t) m->stat_n + n > 0xffffffffull) return fail(GBN_ERR_INVALID, "statistics counters would overflow 32 bits");
    const char *err = nullptr;
    CUDA_TRY(cudaEventRecord(m->ev0, m->stream));
    int launched;
    if (m->sim) {
        launched = launch_sim_sweep(m->net, m->ch, n, (uint32_t) m->sweeps_total, m->stat_n, m->stream, &err);
    } else if (m->kernel == GBN_KERNEL_FAST) {
        // the S nodes count into 8-bit per-batch counters: fold them into the totals before one could overflow
        launched = 0;
        for (uint32_t done = 0; done < n; ) {
            if (m->pending + std::min(n - done, FAST_MAX_PENDING) > FAST_MAX_PENDING)
                if (int rc = fold_counts(m)) return rc;
            const uint32_t chunk = std::min(n - done, FAST_MAX_PENDING - m->pending);
            g_error.clear();
            int l;
            if (m->small) {
                l = launch_sweep_small(m->net, m->ch, chunk, (uint32_t) m->sweeps_total + done, m->stat_n + done, m->stream, &err);
                if (l >= 0) { m->pending += chunk; m->s_bytes_valid = false; }
            } else
                l = fast_batch(m, chunk, (uint32_t) m->sweeps_total + done, m->stat_n + done, &err);
            if (l < 0) { launched = -1; break; }
            launched += l;
            done += chunk;
            if (m->ch.inj_flat && done < n) m->ch.inj_pos += chunk;   // the last chunk's rows are consumed below
        }
        if (launched < 0 && !err) return (int) GBN_ERR_CUDA;        // a CUDA_TRY inside fast_batch set the message
    } else
        launched = launch_sweep_strict(m->net, m->ch, n, (uint32_t) m->sweeps_total, m->stat_n, m->stream, &err);
    if (launched < 0) return fail(GBN_ERR_CUDA, std::string("sweep launch: ") + err);
    CUDA_TRY(cudaEventRecord(m->ev1, m->stream));
    m->launches += (uint64_t) launched;
    m->sweeps_total += n;
    m->stat_n += n;
    m->moments_valid = false;
    if (m->ch.inj_flat) {
        m->ch.inj_pos += (m->kernel == GBN_KERNEL_FAST && !m->sim) ? (n - 1) % FAST_MAX_PENDING + 1 : n;
        if (m->ch.inj_pos >= m->ch.inj_sweeps) { m->ch.inj_flat = m->ch.inj_beta = m->ch.inj_exp = nullptr; m->ch.inj_sweeps = m->ch.inj_pos = 0; }
    }
    return 0;
}

}  // namespace

extern "C" void gbn_params_default(gbn_params *p)
{
    std::memset(p, 0, sizeof *p);
    const double sp[9] = { 0.40, 0.40, 0.20, 0.30, 0.40, 0.30, 0.20, 0.40, 0.40 };   // GraphORNOR.h:31-33
    std::memcpy(p->sprior, sp, sizeof sp);
    p->z_alpha = 25.; p->z_beta = 25.; p->z0_alpha = 25.; p->z0_beta = 25.;          // pyx:13-16, 20-23
    p->t_alpha = 2.; p->t_beta = 2.;
    p->n_graphs = 3;
    p->noise_listen_children = 1; p->comp_yprob = 0; p->const_params = 0;
    p->t_focus = 2.; p->t_lmargin = 2.; p->t_hmargin = 2.;
    p->zn_focus = 2.; p->zn_lmargin = 8.; p->zn_hmargin = 2.;
    p->seed = 0; p->device = 0; p->kernel = GBN_KERNEL_AUTO;
    p->chain_offset = 0; p->n_graphs_local = 0;
}

extern "C" void gbn_model_destroy(gbn_model *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
    if (m->h_flags) cudaFreeHost(m->h_flags);
    for (void *p : m->owned) if (p) cudaFree(p);
    for (cudaEvent_t e : m->phase_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : m->gdone) cudaEventDestroy(e);
    for (cudaStream_t st : m->gstream) cudaStreamDestroy(st);
    for (cudaStream_t st : m->xstream) cudaStreamDestroy(st);
    for (cudaEvent_t e : m->gevx) cudaEventDestroy(e);
    for (cudaEvent_t e : m->gevs) cudaEventDestroy(e);
    if (m->gfork) cudaEventDestroy(m->gfork);
    if (m->tev0) cudaEventDestroy(m->tev0);
    if (m->tev1) cudaEventDestroy(m->tev1);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

extern "C" int gbn_model_create(const gbn_network *netin, const gbn_evidence *ev, const gbn_params *p, gbn_model **out)
{
    if (!netin || !p || !out) return fail(GBN_ERR_INVALID, "null argument");
    *out = nullptr;
    if (p->n_graphs == 0) return fail(GBN_ERR_INVALID, "n_graphs must be positive");
    // empty evidence selects the reference's simulation mode (GraphORNOR.cpp:28)
    const bool sim = !ev || ev->n_evid == 0 || ev->n_datasets == 0;
    if (sim && ev && ev->n_datasets > 1) return fail(GBN_ERR_INVALID, "simulation mode takes no datasets");
    for (int r = 0; r < 3; r++)
        for (int v = 0; v < 3; v++)
            if (!(p->sprior[3 * r + v] >= 0.)) return fail(GBN_ERR_INVALID, "sprior entries must be non-negative");   // Multinomial.cpp:39
    if (!(p->z_alpha > 0.) || !(p->z_beta > 0.) || !(p->z0_alpha > 0.) || !(p->z0_beta > 0.))
        return fail(GBN_ERR_INVALID, "Beta parameters must be positive");                                            // Beta.cpp:34

    int ndev = gbn_device_count();
    if (ndev <= 0) return fail(GBN_ERR_CUDA, "no CUDA device: gbnet_b200 has no CPU path");
    if (p->device < 0 || p->device >= ndev) return fail(GBN_ERR_INVALID, "device ordinal out of range");

    gbn_model *m = new gbn_model();
    m->p = *p;
    m->device = p->device;
    auto bail = [&](int rc) { std::string keep = g_error; gbn_model_destroy(m); g_error = keep; return rc; };
    if (int rc = use_device(m)) return bail(rc);

    try {
        m->tp = compile_topology(netin->n_edge, netin->src, netin->trg, netin->mor, netin->n_active, netin->active_tf,
                                 p->t_focus, p->t_lmargin, p->t_hmargin);
    } catch (const std::exception &ex) {
        return bail(fail(GBN_ERR_INVALID, ex.what()));
    }
    const Topology &tp = m->tp;
    m->n_datasets = sim ? 1 : ev->n_datasets;
    m->sim = sim;
    m->n_graphs = p->n_graphs;
    m->n_graphs_local = p->n_graphs_local ? p->n_graphs_local : p->n_graphs;
    m->chain_offset = p->chain_offset;
    if (m->chain_offset + m->n_graphs_local > m->n_graphs) return bail(fail(GBN_ERR_INVALID, "chain shard exceeds n_graphs"));
    m->n_chains = m->n_datasets * m->n_graphs_local;
    if (m->n_chains > 65535) return bail(fail(GBN_ERR_INVALID, "more than 65535 chains on one device"));

    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess)
        return bail(fail(GBN_ERR_CUDA, "cannot create stream / events"));

    {
        void *hp = nullptr, *dp = nullptr;
        if (cudaHostAlloc(&hp, sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&dp, hp, 0) != cudaSuccess)
            return bail(fail(GBN_ERR_CUDA, "cannot allocate the host-mapped flag word"));
        m->h_flags = (uint32_t *) hp;
        *m->h_flags = 0;
        m->ch.flags = (volatile uint32_t *) dp;
    }
    // ---- network
    DevNet &n = m->net;
    n.nX = tp.nX; n.nH = tp.nH; n.E = tp.E; n.Eu = tp.Eu; n.max_indeg = tp.max_indeg; n.max_outdeg = tp.max_outdeg;
    n.Ep = tp.Ep; n.n_groups = tp.n_groups;
    {
        uint32_t *u32 = nullptr; uint8_t *u8 = nullptr; double *f64 = nullptr;
#define UP(field, vec, tmp) { if (int rc = dev_upload(m, &tmp, vec)) return bail(rc); n.field = tmp; }
        UP(deg, tp.deg, u32) UP(gbase, tp.gbase, u32) UP(ref_of_dt, tp.ref_of_dt, u32)
        UP(par_tf, tp.par_tf, u32) UP(par_edge, tp.par_edge, u32)
        UP(slot_of_edge, tp.slot_of_edge, u32) UP(slot_trg, tp.slot_trg, u32) UP(mor_idx, tp.mor_idx, u8)
        UP(tf_ptr, tp.tf_ptr, u32) UP(tf_child, tp.tf_child, u32) UP(tf_slot, tp.tf_slot, u32)
        UP(tfpos_of_slot, tp.tfpos_of_slot, u32) UP(par_pack, tp.par_pack, u32)
        UP(sbase, tp.sbase, u32) UP(dbase, tp.dbase, u32) UP(sw_grp, tp.sw_grp, u32) UP(dblk_grp, tp.dblk_grp, u32)
        UP(mor2, tp.mor2, u32)
        UP(fp_ptr, tp.fp_ptr, u32) UP(fp_child, tp.fp_child, u32) UP(fp_slot, tp.fp_slot, u32) UP(fp_abit, tp.fp_abit, u32)
        n.Efp = tp.Efp;
        { uint32_t *t4 = nullptr; if (int rc = dev_upload(m, &t4, tp.tfpos4)) return bail(rc); n.tfpos4 = reinterpret_cast<const uint4 *>(t4); }
        { uint16_t *u16 = nullptr; if (int rc = dev_upload(m, &u16, tp.par16)) return bail(rc); n.par16 = reinterpret_cast<const uint2 *>(u16); }
        n.nSw = tp.nSw; n.nD = tp.nD;
        {
            std::vector<uint32_t> rec(4 * (size_t) tp.n_groups), cmap(2 * (size_t) m->n_chains);
            for (uint32_t g = 0; g < tp.n_groups; g++) {
                rec[4 * g] = tp.gbase[g]; rec[4 * g + 1] = (tp.gbase[g + 1] - tp.gbase[g]) >> 5;
                rec[4 * g + 2] = tp.sbase[g]; rec[4 * g + 3] = tp.dbase[g];
            }
            for (uint32_t c = 0; c < m->n_chains; c++) {
                cmap[2 * c] = c / m->n_graphs_local;
                cmap[2 * c + 1] = (c / m->n_graphs_local) * m->n_graphs + m->chain_offset + c % m->n_graphs_local;
            }
            // the packed views of the initial state (every chain starts with S = mor + 1 on every edge)
            std::vector<uint32_t> swi(tp.nSw, 0x55555555u), sti((tp.Efp + 19) >> 4, 0x55555555u);
            for (uint32_t i = 0; i < tp.nSw; i++) {
                const uint32_t row = i >> 5, lane = i & 31, g = tp.sw_grp[row], jw = row - (tp.sbase[g] >> 5);
                const uint32_t gb = tp.gbase[g], W = (tp.gbase[g + 1] - gb) >> 5;
                uint32_t w = 0;
                for (uint32_t u = 0; u < 16; u++) {
                    const uint32_t j = 16 * jw + u;
                    w |= (j < W ? (uint32_t) tp.mor_idx[gb + lane + 32 * j] : 1u) << (2 * u);
                }
                swi[i] = w;
            }
            for (uint32_t pos = 0; pos < tp.Efp; pos++)
                if (tp.fp_slot[pos] != UINT32_MAX)
                    sti[pos >> 4] = (sti[pos >> 4] & ~(3u << (2 * (pos & 15)))) | ((uint32_t) tp.mor_idx[tp.fp_slot[pos]] << (2 * (pos & 15)));
            uint32_t *t = nullptr;
            if (int rc = dev_upload(m, &t, swi)) return bail(rc);
            n.sw_init = t;
            t = nullptr;
            if (int rc = dev_upload(m, &t, sti)) return bail(rc);
            n.stfp_init = t;
            t = nullptr;
            if (int rc = dev_upload(m, &t, rec)) return bail(rc);
            n.grec = reinterpret_cast<const uint4 *>(t);
            t = nullptr;
            if (int rc = dev_upload(m, &t, cmap)) return bail(rc);
            n.chain_map = reinterpret_cast<const uint2 *>(t);
        }
        UP(x_prior_active, tp.x_prior_active, u8)
        UP(t_a, tp.t_a, f64) UP(t_b, tp.t_b, f64) UP(t_mean, tp.t_mean, f64) UP(t_lnB, tp.t_lnB, f64)
#undef UP
    }
    const double xprob[2][2] = { { 0.99, 0.01 }, { 0.10, 0.90 } };                    // GraphORNOR.h:40-41
    const double ynoise[3][3] = { { 0.945, 0.050, 0.005 }, { 0.050, 0.900, 0.050 }, { 0.005, 0.050, 0.945 } };   // GraphORNOR.h:46-48
    std::memcpy(n.xprob, xprob, sizeof xprob);
    std::memcpy(n.ynoise, ynoise, sizeof ynoise);
    std::memcpy(n.sprob, p->sprior, sizeof(double) * 9);                              // GraphORNOR.cpp:25
    n.sprob[3][0] = 0.; n.sprob[3][1] = 1.; n.sprob[3][2] = 0.;                       // pad steps of the fast S kernel: never move
    n.z_y = p->z_alpha / (p->z_alpha + p->z_beta);                                    // GraphORNOR.cpp:155-164
    n.z_n = p->z0_alpha / (p->z0_alpha + p->z0_beta);
    n.rz_y = 1. / n.z_y; n.rz_n = 1. / n.z_n;
    n.n_datasets = m->n_datasets; n.n_graphs_local = m->n_graphs_local; n.chain_offset = m->chain_offset; n.n_graphs = m->n_graphs;
    n.sim = sim ? 1 : 0;
    n.listen = p->noise_listen_children ? 1 : 0;
    n.const_params = p->const_params ? 1 : 0;
    n.seed = p->seed;
    if (int rc = dev_alloc(m, &m->d_y, (size_t) m->n_datasets * tp.nH)) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_yprob, (size_t) m->n_datasets * 3)) return bail(rc);
    n.stab_D = (tp.max_indeg + 2 + 1) & ~1u;     // even: the arrays behind the tables stay 16-byte aligned in shared memory
    n.groups_desc = 1;
    for (uint32_t g = 0; g + 1 < tp.n_groups; g++)
        if (tp.gbase[g + 2] - tp.gbase[g + 1] > tp.gbase[g + 1] - tp.gbase[g]) n.groups_desc = 0;
    if (int rc = dev_alloc(m, &m->d_stab, (size_t) m->n_datasets * STAB_ROWS * n.stab_D)) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_uf, (size_t) m->n_datasets * tp.nX)) return bail(rc);
    n.y = m->d_y; n.yprob = m->d_yprob; n.stab = m->d_stab; n.tf_uf = m->d_uf;
    if (int rc = upload_evidence(m, ev)) return bail(rc);

    // ---- kernel choice
    const char *why = nullptr;
    bool fast_ok = fast_supported(n, &why);
    if (!sim && p->kernel == GBN_KERNEL_FAST && !fast_ok)
        return bail(fail(GBN_ERR_UNSUPPORTED, std::string("fast kernel unavailable: ") + why));
    m->kernel = (p->kernel == GBN_KERNEL_STRICT) ? GBN_KERNEL_STRICT : (fast_ok ? GBN_KERNEL_FAST : GBN_KERNEL_STRICT);
    if (sim) m->kernel = GBN_KERNEL_STRICT;          // simulation has its own kernels (sim_kernels.cu), reference arithmetic
    {
        // the persistent shared-memory kernel wins while every chain has an SM to itself (one CTA per chain, its phases
        // bound by latency: C1 24 vs 33 us per sweep, C2 x 64 chains 51 vs 72); with several waves of chains the
        // multi-kernel path's throughput wins (C2 x 512: 184 vs 203 us).  GBN_SMALL=3 forces it for any chain count.
        const char *w2 = nullptr;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device) != cudaSuccess) sms = 0;
        m->small = m->kernel == GBN_KERNEL_FAST && small_supported(n, &w2) &&
                   (m->n_datasets * m->n_graphs_local <= (uint32_t) sms || env_u32("GBN_SMALL", 1) == 3);
    }
    if (!sim && m->kernel == GBN_KERNEL_STRICT && strict_smem_bytes(n) > 227 * 1024)
        return bail(fail(GBN_ERR_UNSUPPORTED, "strict kernel: n_tf too large for shared memory"));

    // ---- chains
    DevChains &ch = m->ch;
    const size_t C = m->n_chains, nX = tp.nX, Ep = tp.Ep, nH = tp.nH;
    ch.n_chains = m->n_chains;
    if (int rc = dev_alloc(m, &ch.X, C * nX)) return bail(rc);
    if (int rc = dev_alloc(m, &ch.T, C * nX)) return bail(rc);
    if (int rc = dev_alloc(m, &ch.T2, C * nX)) return bail(rc);
    if (int rc = dev_alloc(m, &ch.S, C * Ep)) return bail(rc);
    if (int rc = dev_alloc(m, &ch.cX1, C * nX)) return bail(rc);
    if (int rc = dev_alloc(m, &ch.tMean, C * nX)) return bail(rc);
    if (int rc = dev_alloc(m, &ch.tM2, C * nX)) return bail(rc);
    if (int rc = dev_alloc(m, &ch.cS, C * Ep)) return bail(rc);
    if (sim) {
        if (int rc = dev_alloc(m, &ch.H, C * nH)) return bail(rc);
        if (int rc = dev_alloc(m, &ch.Y, C * nH)) return bail(rc);
        if (int rc = dev_alloc(m, &ch.cH, C * nH)) return bail(rc);
        if (int rc = dev_alloc(m, &ch.cY, C * nH)) return bail(rc);
    }
    if (m->kernel == GBN_KERNEL_FAST) {
        const size_t xq = fast_x_bundle();
        const size_t n_tc2 = (C + xq - 1) / xq * xq * (nH + 1);                              // interleaved by chain bundle;
        if (int rc = dev_alloc(m, &ch.tc2, n_tc2)) return bail(rc);                          // entry nH of a bundle stays {0, 0}
        if (cudaMemset(ch.tc2, 0, sizeof(double2) * n_tc2) != cudaSuccess) return bail(fail(GBN_ERR_CUDA, "cudaMemset failed"));
        if (int rc = dev_alloc(m, &ch.tL, n_tc2)) return bail(rc);
        if (int rc = dev_alloc(m, &ch.tq, n_tc2)) return bail(rc);
        if (cudaMemset(ch.tq, 0, n_tc2) != cudaSuccess) return bail(fail(GBN_ERR_CUDA, "cudaMemset failed"));
        {
            std::vector<double> ones(n_tc2, 1.);                                             // entry nH of a bundle stays 1
            if (cudaMemcpy(ch.tL, ones.data(), sizeof(double) * n_tc2, cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(GBN_ERR_CUDA, "cudaMemcpy failed"));
        }
        n.x1 = fast_x1_supported(n) ? 1 : 0;
        if (n.x1) {                                                                          // rows of the single-chain X kernel; entries
            const size_t nw = C * n_w32(n), nq = C * n_q8(n);                                // nH .. of a row stay 0
            if (int rc = dev_alloc(m, &ch.w0f, nw)) return bail(rc);
            if (int rc = dev_alloc(m, &ch.w2f, nw)) return bail(rc);
            if (int rc = dev_alloc(m, &ch.tq1, nq)) return bail(rc);
            if (cudaMemset(ch.w0f, 0, sizeof(float) * nw) != cudaSuccess || cudaMemset(ch.w2f, 0, sizeof(float) * nw) != cudaSuccess ||
                cudaMemset(ch.tq1, 0, nq) != cudaSuccess) return bail(fail(GBN_ERR_CUDA, "cudaMemset failed"));
        }
        if (int rc = dev_alloc(m, &ch.Sw, C * tp.nSw)) return bail(rc);
        if (int rc = dev_alloc(m, &ch.Aw, C * tp.nSw)) return bail(rc);
        if (int rc = dev_alloc(m, &ch.dC, C * tp.nD)) return bail(rc);
        if (int rc = dev_alloc(m, &ch.Stfp, C * n_stf_words(n))) return bail(rc);
        if (int rc = dev_alloc(m, &ch.xU, C * ((nX + 3) & ~(size_t) 3))) return bail(rc);
        if (int rc = dev_alloc(m, &ch.xqp, C * nX)) return bail(rc);
        if (cudaMemset(ch.xqp, 0, sizeof(uint16_t) * C * nX) != cudaSuccess) return bail(fail(GBN_ERR_CUDA, "cudaMemset failed"));
        if (int rc = dev_alloc(m, &ch.FXp, C * nX)) return bail(rc);
    }
    const size_t ns = (size_t) m->n_datasets * n_stats_of(n);
    if (int rc = dev_alloc(m, &m->d_sums, 2 * ns)) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_psums, ns)) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_dev, ns)) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_pdev, ns)) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_isums, (size_t) m->n_datasets * moment_isum_words(n))) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_tsum, 3 * (size_t) m->n_datasets * n_t_of(n))) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_tdev, 2 * (size_t) m->n_datasets * n_t_of(n))) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_gr, (size_t) n_nodes_of(n))) return bail(rc);
    if (int rc = dev_alloc(m, &m->d_max, (size_t) m->n_datasets)) return bail(rc);

    {
        const bool fv = m->kernel == GBN_KERNEL_FAST;
        if (int rc = init_chains(m, fv)) return bail(rc);
        if (int rc = zero_stats(m)) return bail(rc);
        if (int rc = refresh_fast(m, fv)) return bail(rc);
    }
    if (cudaStreamSynchronize(m->stream) != cudaSuccess) return bail(fail(GBN_ERR_CUDA, "initialisation failed on the device"));
    *out = m;
    return 0;
}

extern "C" int gbn_model_dims(const gbn_model *m, gbn_dims *out)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    out->n_tf = m->tp.nX; out->n_trg = m->tp.nH; out->n_edge = m->tp.E;
    out->n_nodes = n_nodes_of(m->net);
    out->n_datasets = m->n_datasets; out->n_graphs = m->n_graphs;
    out->n_graphs_local = m->n_graphs_local; out->chain_offset = m->chain_offset;
    out->n_stats = n_stats_of(m->net);
    out->simulation = m->sim ? 1 : 0;
    return 0;
}

extern "C" int gbn_model_kernel(const gbn_model *m) { return m ? m->kernel : (int) GBN_ERR_INVALID; }
extern "C" int gbn_model_path(const gbn_model *m)
{
    if (!m) return (int) GBN_ERR_INVALID;
    return m->kernel == GBN_KERNEL_FAST && !m->sim ? (m->small ? 2 : 1) : 0;
}

extern "C" int gbn_model_tf_ids(const gbn_model *m, uint32_t *out)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    std::memcpy(out, m->tp.tf_uid.data(), sizeof(uint32_t) * m->tp.nX);
    return 0;
}

extern "C" int gbn_model_trg_ids(const gbn_model *m, uint32_t *out)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    std::memcpy(out, m->tp.trg_uid.data(), sizeof(uint32_t) * m->tp.nH);
    return 0;
}

// after a stream synchronize: fold the recorded per-kernel events into phase_ms
static int collect_phase_times(gbn_model *m)
{
    for (uint32_t g = 0; g < m->phase_groups; g++)
        for (uint32_t sw = 0; sw < m->phase_pending; sw++)
            for (int ph = 0; ph < 3; ph++) {
                float ms = 0.f;
                const size_t o = (size_t) g * m->phase_stride + 3 * sw + ph;
                CUDA_TRY(cudaEventElapsedTime(&ms, m->phase_ev[o], m->phase_ev[o + 1]));
                m->phase_ms[ph] += ms;
            }
    m->phase_launches += (uint64_t) m->phase_pending * m->phase_groups;   // launches of each kernel
    m->phase_pending = 0;
    return 0;
}

extern "C" int gbn_model_sample_n(gbn_model *m, uint32_t n)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (n == 0) { m->last_ms = 0.; return 0; }
    if (int rc = use_device(m)) return rc;
    if (int rc = sample_n(m, n)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (int rc = check_flags(m)) return rc;
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
    m->last_ms = ms;
    return collect_phase_times(m);
}

extern "C" int gbn_model_set_phase_timing(gbn_model *m, int on)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    m->phase_timing = on != 0;
    m->phase_ms[0] = m->phase_ms[1] = m->phase_ms[2] = 0.;
    m->phase_launches = 0;
    m->phase_pending = 0;
    return 0;
}

extern "C" int gbn_model_set_stream_groups(gbn_model *m, uint32_t groups)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (groups > 64) return fail(GBN_ERR_INVALID, "at most 64 stream groups");
    m->stream_groups = groups;
    return 0;
}

extern "C" int gbn_model_phase_ms(const gbn_model *m, double ms_out[3], uint64_t *sweeps)
{
    if (!m || !ms_out || !sweeps) return fail(GBN_ERR_INVALID, "null argument");
    for (int i = 0; i < 3; i++) ms_out[i] = m->phase_ms[i];
    *sweeps = m->phase_launches;
    return 0;
}

// ModelBase::sample (ModelBase.cpp:114-129).  The for-loop's increment expression clips dN
// BEFORE adding it to n, so the batch that just ran is never clipped: sample(20, 30) draws 30.
// want_rhat = false skips the Gelman-Rubin evaluation after a batch when it cannot end the call (N <= dN:
// one batch whatever its value) and does not wait for the device; the caller synchronises.
static int sample_loop(gbn_model *m, uint32_t N, uint32_t dN, bool want_rhat, double *total_ms)
{
    uint32_t n;
    double gr = std::numeric_limits<double>::infinity();
    const bool lazy = !want_rhat && N <= dN;
    for (n = 0; n < N && gr > 1.10 && g_signal == 0; ) {
        if (int rc = sample_n(m, dN)) return rc;
        if (!lazy) {
            CUDA_TRY(cudaStreamSynchronize(m->stream));
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
            *total_ms += ms;
            if (int rc = collect_phase_times(m)) return rc;
        }
        dN = std::min(dN, N - n);
        n += dN;
        if (!lazy) if (int rc = max_gelman_rubin(m, &gr)) return rc;
        if (dN == 0) break;        // N - n reached 0: the for-condition ends the loop
    }
    return 0;
}

extern "C" int gbn_model_sample(gbn_model *m, uint32_t N, uint32_t dN)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (int rc = use_device(m)) return rc;
    m->last_ms = 0.;
    if (N == 0) return 0;
    if (dN == 0) return fail(GBN_ERR_INVALID, "deltaN must be positive (the reference's loop would never end)");
    SigintScope sigint;
    double total_ms = 0.;
    if (int rc = sample_loop(m, N, dN, true, &total_ms)) return rc;
    m->last_ms = total_ms;
    return 0;
}

// gbnet/utils.py:93-135 (sample_model) as one call
extern "C" int gbn_model_run_protocol(gbn_model *m, uint32_t burn_its, uint32_t max_its, uint32_t N, uint32_t dN,
                                      uint32_t *its_out, double *rhat_out)
{
    if (!m || !its_out || !rhat_out) return fail(GBN_ERR_INVALID, "null argument");
    if (N == 0 || dN == 0) return fail(GBN_ERR_INVALID, "N and deltaN must be positive");
    if (int rc = use_device(m)) return rc;
    SigintScope sigint;
    auto t0 = std::chrono::steady_clock::now();
    uint32_t it = 0;
    double ms = 0., rhat = std::numeric_limits<double>::infinity();
    *its_out = 0; *rhat_out = rhat;
    for (uint32_t i = 0; i < burn_its && g_signal == 0; i++) {                  // utils.py:94-96
        it += 1;
        if (int rc = sample_loop(m, N, dN, false, &ms)) return rc;
    }
    if (int rc = zero_stats(m)) return rc;                                      // utils.py:105: model.burn_stats()
    while (rhat > 1.1 && g_signal == 0) {                                       // utils.py:123-135
        it += 1;
        if (int rc = sample_loop(m, N, dN, true, &ms)) return rc;
        if (int rc = max_gelman_rubin(m, &rhat)) return rc;
        if (it >= max_its) break;
    }
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    m->phase_pending = 0;
    m->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *its_out = it; *rhat_out = rhat;
    return 0;
}

extern "C" int gbn_model_burn_stats(gbn_model *m)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (int rc = use_device(m)) return rc;
    if (int rc = zero_stats(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

extern "C" int gbn_model_chain_length(const gbn_model *m, double *out)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    *out = (double) m->stat_n;
    return 0;
}

extern "C" int gbn_model_max_gelman_rubin(gbn_model *m, double *out)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    if (int rc = use_device(m)) return rc;
    return max_gelman_rubin(m, out);
}

extern "C" int gbn_model_gelman_rubin(gbn_model *m, uint32_t dataset, double *out)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    if (dataset >= m->n_datasets) return fail(GBN_ERR_INVALID, "dataset out of range");
    if (int rc = use_device(m)) return rc;
    const uint32_t nn = n_nodes_of(m->net);
    if (m->stat_n == 0) { for (uint32_t i = 0; i < nn; i++) out[i] = std::numeric_limits<double>::infinity(); return 0; }
    if (int rc = compute_moments(m)) return rc;
    CUDA_TRY(cudaMemsetAsync(m->d_max, 0, sizeof(double) * m->n_datasets, m->stream));
    launch_gelman_rubin(m->net, (double) m->stat_n, m->d_sums, m->d_dev, dataset, m->d_gr, m->d_max, m->stream);
    m->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(out, m->d_gr, sizeof(double) * nn, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

// ModelBase::get_posterior_means / sdevs (ModelBase.cpp:160-209): gsl_rstat over the per-chain
// means of chains 1..M-1
static int posterior(gbn_model *m, uint32_t dataset, char kind, double *out, uint32_t *n_out, bool want_sd)
{
    if (!m || !n_out) return fail(GBN_ERR_INVALID, "null argument");
    if (dataset >= m->n_datasets) return fail(GBN_ERR_INVALID, "dataset out of range");
    const uint32_t nX = m->tp.nX, nH = m->tp.nH, E = m->tp.E, nT = m->net.const_params ? 0 : nX;
    uint32_t off, cnt;
    int hy = -1;                              // simulation mode: 0 = H, 1 = Y (interleaved per target)
    if (m->sim) {
        switch (kind) {                       // X and S are not random nodes there (GraphORNOR.cpp:58-63, 247-248)
        case 'H': off = 0; cnt = 3 * nH; hy = 0; break;
        case 'Y': off = 0; cnt = 3 * nH; hy = 1; break;
        case 'T': off = 6 * nH; cnt = nT; break;
        default: *n_out = 0; return 0;
        }
    } else switch (kind) {
    case 'X': off = 0; cnt = 2 * nX; break;
    case 'T': off = 2 * nX; cnt = nT; break;
    case 'S': off = 2 * nX + nT; cnt = 3 * E; break;
    default: *n_out = 0; return 0;            // no node carries that name: empty result, as in the reference
    }
    *n_out = cnt;
    if (!out || cnt == 0) return 0;
    if (int rc = use_device(m)) return rc;
    const double M = (double) m->n_graphs;
    if (m->stat_n == 0) {
        // every per-chain mean is gsl_rstat_mean of an empty accumulator = 0
        for (uint32_t i = 0; i < cnt; i++) out[i] = 0.;
        return 0;
    }
    if (int rc = compute_moments(m)) return rc;
    const size_t base = (size_t) dataset * n_stats_of(m->net) + off;
    std::vector<double> h(hy >= 0 ? 2 * cnt : cnt);
    CUDA_TRY(cudaMemcpyAsync(h.data(), (want_sd ? m->d_pdev : m->d_psums) + base, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (hy >= 0)                              // pick H_i (or Y_i) out of H_0 Y_0 H_1 Y_1 ...
        for (uint32_t i = 0; i < nH; i++)
            for (uint32_t j = 0; j < 3; j++) h[3 * i + j] = h[6 * i + 3 * hy + j];
    const double n = M - 1.;
    for (uint32_t i = 0; i < cnt; i++) {
        if (want_sd) out[i] = n > 1. ? std::sqrt(h[i] / (n - 1.)) : 0.;
        else out[i] = n > 0. ? h[i] / n : 0.;
    }
    return 0;
}

extern "C" int gbn_model_posterior_means(gbn_model *m, uint32_t dataset, char kind, double *out, uint32_t *n)
{ return posterior(m, dataset, kind, out, n, false); }
extern "C" int gbn_model_posterior_sdevs(gbn_model *m, uint32_t dataset, char kind, double *out, uint32_t *n)
{ return posterior(m, dataset, kind, out, n, true); }

// ModelBase::print_stats (ModelBase.cpp:144-158): "<id>_<stat>  mean(sd)" over chains 1..M-1
extern "C" int gbn_model_print_stats(gbn_model *m)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (m->sim) return fail(GBN_ERR_UNSUPPORTED, "print_stats is not available in simulation mode");
    const char kinds[3] = { 'X', 'T', 'S' };
    const uint32_t per[3] = { 2, 1, 3 };
    for (int t = 0; t < 3; t++) {
        uint32_t cnt = 0;
        if (int rc = posterior(m, 0, kinds[t], nullptr, &cnt, false)) return rc;
        if (cnt == 0) continue;
        std::vector<double> mu(cnt), sd(cnt);
        if (int rc = posterior(m, 0, kinds[t], mu.data(), &cnt, false)) return rc;
        if (int rc = posterior(m, 0, kinds[t], sd.data(), &cnt, true)) return rc;
        for (uint32_t i = 0; i < cnt; i++) {
            uint32_t node = i / per[t], j = i % per[t];
            if (kinds[t] == 'S') {
                uint32_t q = m->tp.slot_of_edge[node];
                std::printf("S_%u-->%u_%u\t%.4f(%.4f)\n", m->tp.tf_uid[m->tp.par_tf[q]], m->tp.trg_uid[m->tp.ref_of_dt[m->tp.slot_trg[q]]], j, mu[i], sd[i]);
            } else {
                std::printf("%c_%u_%u\t%.4f(%.4f)\n", kinds[t], m->tp.tf_uid[node], j, mu[i], sd[i]);
            }
        }
    }
    std::fflush(stdout);
    return 0;
}

extern "C" int gbn_model_set_evidence(gbn_model *m, const gbn_evidence *ev)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (int rc = use_device(m)) return rc;
    if (m->sim) return fail(GBN_ERR_UNSUPPORTED, "a simulation-mode model has no evidence to replace");
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    // GBN_TRACE=1: wall time of every stage on stderr (each stage is synchronised, so the total grows)
    static const bool trace = env_u32("GBN_TRACE", 0) == 1;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(m->stream);
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[gbnet_b200] set_evidence %-16s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    if (int rc = upload_evidence(m, ev)) return rc;
    lap("upload");
    m->sweeps_total = 0;
    m->ch.inj_flat = m->ch.inj_beta = m->ch.inj_exp = nullptr; m->ch.inj_sweeps = m->ch.inj_pos = 0;
    const bool fv = m->kernel == GBN_KERNEL_FAST;
    if (int rc = init_chains(m, fv)) return rc;
    lap("init_chains");
    if (int rc = zero_stats(m)) return rc;
    lap("zero_stats");
    if (int rc = refresh_fast(m, fv)) return rc;
    lap("refresh_fast");
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

// ------------------------------------------------------------------------------ host-side probe
extern "C" int gbn_topology_probe(const gbn_network *netin, const gbn_evidence *ev, int comp_yprob, gbn_topology_dims *dims,
                                  uint32_t *slot_of_edge, uint32_t *deg, uint32_t *ref_of_dt, uint32_t *gbase,
                                  uint32_t *tf_ptr, uint32_t *tf_child, uint32_t *tf_slot, uint8_t *y, double yprob[3])
{
    if (!netin) return fail(GBN_ERR_INVALID, "null argument");
    Topology tp;
    try {
        tp = compile_topology(netin->n_edge, netin->src, netin->trg, netin->mor, netin->n_active, netin->active_tf, 2., 2., 2.);
    } catch (const std::exception &ex) {
        return fail(GBN_ERR_INVALID, ex.what());
    }
    if (dims) {
        dims->n_tf = tp.nX; dims->n_trg = tp.nH; dims->n_edge = tp.E; dims->n_edge_unique = tp.Eu;
        dims->n_slots = tp.Ep; dims->n_groups = tp.n_groups; dims->max_indeg = tp.max_indeg; dims->max_outdeg = tp.max_outdeg;
    }
    auto put = [](uint32_t *dst, const std::vector<uint32_t> &v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(uint32_t) * v.size()); };
    put(slot_of_edge, tp.slot_of_edge); put(deg, tp.deg); put(ref_of_dt, tp.ref_of_dt); put(gbase, tp.gbase);
    put(tf_ptr, tp.tf_ptr); put(tf_slot, tp.tf_slot);
    if (tf_child) std::memcpy(tf_child, tp.tf_child.data(), sizeof(uint32_t) * tp.Eu);       // without the sentinel entry
    if (ev && ev->n_evid > 0 && ev->n_datasets > 0 && (y || yprob)) {
        std::vector<uint8_t> yy(tp.nH);
        double yp[3];
        try {
            compile_evidence(tp, ev->n_evid, ev->evid_trg, ev->evid_val, comp_yprob != 0, yy.data(), yp);
        } catch (const std::exception &ex) {
            return fail(GBN_ERR_INVALID, ex.what());
        }
        if (y) std::memcpy(y, yy.data(), tp.nH);
        if (yprob) { yprob[0] = yp[0]; yprob[1] = yp[1]; yprob[2] = yp[2]; }
    }
    return 0;
}

// ------------------------------------------------------------------------------ multi-GPU
extern "C" int gbn_comm_unique_id(void *id_out)
{
    std::string why;
    if (!g_nccl.load(why)) return fail(GBN_ERR_COMM, why);
    int rc = g_nccl.GetUniqueId(id_out);
    if (rc != 0) return fail(GBN_ERR_COMM, "ncclGetUniqueId failed");
    return 0;
}

extern "C" int gbn_model_attach_comm(gbn_model *m, int rank, int world, const void *id)
{
    if (!m || !id) return fail(GBN_ERR_INVALID, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(GBN_ERR_INVALID, "bad rank / world");
    if (int rc = use_device(m)) return rc;
    std::string why;
    if (!g_nccl.load(why)) return fail(GBN_ERR_COMM, why);
    NcclApi::Uid uid;
    std::memcpy(uid.b, id, 128);
    void *comm = nullptr;
    int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (rc != 0) return fail(GBN_ERR_COMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    m->comm = comm; m->rank = rank; m->world = world;
    m->moments_valid = false;
    return 0;
}

// ------------------------------------------------------------------------------ parity hooks
extern "C" int gbn_model_get_state(gbn_model *m, uint32_t chain, double *out)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    if (chain >= m->n_chains) return fail(GBN_ERR_INVALID, "chain out of range");
    if (int rc = use_device(m)) return rc;
    const uint32_t nX = m->tp.nX, E = m->tp.E, Ep = m->tp.Ep;
    std::vector<uint8_t> X(nX), S(Ep);
    std::vector<double> T(nX);
    if (int rc = ensure_s_bytes(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (m->sim) {           // random_nodes order of simulation mode: H_0 Y_0 H_1 Y_1 ... then T
        const uint32_t nH = m->tp.nH;
        std::vector<uint8_t> H(nH), Y(nH);
        CUDA_TRY(cudaMemcpy(H.data(), m->ch.H + (size_t) chain * nH, nH, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(Y.data(), m->ch.Y + (size_t) chain * nH, nH, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(T.data(), m->ch.T + (size_t) chain * nX, sizeof(double) * nX, cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < nH; r++) { out[2 * r] = H[m->tp.dt_of_ref[r]]; out[2 * r + 1] = Y[m->tp.dt_of_ref[r]]; }
        if (!m->net.const_params) for (uint32_t k = 0; k < nX; k++) out[2 * nH + k] = T[k];
        return 0;
    }
    CUDA_TRY(cudaMemcpy(X.data(), m->ch.X + (size_t) chain * nX, nX, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(T.data(), m->ch.T + (size_t) chain * nX, sizeof(double) * nX, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(S.data(), m->ch.S + (size_t) chain * Ep, Ep, cudaMemcpyDeviceToHost));
    uint32_t o = 0;
    for (uint32_t k = 0; k < nX; k++) out[o++] = X[k];
    if (!m->net.const_params) for (uint32_t k = 0; k < nX; k++) out[o++] = T[k];
    for (uint32_t e = 0; e < E; e++) out[o++] = S[m->tp.slot_of_edge[e]];
    return 0;
}

extern "C" int gbn_model_set_state(gbn_model *m, uint32_t chain, const double *in)
{
    if (!m || !in) return fail(GBN_ERR_INVALID, "null argument");
    if (chain >= m->n_chains) return fail(GBN_ERR_INVALID, "chain out of range");
    if (int rc = use_device(m)) return rc;
    const uint32_t nX = m->tp.nX, E = m->tp.E, Ep = m->tp.Ep;
    std::vector<uint8_t> X(nX), S(Ep, 1);             // pad slots stay at 1 (edge off)
    std::vector<double> T(nX);
    if (m->sim) {
        const uint32_t nH = m->tp.nH;
        std::vector<uint8_t> H(nH), Y(nH);
        for (uint32_t r = 0; r < nH; r++) {
            const double h = in[2 * r], y = in[2 * r + 1];
            if ((h != 0. && h != 1. && h != 2.) || (y != 0. && y != 1. && y != 2.)) return fail(GBN_ERR_INVALID, "H / Y values must be 0, 1 or 2");
            H[m->tp.dt_of_ref[r]] = (uint8_t) h; Y[m->tp.dt_of_ref[r]] = (uint8_t) y;
        }
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        CUDA_TRY(cudaMemcpy(m->ch.H + (size_t) chain * nH, H.data(), nH, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(m->ch.Y + (size_t) chain * nH, Y.data(), nH, cudaMemcpyHostToDevice));
        if (!m->net.const_params) {
            for (uint32_t k = 0; k < nX; k++) T[k] = in[2 * nH + k];
            CUDA_TRY(cudaMemcpy(m->ch.T + (size_t) chain * nX, T.data(), sizeof(double) * nX, cudaMemcpyHostToDevice));
        }
        return 0;
    }
    if (int rc = ensure_s_bytes(m)) return rc;          // the refresh below re-packs every chain from the byte view
    uint32_t o = 0;
    for (uint32_t k = 0; k < nX; k++) { double v = in[o++]; if (v != 0. && v != 1.) return fail(GBN_ERR_INVALID, "X values must be 0 or 1"); X[k] = (uint8_t) v; }
    if (!m->net.const_params) for (uint32_t k = 0; k < nX; k++) T[k] = in[o++];
    for (uint32_t e = 0; e < E; e++) { double v = in[o++]; if (v != 0. && v != 1. && v != 2.) return fail(GBN_ERR_INVALID, "S values must be 0, 1 or 2"); S[m->tp.slot_of_edge[e]] = (uint8_t) v; }
    if (m->kernel == GBN_KERNEL_FAST && !m->net.const_params)
        for (uint32_t k = 0; k < nX; k++)
            if (T[k] == 1.)                 // zero edge factor: see check_flags
                return fail(GBN_ERR_UNSUPPORTED, "T = 1.0 exactly: the fast kernels cannot represent the zero edge factor; "
                                                 "build the model with kernel = GBN_KERNEL_STRICT for this state");
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (!m->net.const_params)
        CUDA_TRY(cudaMemcpy(m->ch.T + (size_t) chain * nX, T.data(), sizeof(double) * nX, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->ch.X + (size_t) chain * nX, X.data(), nX, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->ch.S + (size_t) chain * Ep, S.data(), Ep, cudaMemcpyHostToDevice));
    if (int rc = refresh_fast(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

static int eval_to_host(gbn_model *m, uint32_t chain, double *out, size_t n, int which)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    if (chain >= m->n_chains) return fail(GBN_ERR_INVALID, "chain out of range");
    if (int rc = use_device(m)) return rc;
    if (int rc = ensure_scratch(m, sizeof(double) * n)) return rc;
    if (int rc = ensure_s_bytes(m)) return rc;
    if (m->sim && which != 0) return fail(GBN_ERR_UNSUPPORTED, "only eval_conditionals is available in simulation mode");
    if (m->sim) launch_sim_conditionals(m->net, m->ch, chain, m->d_scratch, m->stream);
    else if (which == 0) launch_conditionals(m->net, m->ch, chain, m->d_scratch, m->stream);
    else if (which == 1) launch_target_likelihoods(m->net, m->ch, chain, m->d_scratch, m->stream);
    else launch_joint_logprob(m->net, m->ch, chain, m->d_scratch, m->stream);
    m->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, m->d_scratch, sizeof(double) * n, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

extern "C" int gbn_model_eval_conditionals(gbn_model *m, uint32_t chain, double *out)
{ return eval_to_host(m, chain, out, m ? (size_t) 3 * n_nodes_of(m->net) : 0, 0); }
extern "C" int gbn_model_target_likelihoods(gbn_model *m, uint32_t chain, double *out)
{ return eval_to_host(m, chain, out, m ? (size_t) m->tp.nH : 0, 1); }
extern "C" int gbn_model_joint_logprob(gbn_model *m, uint32_t chain, double *out)
{ return eval_to_host(m, chain, out, 1, 2); }

extern "C" int gbn_model_inject_draws(gbn_model *m, uint32_t n_sweeps, const double *flat_u, const double *beta_v, const double *exp_v)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (int rc = use_device(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    for (double **p : { &m->d_inj_flat, &m->d_inj_beta, &m->d_inj_exp })
        if (*p) { for (auto &q : m->owned) if (q == *p) q = nullptr; cudaFree(*p); *p = nullptr; }
    m->ch.inj_flat = m->ch.inj_beta = m->ch.inj_exp = nullptr; m->ch.inj_sweeps = m->ch.inj_pos = 0;
    if (n_sweeps == 0) return 0;
    if (!flat_u || !beta_v || !exp_v) return fail(GBN_ERR_INVALID, "null draw arrays");
    const size_t C = m->n_chains, nX = m->tp.nX, E = m->tp.E;
    // per sweep: X then S in edge input order; in simulation mode H_0 Y_0 H_1 Y_1 ... (reference target order)
    const size_t nf = C * n_sweeps * (m->sim ? 2 * (size_t) m->tp.nH : nX + E), nb = C * n_sweeps * nX;
    if (int rc = dev_alloc(m, &m->d_inj_flat, nf)) return rc;
    if (int rc = dev_alloc(m, &m->d_inj_beta, nb)) return rc;
    if (int rc = dev_alloc(m, &m->d_inj_exp, nb)) return rc;
    CUDA_TRY(cudaMemcpy(m->d_inj_flat, flat_u, sizeof(double) * nf, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->d_inj_beta, beta_v, sizeof(double) * nb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->d_inj_exp, exp_v, sizeof(double) * nb, cudaMemcpyHostToDevice));
    m->ch.inj_flat = m->d_inj_flat; m->ch.inj_beta = m->d_inj_beta; m->ch.inj_exp = m->d_inj_exp;
    m->ch.inj_sweeps = n_sweeps; m->ch.inj_pos = 0;
    return 0;
}

extern "C" int gbn_model_export_draws(gbn_model *m, uint32_t chain, uint32_t sweep0, uint32_t n, double *flat_u, double *beta_v, double *exp_v)
{
    if (!m || !flat_u || !beta_v || !exp_v) return fail(GBN_ERR_INVALID, "null argument");
    if (chain >= m->n_chains) return fail(GBN_ERR_INVALID, "chain out of range");
    if (int rc = use_device(m)) return rc;
    const size_t nX = m->tp.nX, E = m->tp.E;
    const size_t nf = (size_t) n * (m->sim ? 2 * (size_t) m->tp.nH : nX + E), nb = (size_t) n * nX;
    if (int rc = ensure_scratch(m, sizeof(double) * (nf + 2 * nb))) return rc;
    double *df = m->d_scratch, *db = df + nf, *de = db + nb;
    if (m->sim) launch_sim_export_draws(m->net, chain_global_id(m->net, chain), sweep0, n, df, db, de, m->stream);
    else launch_export_draws(m->net, chain_global_id(m->net, chain), sweep0, n, df, db, de, m->stream);
    m->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(flat_u, df, sizeof(double) * nf, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(beta_v, db, sizeof(double) * nb, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(exp_v, de, sizeof(double) * nb, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

extern "C" int gbn_model_node_moments(gbn_model *m, uint32_t chain, double *mean, double *variance)
{
    if (!m || !mean || !variance) return fail(GBN_ERR_INVALID, "null argument");
    if (chain >= m->n_chains) return fail(GBN_ERR_INVALID, "chain out of range");
    if (int rc = use_device(m)) return rc;
    const size_t ns = n_stats_of(m->net);
    if (int rc = ensure_scratch(m, sizeof(double) * 2 * ns)) return rc;
    if (m->stat_n == 0) { for (size_t i = 0; i < ns; i++) { mean[i] = 0.; variance[i] = 0.; } return 0; }
    if (int rc = fold_counts(m)) return rc;
    launch_node_moments(m->net, m->ch, chain, (double) m->stat_n, m->d_scratch, m->d_scratch + ns, m->stream);
    m->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(mean, m->d_scratch, sizeof(double) * ns, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(variance, m->d_scratch + ns, sizeof(double) * ns, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

extern "C" int gbn_philox_kat(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    if (gbn_device_count() <= 0) return fail(GBN_ERR_CUDA, "no CUDA device: gbnet_b200 has no CPU path");
    CUDA_TRY(cudaSetDevice(device));
    uint32_t *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 16));
    launch_philox_kat(ctr, key, d, 0);
    cudaError_t e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(GBN_ERR_CUDA, cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------ checkpoint / resume
// The reference has no serialisation of chain state or statistics; this is the "next"
// row of §8(f): everything a bit-exact continuation needs - X, T, S, the counters and Welford pairs,
// the sweep / statistics counters - as one flat blob.  Layout: 4 x u64 header {magic, n_chains,
// sweeps_total, stat_n} then the arrays in the order below.
namespace {
constexpr uint64_t CKPT_MAGIC = 0x67626e5f636b7031ull;      // "gbn_ckp1"
struct CkptPart { void *ptr; size_t bytes; };
std::vector<CkptPart> ckpt_parts(gbn_model *m)
{
    const size_t C = m->n_chains, nX = m->tp.nX, Ep = m->tp.Ep;
    std::vector<CkptPart> parts = { { m->ch.X, C * nX }, { m->ch.T, sizeof(double) * C * nX }, { m->ch.S, C * Ep },
             { m->ch.cX1, sizeof(uint32_t) * C * nX }, { m->ch.tMean, sizeof(double) * C * nX },
             { m->ch.tM2, sizeof(double) * C * nX }, { m->ch.cS, sizeof(uint2) * C * Ep } };
    if (m->sim) {
        const size_t nH = m->tp.nH;
        parts.push_back({ m->ch.H, C * nH }); parts.push_back({ m->ch.Y, C * nH });
        parts.push_back({ m->ch.cH, sizeof(uint2) * C * nH }); parts.push_back({ m->ch.cY, sizeof(uint2) * C * nH });
    }
    return parts;
}
}  // namespace

extern "C" int gbn_model_checkpoint_size(gbn_model *m, size_t *bytes)
{
    if (!m || !bytes) return fail(GBN_ERR_INVALID, "null argument");
    size_t n = 4 * sizeof(uint64_t);
    for (auto &p : ckpt_parts(m)) n += p.bytes;
    *bytes = n;
    return 0;
}

extern "C" int gbn_model_checkpoint_save(gbn_model *m, void *out, size_t bytes)
{
    if (!m || !out) return fail(GBN_ERR_INVALID, "null argument");
    size_t need = 0;
    gbn_model_checkpoint_size(m, &need);
    if (bytes < need) return fail(GBN_ERR_INVALID, "checkpoint buffer too small");
    if (int rc = use_device(m)) return rc;
    if (int rc = ensure_s_bytes(m)) return rc;
    if (int rc = fold_counts(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    uint64_t hdr[4] = { CKPT_MAGIC, m->n_chains, m->sweeps_total, m->stat_n };
    std::memcpy(out, hdr, sizeof hdr);
    char *o = (char *) out + sizeof hdr;
    for (auto &p : ckpt_parts(m)) { CUDA_TRY(cudaMemcpy(o, p.ptr, p.bytes, cudaMemcpyDeviceToHost)); o += p.bytes; }
    return 0;
}

extern "C" int gbn_model_checkpoint_load(gbn_model *m, const void *in, size_t bytes)
{
    if (!m || !in) return fail(GBN_ERR_INVALID, "null argument");
    size_t need = 0;
    gbn_model_checkpoint_size(m, &need);
    uint64_t hdr[4];
    if (bytes < sizeof hdr) return fail(GBN_ERR_INVALID, "checkpoint truncated");
    std::memcpy(hdr, in, sizeof hdr);
    if (hdr[0] != CKPT_MAGIC || hdr[1] != m->n_chains || bytes != need)
        return fail(GBN_ERR_INVALID, "checkpoint does not belong to a model of this shape");
    if (int rc = use_device(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    const char *i = (const char *) in + sizeof hdr;
    for (auto &p : ckpt_parts(m)) { CUDA_TRY(cudaMemcpy(p.ptr, i, p.bytes, cudaMemcpyHostToDevice)); i += p.bytes; }
    m->sweeps_total = hdr[2];
    m->stat_n = (uint32_t) hdr[3];
    m->moments_valid = false;
    if (m->ch.dC) CUDA_TRY(cudaMemsetAsync(m->ch.dC, 0, sizeof(uint2) * (size_t) m->n_chains * m->tp.nD, m->stream));
    m->pending = 0;
    m->cs_zero = false;                                     // the totals came with the checkpoint
    m->ch.inj_flat = m->ch.inj_beta = m->ch.inj_exp = nullptr; m->ch.inj_sweeps = m->ch.inj_pos = 0;
    if (int rc = refresh_fast(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}

// ------------------------------------------------------------------------------ measurement
extern "C" int gbn_model_last_device_ms(const gbn_model *m, double *ms)
{
    if (!m || !ms) return fail(GBN_ERR_INVALID, "null argument");
    *ms = m->last_ms;
    return 0;
}

extern "C" int gbn_model_launch_count(const gbn_model *m, uint64_t *n)
{
    if (!m || !n) return fail(GBN_ERR_INVALID, "null argument");
    *n = m->launches;
    return 0;
}

extern "C" int gbn_model_sweep_count(const gbn_model *m, uint64_t *n)
{
    if (!m || !n) return fail(GBN_ERR_INVALID, "null argument");
    *n = m->sweeps_total;
    return 0;
}

// device time between two points of the caller's choosing, on the model's stream
extern "C" int gbn_model_timer_start(gbn_model *m)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (int rc = use_device(m)) return rc;
    if (!m->tev0) { CUDA_TRY(cudaEventCreate(&m->tev0)); CUDA_TRY(cudaEventCreate(&m->tev1)); }
    CUDA_TRY(cudaEventRecord(m->tev0, m->stream));
    return 0;
}

extern "C" int gbn_model_timer_stop(gbn_model *m, double *ms)
{
    if (!m || !ms) return fail(GBN_ERR_INVALID, "null argument");
    if (!m->tev0) return fail(GBN_ERR_INVALID, "timer not started");
    if (int rc = use_device(m)) return rc;
    CUDA_TRY(cudaEventRecord(m->tev1, m->stream));
    CUDA_TRY(cudaEventSynchronize(m->tev1));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, m->tev0, m->tev1));
    *ms = f;
    return 0;
}

// diagnostics of the speculative X windows: rounds, TFs committed, flips (since enabled)
extern "C" int gbn_model_debug_counters(gbn_model *m, int enable, uint64_t out[16])
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (int rc = use_device(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (out) {
        for (int i = 0; i < 16; i++) out[i] = 0;
        if (m->ch.dbg) CUDA_TRY(cudaMemcpy(out, m->ch.dbg, 128, cudaMemcpyDeviceToHost));
    }
    if (enable && !m->ch.dbg) {
        unsigned long long *p = nullptr;
        if (int rc = dev_alloc(m, &p, 16)) return rc;
        m->ch.dbg = p;
    }
    if (m->ch.dbg) CUDA_TRY(cudaMemset(m->ch.dbg, 0, 128));
    if (!enable) m->ch.dbg = nullptr;
    return 0;
}

extern "C" int gbn_model_synchronize(gbn_model *m)
{
    if (!m) return fail(GBN_ERR_INVALID, "null model");
    if (int rc = use_device(m)) return rc;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return 0;
}
