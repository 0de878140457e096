// Enrichment pre-screen counts for many contrasts at once.
//
// Replaces the per-contrast pandas group-bys of the reference's get_tests (gbnet/utils.py:10-35): for every
// source TF the number of its edges whose target is / is not differentially expressed and the 2x2 counts of
// edge sign against DE sign.  The edge list is sorted by source once on the host (a CSR by TF, shared by all
// contrasts); on the device one warp per (contrast, TF) walks the TF's edges - consecutive lanes read
// consecutive edges - and reduces six counters with shuffles.  The Fisher p-values themselves are evaluated by
// the caller from these counts (gbnet_b200/utils.py, hypergeometric survival function for all TFs at once).
#include "../../include/gbnet_b200.h"

#include <algorithm>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

namespace {

constexpr uint32_t NO_KEY = 0xffffffffu;

// edges of TF k: [ptr[k], ptr[k+1]); key = index into the evidence table (NO_KEY: the target has no evidence row)
__global__ void k_fisher_counts(uint32_t n_tf, uint32_t n_evid, const uint32_t *ptr, const uint32_t *key, const int8_t *type,
                                const int8_t *val /* [n_datasets][n_evid] */, uint32_t *counts /* [n_datasets][n_tf][6] */)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t d = blockIdx.y;
    if (k >= n_tf) return;
    const int8_t *v = val + (size_t) d * n_evid;
    uint32_t c[6] = { 0, 0, 0, 0, 0, 0 };
    for (uint32_t e = ptr[k] + lane; e < ptr[k + 1]; e += 32) {
        const uint32_t q = key[e];
        const int x = q == NO_KEY ? GBN_EVID_ABSENT : (int) v[q];
        const int t = type[e];
        c[0] += x != 0;                            // pandas: NaN != 0 holds, so a target without evidence counts here
        c[1] += x == 0;
        c[2] += t == 1 && x == 1;
        c[3] += t == 1 && x == -1;
        c[4] += t == -1 && x == 1;
        c[5] += t == -1 && x == -1;
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c[j] += __shfl_xor_sync(0xffffffffu, c[j], o);
    }
    if (lane == 0) {
        uint32_t *out = counts + ((size_t) d * n_tf + k) * 6;
#pragma unroll
        for (int j = 0; j < 6; j++) out[j] = c[j];
    }
}

thread_local std::string f_error;

}  // namespace

extern "C" const char *gbn_fisher_last_error(void) { return f_error.c_str(); }

extern "C" int gbn_fisher_counts(int device, const gbn_network *net, uint32_t n_datasets, uint32_t n_evid,
                                 const uint32_t *evid_trg, const int8_t *evid_val,
                                 uint32_t *n_tf_out, uint32_t *tf_ids_out, uint32_t *counts)
{
    auto fail = [](gbn_status s, const std::string &msg) { f_error = msg; return (int) s; };
    if (!net || !n_tf_out) return fail(GBN_ERR_INVALID, "null argument");
    const uint32_t E = net->n_edge;
    std::vector<uint32_t> tf(net->src, net->src + E);
    std::sort(tf.begin(), tf.end());
    tf.erase(std::unique(tf.begin(), tf.end()), tf.end());
    const uint32_t n_tf = (uint32_t) tf.size();
    *n_tf_out = n_tf;
    if (tf_ids_out) std::copy(tf.begin(), tf.end(), tf_ids_out);
    if (!counts) return 0;
    if (n_datasets == 0 || n_tf == 0) return 0;
    if (n_evid && (!evid_trg || !evid_val)) return fail(GBN_ERR_INVALID, "null evidence arrays");
    // CSR by TF over the edge rows (counting sort, edge order kept inside a TF)
    std::vector<uint32_t> ptr(n_tf + 1, 0), rank(E), key(E);
    std::vector<int8_t> type(E);
    for (uint32_t e = 0; e < E; e++) {
        rank[e] = (uint32_t) (std::lower_bound(tf.begin(), tf.end(), net->src[e]) - tf.begin());
        ptr[rank[e] + 1]++;
    }
    for (uint32_t k = 0; k < n_tf; k++) ptr[k + 1] += ptr[k];
    std::unordered_map<uint32_t, uint32_t> where;              // target id -> row of the evidence table (first one wins)
    where.reserve(2 * (size_t) n_evid);
    for (uint32_t i = 0; i < n_evid; i++) where.emplace(evid_trg[i], i);
    {
        std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
        for (uint32_t e = 0; e < E; e++) {
            const uint32_t p = fill[rank[e]]++;
            auto it = where.find(net->trg[e]);
            key[p] = it == where.end() ? NO_KEY : it->second;
            type[p] = (int8_t) net->mor[e];
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(GBN_ERR_CUDA, "no CUDA device: gbnet_b200 has no CPU path"); }
    if (device < 0 || device >= ndev) return fail(GBN_ERR_INVALID, "device ordinal out of range");
    cudaError_t err = cudaSetDevice(device);
    uint32_t *d_ptr = nullptr, *d_key = nullptr, *d_counts = nullptr;
    int8_t *d_type = nullptr, *d_val = nullptr;
    const size_t n_counts = (size_t) n_datasets * n_tf * 6, n_val = (size_t) n_datasets * n_evid;
    auto upload = [&](void **dst, const void *src, size_t bytes) {
        if (err == cudaSuccess) err = cudaMalloc(dst, bytes ? bytes : 1);
        if (err == cudaSuccess && bytes && src) err = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    upload((void **) &d_ptr, ptr.data(), sizeof(uint32_t) * ptr.size());
    upload((void **) &d_key, key.data(), sizeof(uint32_t) * E);
    upload((void **) &d_type, type.data(), E);
    upload((void **) &d_val, evid_val, n_val);
    upload((void **) &d_counts, nullptr, sizeof(uint32_t) * n_counts);
    if (err == cudaSuccess) {
        const uint32_t wpb = 8;
        k_fisher_counts<<<dim3((n_tf + wpb - 1) / wpb, n_datasets), 32 * wpb>>>(n_tf, n_evid, d_ptr, d_key, d_type, d_val, d_counts);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpy(counts, d_counts, sizeof(uint32_t) * n_counts, cudaMemcpyDeviceToHost);
    cudaFree(d_ptr); cudaFree(d_key); cudaFree(d_type); cudaFree(d_val); cudaFree(d_counts);
    if (err != cudaSuccess) return fail(GBN_ERR_CUDA, std::string("gbn_fisher_counts: ") + cudaGetErrorString(err));
    return 0;
}
