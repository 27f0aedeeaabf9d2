// Device top-words export (reference variational_bayes.py:326-341 export_beta): per topic,
// the words sorted by beta_kv = exp(E_log_eta[k,v] - logsumexp_v E_log_eta[k,:]), descending.  The reference
// argsorts K rows of V probabilities on the host; at V = 1M, K = 500 that is a 4 GB copy-back plus 500 full sorts.
// Here the (V, KP) table already on the device is cut into chunks of topics, every chunk is sorted by one
// segmented radix sort (CUB, pairs of log-probability and word id) and only the first `top` entries of every
// topic cross PCIe.  Not on the E-step hot path.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <string>

namespace pylda {

__global__ void k_topic_keys(const double* __restrict__ Elt, const double* __restrict__ lse, int k0, int nk, int V, int KP,
                             double* __restrict__ keys, int* __restrict__ vals) {
    const long long total = (long long)nk * V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i / V);
        const int v = (int)(i - (long long)kk * V);
        keys[i] = Elt[(size_t)v * KP + k0 + kk] - lse[k0 + kk];
        vals[i] = v;
    }
}

__global__ void k_topic_offsets(int nk, int V, long long* __restrict__ off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nk) off[i] = (long long)i * V;
}

__global__ void k_take_top(const double* __restrict__ keys, const int* __restrict__ vals, int nk, int V, int top,
                           double* __restrict__ prob, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nk * top) {
        const int kk = i / top, t = i - kk * top;
        prob[i] = exp(keys[(size_t)kk * V + t]);
        idx[i] = vals[(size_t)kk * V + t];
    }
}

#define TK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            *err = std::string(#call " failed: ") + cudaGetErrorString(e_);        \
            goto done;                                                             \
        }                                                                          \
    } while (0)

// idx_out / prob_out: host buffers of K * top entries.  Returns 0 on success.
int device_top_words(const double* Elt, const double* lse, int K, int V, int KP, int top, int32_t* idx_out,
                     double* prob_out, cudaStream_t s, std::string* err) {
    int rc = 1;
    // chunk of topics whose keys + values (in and out) stay below ~2 GB
    int chunk = (int)std::max<long long>(1, std::min<long long>(K, (2LL << 30) / (24LL * V)));
    double *keys_in = nullptr, *keys_out = nullptr, *prob = nullptr;
    int *vals_in = nullptr, *vals_out = nullptr, *idx = nullptr;
    long long* off = nullptr;
    void* temp = nullptr;
    size_t temp_bytes = 0;
    const size_t n = (size_t)chunk * V;
    TK(cudaMalloc((void**)&keys_in, n * sizeof(double)));
    TK(cudaMalloc((void**)&keys_out, n * sizeof(double)));
    TK(cudaMalloc((void**)&vals_in, n * sizeof(int)));
    TK(cudaMalloc((void**)&vals_out, n * sizeof(int)));
    TK(cudaMalloc((void**)&off, ((size_t)chunk + 1) * sizeof(long long)));
    TK(cudaMalloc((void**)&prob, (size_t)chunk * top * sizeof(double)));
    TK(cudaMalloc((void**)&idx, (size_t)chunk * top * sizeof(int)));
    TK(cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, temp_bytes, keys_in, keys_out, vals_in, vals_out, (long long)n,
                                                          chunk, off, off + 1, 0, 64, s));
    TK(cudaMalloc(&temp, temp_bytes));
    for (int k0 = 0; k0 < K; k0 += chunk) {
        const int nk = std::min(chunk, K - k0);
        k_topic_keys<<<1184, 256, 0, s>>>(Elt, lse, k0, nk, V, KP, keys_in, vals_in);
        k_topic_offsets<<<(nk + 256) / 256, 256, 0, s>>>(nk, V, off);
        TK(cub::DeviceSegmentedRadixSort::SortPairsDescending(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out,
                                                              (long long)nk * V, nk, off, off + 1, 0, 64, s));
        k_take_top<<<(nk * top + 255) / 256, 256, 0, s>>>(keys_out, vals_out, nk, V, top, prob, idx);
        TK(cudaGetLastError());
        TK(cudaMemcpyAsync(prob_out + (size_t)k0 * top, prob, (size_t)nk * top * sizeof(double), cudaMemcpyDeviceToHost, s));
        TK(cudaMemcpyAsync(idx_out + (size_t)k0 * top, idx, (size_t)nk * top * sizeof(int), cudaMemcpyDeviceToHost, s));
        TK(cudaStreamSynchronize(s));
    }
    rc = 0;
done:
    cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_in); cudaFree(vals_out); cudaFree(off); cudaFree(prob); cudaFree(idx);
    cudaFree(temp);
    return rc;
}

}  // namespace pylda
