// api_pool.cu — kernel K0 (mean pooling + normalisation) behind the C ABI: the caller-side step
// immediately before the scan path (reference: src/semantic/embeddings.rs:61-91).
#include "index_impl.cuh"
#include "k0_pool.cuh"

using namespace sema;
using namespace sema_impl;

namespace sema_impl {

int launch_pool(sema_index *s, cudaStream_t stream, const float *tokens_dev, const float *mask_dev, uint64_t n,
                uint32_t seq_len, int skip_masked, float *out_dev, uint64_t out_ld)
{
    SEMA_NVTX("sema.K0.pool");
    const size_t smem = (3 * (size_t)seq_len + s->dim) * sizeof(float);
    if (smem > 200 * 1024) return fail(SEMA_ERR_UNSUPPORTED, "3 x seq_len %u + dim %u floats exceed shared memory", seq_len, s->dim);
    static bool attr_set[64] = {false};
    if (smem > 48 * 1024 && !attr_set[s->device & 63]) {
        CK(cudaFuncSetAttribute(pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[s->device & 63] = true;
    }
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(s->qscratch + 65536 + (stream == s->ingest_stream ? 16 : 8));
    CK(cudaMemsetAsync(ctr, 0, sizeof *ctr, stream));
    // float4 columns when the layout allows (dim 384: 96 threads per text, ~4x more texts resident per SM)
    const bool v4 = (s->dim & 3u) == 0 && (seq_len & 3u) == 0 && s->dim / 4 <= 256 && smem <= 48 * 1024 &&
                    (reinterpret_cast<uintptr_t>(tokens_dev) & 15) == 0;
    if (v4) {
        const uint32_t threads = ((s->dim / 4) + 31u) & ~31u;
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pool_kernel_v4, (int)threads, smem));
        if (per_sm < 1) return fail(SEMA_ERR_UNSUPPORTED, "pool kernel does not fit on an SM");
        uint64_t grid = (uint64_t)s->num_sms * per_sm;
        if (grid > n) grid = n;
        pool_kernel_v4<<<(unsigned)grid, threads, smem, stream>>>(tokens_dev, mask_dev, n, seq_len, s->dim, out_dev, out_ld,
                                                                  skip_masked, ctr);
    } else {
        // one thread per column, in as many equal passes as the block size allows (dim 768 -> 2 x 384)
        const uint32_t passes = (s->dim + POOL_MAX_THREADS - 1) / POOL_MAX_THREADS;
        uint32_t threads = (((s->dim + passes - 1) / passes) + 31u) & ~31u;
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pool_kernel, (int)threads, smem));
        if (per_sm < 1) return fail(SEMA_ERR_UNSUPPORTED, "pool kernel does not fit on an SM");
        uint64_t grid = (uint64_t)s->num_sms * per_sm;          // exactly the resident blocks: texts are claimed dynamically
        if (grid > n) grid = n;
        pool_kernel<<<(unsigned)grid, threads, smem, stream>>>(tokens_dev, mask_dev, n, seq_len, s->dim, out_dev, out_ld,
                                                               skip_masked, ctr);
    }
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

}  // namespace sema_impl

extern "C" {

int sema_mean_pool_device(sema_index *s, const float *tokens_dev, const float *mask_dev, uint64_t n,
                          uint32_t seq_len, int skip_masked, float *out_dev)
{
    if (!s || !tokens_dev || !mask_dev || !out_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (seq_len == 0) return fail(SEMA_ERR_INVALID, "seq_len = 0");
    if (n == 0) return SEMA_OK;
    CK(cudaSetDevice(s->device));
    return launch_pool(s, s->stream, tokens_dev, mask_dev, n, seq_len, skip_masked, out_dev, s->dim);
}

int sema_mean_pool(sema_index *s, const float *tokens, const float *mask, uint64_t n, uint32_t seq_len,
                   int skip_masked, float *out)
{
    if (!s || !tokens || !mask || !out) return fail(SEMA_ERR_INVALID, "null argument");
    if (seq_len == 0) return fail(SEMA_ERR_INVALID, "seq_len = 0");
    if (n == 0) return SEMA_OK;
    CK(cudaSetDevice(s->device));
    const size_t tb = (size_t)n * seq_len * s->dim * sizeof(float), mb = (size_t)n * seq_len * sizeof(float),
                 ob = (size_t)n * s->dim * sizeof(float);
    float *t_d = nullptr, *m_d = nullptr, *o_d = nullptr;
    cudaError_t e = cudaMalloc(&t_d, tb);
    if (e == cudaSuccess) e = cudaMalloc(&m_d, mb);
    if (e == cudaSuccess) e = cudaMalloc(&o_d, ob);
    int rc = SEMA_OK;
    if (e != cudaSuccess) rc = fail(e == cudaErrorMemoryAllocation ? SEMA_ERR_NOMEM : SEMA_ERR_CUDA, "mean_pool staging: %s", cudaGetErrorString(e));
    if (!rc && (e = cudaMemcpyAsync(t_d, tokens, tb, cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) rc = fail(SEMA_ERR_CUDA, "%s", cudaGetErrorString(e));
    if (!rc && (e = cudaMemcpyAsync(m_d, mask, mb, cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) rc = fail(SEMA_ERR_CUDA, "%s", cudaGetErrorString(e));
    if (!rc) rc = launch_pool(s, s->stream, t_d, m_d, n, seq_len, skip_masked, o_d, s->dim);
    if (!rc && (e = cudaMemcpyAsync(out, o_d, ob, cudaMemcpyDeviceToHost, s->stream)) != cudaSuccess) rc = fail(SEMA_ERR_CUDA, "%s", cudaGetErrorString(e));
    if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess && !rc) rc = fail(SEMA_ERR_CUDA, "%s", cudaGetErrorString(e));
    cudaFree(t_d); cudaFree(m_d); cudaFree(o_d);
    cudaGetLastError();
    return rc;
}

}  // extern "C"
