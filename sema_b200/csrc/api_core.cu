// api_core.cu — lifecycle, ingest (K1), tombstones, compaction, disk cache, properties.
#include "index_impl.cuh"
#include "k1_ingest.cuh"

using namespace sema;
using namespace sema_impl;

namespace sema_impl {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int poll_ingest(sema_index *s, bool wait)
{
    if (s->poisoned) return fail(SEMA_ERR_CUDA, "index unusable: an earlier compaction failed half-way");
    while (!s->pending.empty()) {
        Pending &p = s->pending.front();
        cudaError_t e = wait ? cudaEventSynchronize(p.ev) : cudaEventQuery(p.ev);
        if (e == cudaErrorNotReady) break;
        if (e != cudaSuccess) return fail(SEMA_ERR_CUDA, "ingest event: %s", cudaGetErrorString(e));
        s->n_visible = p.rows_after;
        cudaEventDestroy(p.ev);
        s->pending.pop_front();
    }
    return SEMA_OK;
}

int ensure(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return SEMA_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CK(cudaMalloc(p, need));
    *cap = need;
    return SEMA_OK;
}

// K1 grid: a warp owns groups of 32 rows; enough blocks to fill the GPU, never more than the groups need
static unsigned ingest_blocks(const sema_index *s, uint64_t n)
{
    const uint64_t groups = (n + 31) / 32;
    uint64_t blocks = (groups + INGEST_SEQ_WARPS - 1) / INGEST_SEQ_WARPS;
    const uint64_t maxb = (uint64_t)s->num_sms * 12;
    if (blocks > maxb) blocks = maxb;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

// K1 on query vectors in place (nq rows of `stride` floats on the device, on the query stream)
int normalize_queries_dev(sema_index *s, float *q, uint64_t stride, uint32_t nq)
{
    for (uint32_t done = 0; done < nq; done += 65536) {
        const uint32_t m = (nq - done) < 65536u ? (nq - done) : 65536u;
        float *base = q + (size_t)done * stride;
        // src == dst, same stride; pad columns [dim, stride) are rewritten as zeros.  The same kernel
        // (the reference's sequential sum order) normalises stored rows, and the host-query path
        // repeats that order in K2's registers (k2_scan_tma.cuh), so every path produces the same
        // normalised query bit for bit.
        const bool vec4 = stride == s->dim && (s->dim & 3u) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
        const unsigned blocks = ingest_blocks(s, m);
        if (vec4)
            ingest_kernel<4><<<blocks, INGEST_SEQ_WARPS * 32, 0, s->stream>>>(
                base, stride, base, (uint32_t)stride, s->dim, m, nullptr, s->qscratch,
                1, reinterpret_cast<float *>(s->qscratch + 65536));
        else
            ingest_kernel<1><<<blocks, INGEST_SEQ_WARPS * 32, 0, s->stream>>>(
                base, stride, base, (uint32_t)stride, s->dim, m, nullptr, s->qscratch,
                1, reinterpret_cast<float *>(s->qscratch + 65536));
        CK(cudaGetLastError());
        s->launches++;
    }
    return SEMA_OK;
}

}  // namespace sema_impl

namespace {

int launch_ingest(sema_index *s, const float *src, uint64_t src_ld, uint64_t first, uint64_t n,
                  const uint8_t *valid_in, int normalize, bool vec4)
{
    SEMA_NVTX("sema.K1.ingest");
    float *dst = s->X + first * s->ld;
    const unsigned blocks = ingest_blocks(s, n);
    if (vec4)
        ingest_kernel<4><<<blocks, INGEST_SEQ_WARPS * 32, 0, s->ingest_stream>>>(
            src, src_ld, dst, s->ld, s->dim, n, valid_in, s->valid + first, normalize, s->max_norm2);
    else
        ingest_kernel<1><<<blocks, INGEST_SEQ_WARPS * 32, 0, s->ingest_stream>>>(
            src, src_ld, dst, s->ld, s->dim, n, valid_in, s->valid + first, normalize, s->max_norm2);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

int publish(sema_index *s, uint64_t n)
{
    s->n_rows += n;
    Pending p;
    CK(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming));
    CK(cudaEventRecord(p.ev, s->ingest_stream));
    p.rows_after = s->n_rows;
    s->pending.push_back(p);
    return SEMA_OK;
}


int check_append(sema_index *s, uint64_t n)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (s->poisoned) return fail(SEMA_ERR_CUDA, "index unusable: an earlier compaction failed half-way");
    if (s->n_rows + n > s->capacity)
        return fail(SEMA_ERR_CAPACITY, "append of %llu rows exceeds capacity %llu (size %llu)",
                    (unsigned long long)n, (unsigned long long)s->capacity, (unsigned long long)s->n_rows);
    if ((uint64_t)s->row_base + s->n_rows + n > 0xfffffffeull)
        return fail(SEMA_ERR_CAPACITY, "global row ids must stay below 2^32-1");
    if (s->growable && n) {   // back the new rows with physical memory (existing mappings are untouched: searches may be running)
        CK(cudaSetDevice(s->device));
        int rc = growbuf_commit(s->gX, (size_t)(s->n_rows + n) * s->ld * sizeof(float));
        if (rc) return rc;
        rc = growbuf_commit(s->gValid, (size_t)(s->n_rows + n));
        if (rc) return rc;
    }
    return SEMA_OK;
}

}  // namespace

extern "C" {

const char *sema_last_error(void) { return g_err; }
const char *sema_version(void) { return "sema_b200 0.1 (sm_100a)"; }

int sema_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int sema_host_alloc(void **out, size_t bytes)
{
    if (!out) return fail(SEMA_ERR_INVALID, "null out");
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return SEMA_OK;
}

int sema_host_free(void *p)
{
    if (p) CK(cudaFreeHost(p));
    return SEMA_OK;
}

static int create_impl(int device, uint32_t dim, uint64_t capacity_rows, int metric, bool growable, sema_index **out);

int sema_index_create(int device, uint32_t dim, uint64_t capacity_rows, int metric, sema_index **out)
{
    return create_impl(device, dim, capacity_rows, metric, false, out);
}

int sema_index_create_growable(int device, uint32_t dim, uint64_t max_rows, int metric, sema_index **out)
{
    if (max_rows == 0) return fail(SEMA_ERR_INVALID, "max_rows = 0");
    return create_impl(device, dim, max_rows, metric, true, out);
}

static int create_impl(int device, uint32_t dim, uint64_t capacity_rows, int metric, bool growable, sema_index **out)
{
    if (!out) return fail(SEMA_ERR_INVALID, "null out");
    *out = nullptr;
    if (dim == 0 || dim > SEMA_MAX_DIM) return fail(SEMA_ERR_INVALID, "dim %u outside [1, %u]", dim, SEMA_MAX_DIM);
    if (metric != SEMA_METRIC_COSINE && metric != SEMA_METRIC_L2) return fail(SEMA_ERR_INVALID, "unknown metric %d", metric);
    if (capacity_rows > 0xfffffffeull) return fail(SEMA_ERR_INVALID, "capacity_rows must be < 2^32-1");
    int ndev = sema_device_count();
    if (ndev == 0) return fail(SEMA_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(SEMA_ERR_INVALID, "device %d outside [0, %d)", device, ndev);
    CK(cudaSetDevice(device));
    sema_index *s = new (std::nothrow) sema_index();
    if (!s) return fail(SEMA_ERR_NOMEM, "host allocation failed");
    s->device = device;
    s->dim = dim;
    s->ld = (dim + 3u) & ~3u;
    s->capacity = capacity_rows;
    s->metric = metric;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete s; return fail(SEMA_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    s->num_sms = prop.multiProcessorCount;
    const size_t xbytes = (size_t)(capacity_rows ? capacity_rows : 1) * s->ld * sizeof(float);
    const size_t res_bytes = 8 + (size_t)SEMA_MAX_K * 12;
#define CKD(call)                                                                           \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            int code_ = e_ == cudaErrorMemoryAllocation ? SEMA_ERR_NOMEM : SEMA_ERR_CUDA;   \
            fail(code_, "%s failed: %s", #call, cudaGetErrorString(e_));                    \
            cudaGetLastError();                                                             \
            sema_index_destroy(s);                                                          \
            return code_;                                                                   \
        }                                                                                   \
    } while (0)
    CKD(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    CKD(cudaStreamCreateWithFlags(&s->ingest_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    if (growable) {
        s->growable = true;
        // 256 MB steps for the matrix (the growth is amortised over ~170 k rows of dim 384), granularity-sized ones for the validity bytes
        int rc_ = growbuf_reserve(s->gX, device, xbytes, (size_t)256 << 20);
        if (rc_ == SEMA_OK) rc_ = growbuf_reserve(s->gValid, device, capacity_rows, 0);
        if (rc_ != SEMA_OK) { sema_index_destroy(s); return rc_; }
        s->X = reinterpret_cast<float *>(s->gX.base);
        s->valid = reinterpret_cast<uint8_t *>(s->gValid.base);
    } else {
        CKD(cudaMalloc(&s->X, xbytes));
        CKD(cudaMalloc(&s->valid, capacity_rows ? capacity_rows : 1));
    }
    CKD(cudaMalloc(&s->q_dev, s->ld * sizeof(float)));
    CKD(cudaMemset(s->q_dev, 0, s->ld * sizeof(float)));
    CKD(cudaHostAlloc(&s->q_pin, s->ld * sizeof(float), cudaHostAllocPortable));
    memset(s->q_pin, 0, s->ld * sizeof(float));
    CKD(cudaMalloc(&s->partials, (size_t)s->num_sms * MAX_BLOCKS_PER_SM * K_PASS * sizeof(uint64_t)));
    CKD(cudaMalloc(&s->ticket, 4 * sizeof(unsigned int)));
    CKD(cudaMemset(s->ticket, 0, 4 * sizeof(unsigned int)));
    CKD(cudaMalloc(&s->stream_ctl, STREAM_CTL_WORDS * sizeof(unsigned int)));
    CKD(cudaMemset(s->stream_ctl, 0, STREAM_CTL_WORDS * sizeof(unsigned int)));
    CKD(cudaMalloc(&s->keys_dev, SEMA_MAX_K * sizeof(uint64_t)));
    CKD(cudaMalloc(&s->max_norm2, 2 * sizeof(float)));
    CKD(cudaMalloc(&s->qscratch, 65536 + 32));
    {
        const float init[2] = {0.0f, __builtin_inff()};   // [max |x|^2, min |x|^2] over the rows K1 has seen
        CKD(cudaMemcpy(s->max_norm2, init, sizeof init, cudaMemcpyHostToDevice));
    }
    CKD(cudaMalloc(&s->res_dev, res_bytes));
    CKD(cudaHostAlloc(&s->res_pin, res_bytes, cudaHostAllocPortable));
    // mapped result block of the host-query path: the kernel stores results + completion flag here
    CKD(cudaHostAlloc(&s->res_map, RES_MAP_BYTES, cudaHostAllocPortable | cudaHostAllocMapped));
    memset(s->res_map, 0, RES_MAP_BYTES);
    CKD(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s->res_map_dev), s->res_map, 0));
    CKD(cudaDeviceSynchronize());
#undef CKD
    *out = s;
    return SEMA_OK;
}

int sema_index_destroy(sema_index *s)
{
    if (!s) return SEMA_OK;
    cudaSetDevice(s->device);
    if (s->own_stream) cudaStreamSynchronize(s->own_stream);
    if (s->ingest_stream) cudaStreamSynchronize(s->ingest_stream);
    for (auto &p : s->pending) cudaEventDestroy(p.ev);
    if (s->growable) {
        growbuf_free(s->gX); growbuf_free(s->gValid); growbuf_free(s->gPlanes);
        s->X = nullptr; s->valid = nullptr; s->planes = nullptr;
    }
    cudaFree(s->X); cudaFree(s->valid); cudaFree(s->q_dev); cudaFreeHost(s->q_pin);
    cudaFree(s->cs.flags); cudaFree(s->cs.new_valid); cudaFree(s->cs.src); cudaFree(s->cs.tiles); cudaFree(s->cs.map);
    cudaFree(s->cs.head); cudaFree(s->cs.tmp);
    cudaFree(s->partials); cudaFree(s->ticket); cudaFree(s->stream_ctl); cudaFree(s->keys_dev); cudaFree(s->res_dev);
    cudaFreeHost(s->res_pin); cudaFreeHost(s->res_map); cudaFree(s->Q_dev); cudaFree(s->bids_dev); cudaFree(s->bsc_dev);
    cudaFree(s->bnf_dev); cudaFree(s->tomb_dev);
    cudaFree(s->qscratch); cudaFree(s->max_norm2); cudaFree(s->planes); cudaFree(s->Qpad_dev); cudaFree(s->cand_rows);
    cudaFree(s->q_aligned); cudaFree(s->sub_q); cudaFree(s->sub_ids); cudaFree(s->sub_sc); cudaFree(s->sub_nf); cudaFree(s->sub_idx);
    cudaFree(s->cand_thr); cudaFree(s->cand_sc); cudaFree(s->flags_dev); cudaFreeHost(s->flags_pin);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->ingest_stream) cudaStreamDestroy(s->ingest_stream);
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    cudaGetLastError();
    delete s;
    return SEMA_OK;
}

int sema_index_append_async(sema_index *s, const float *rows, uint64_t n, const uint8_t *valid,
                            int normalize, uint64_t *first_row)
{
    int rc = check_append(s, n);
    if (rc) return rc;
    if (n && !rows) return fail(SEMA_ERR_INVALID, "null rows");
    CK(cudaSetDevice(s->device));
    const uint64_t first = s->n_rows;
    if (first_row) *first_row = first;
    if (n == 0) return SEMA_OK;
    float *dst = s->X + first * s->ld;
    // host rows land directly in their final place; K1 then normalises in place
    if (s->ld == s->dim)
        CK(cudaMemcpyAsync(dst, rows, n * s->dim * sizeof(float), cudaMemcpyHostToDevice, s->ingest_stream));
    else
        CK(cudaMemcpy2DAsync(dst, s->ld * sizeof(float), rows, s->dim * sizeof(float),
                             s->dim * sizeof(float), n, cudaMemcpyHostToDevice, s->ingest_stream));
    const uint8_t *vin = nullptr;
    if (valid) {
        CK(cudaMemcpyAsync(s->valid + first, valid, n, cudaMemcpyHostToDevice, s->ingest_stream));
        vin = s->valid + first;
    }
    rc = launch_ingest(s, dst, s->ld, first, n, vin, normalize, s->ld == s->dim);
    if (rc) return rc;
    return publish(s, n);
}

int sema_index_flush(sema_index *s)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->ingest_stream));
    return poll_ingest(s, true);
}

int sema_index_append(sema_index *s, const float *rows, uint64_t n, const uint8_t *valid,
                      int normalize, uint64_t *first_row)
{
    int rc = sema_index_append_async(s, rows, n, valid, normalize, first_row);
    if (rc) return rc;
    return sema_index_flush(s);
}

int sema_index_append_device(sema_index *s, const float *rows_dev, uint64_t n,
                             const uint8_t *valid_dev, int normalize, uint64_t *first_row)
{
    int rc = check_append(s, n);
    if (rc) return rc;
    if (n && !rows_dev) return fail(SEMA_ERR_INVALID, "null rows");
    CK(cudaSetDevice(s->device));
    const uint64_t first = s->n_rows;
    if (first_row) *first_row = first;
    if (n == 0) return SEMA_OK;
    const bool vec4 = s->ld == s->dim && (reinterpret_cast<uintptr_t>(rows_dev) & 15) == 0;
    rc = launch_ingest(s, rows_dev, s->dim, first, n, valid_dev, normalize, vec4);
    if (rc) return rc;
    rc = publish(s, n);
    if (rc) return rc;
    return sema_index_flush(s);
}

int sema_index_append_pooled_device(sema_index *s, const float *tokens_dev, const float *mask_dev, uint64_t n,
                                    uint32_t seq_len, const uint8_t *valid_dev, int skip_masked, uint64_t *first_row)
{
    int rc = check_append(s, n);
    if (rc) return rc;
    if (n && (!tokens_dev || !mask_dev)) return fail(SEMA_ERR_INVALID, "null tokens / mask");
    if (seq_len == 0) return fail(SEMA_ERR_INVALID, "seq_len = 0");
    CK(cudaSetDevice(s->device));
    const uint64_t first = s->n_rows;
    if (first_row) *first_row = first;
    if (n == 0) return SEMA_OK;
    // K0 pools + normalises straight into the rows' final place; K1 (normalize = 0, in place) then
    // only does the column's bookkeeping: null / non-finite rows -> NaN, validity bytes, max norm
    float *dst = s->X + first * s->ld;
    rc = launch_pool(s, s->ingest_stream, tokens_dev, mask_dev, n, seq_len, skip_masked, dst, s->ld);
    if (rc) return rc;
    rc = launch_ingest(s, dst, s->ld, first, n, valid_dev, 0, s->ld == s->dim);
    if (rc) return rc;
    rc = publish(s, n);
    if (rc) return rc;
    return sema_index_flush(s);
}

int sema_index_append_synthetic(sema_index *s, uint64_t seed, uint64_t synth_row0, uint64_t n,
                                int normalize, uint64_t *first_row)
{
    int rc = check_append(s, n);
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    const uint64_t first = s->n_rows;
    if (first_row) *first_row = first;
    if (n == 0) return SEMA_OK;
    float *dst = s->X + first * s->ld;
    synth_kernel<<<s->num_sms * 16, INGEST_THREADS, 0, s->ingest_stream>>>(dst, s->ld, s->dim, seed, synth_row0, n);
    CK(cudaGetLastError());
    s->launches++;
    rc = launch_ingest(s, dst, s->ld, first, n, nullptr, normalize, s->ld == s->dim);
    if (rc) return rc;
    rc = publish(s, n);
    if (rc) return rc;
    return sema_index_flush(s);
}

int sema_index_tombstone(sema_index *s, const uint64_t *rows, uint64_t n)
{
    SEMA_NVTX("sema.tombstone");
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (n && !rows) return fail(SEMA_ERR_INVALID, "null rows");
    if (n == 0) return SEMA_OK;
    int rc = sema_index_flush(s);
    if (rc) return rc;
    for (uint64_t i = 0; i < n; ++i)
        if (rows[i] >= s->n_rows)
            return fail(SEMA_ERR_INVALID, "tombstone row %llu >= size %llu", (unsigned long long)rows[i], (unsigned long long)s->n_rows);
    rc = ensure(reinterpret_cast<void **>(&s->tomb_dev), &s->tomb_cap, n * sizeof(uint64_t));
    if (rc) return rc;
    CK(cudaMemcpyAsync(s->tomb_dev, rows, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    uint64_t blocks = (n * 32 + INGEST_THREADS - 1) / INGEST_THREADS;
    if (blocks > (uint64_t)s->num_sms * 16) blocks = (uint64_t)s->num_sms * 16;
    tombstone_kernel<<<(unsigned)blocks, INGEST_THREADS, 0, s->stream>>>(s->X, s->ld, s->tomb_dev, n, s->n_rows, s->valid);
    CK(cudaGetLastError());
    s->launches++;
    // the bf16 planes of K3 must forget the dead rows too: the same rows are NaN-poisoned in place
    // (16 bytes per plane and 8 columns), so removing a file never re-tiles the planes behind it
    rc = k3_poison_rows(s, s->tomb_dev, n);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));
    return SEMA_OK;
}


int sema_index_compact(sema_index *s, uint64_t *new_row_of_old, uint64_t *n_live_out)
{
    return sema_index_compact_keep(s, nullptr, new_row_of_old, n_live_out);
}

int sema_index_compact_keep(sema_index *s, const uint8_t *keep, uint64_t *new_row_of_old, uint64_t *n_live_out)
{
    SEMA_NVTX("sema.compact");
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    int rc = sema_index_flush(s);
    if (rc) return rc;
    if (s->poisoned) return fail(SEMA_ERR_CUDA, "index unusable: an earlier compaction failed half-way");
    CK(cudaStreamSynchronize(s->stream));
    const uint64_t n = s->n_rows;
    if (n == 0) {
        if (n_live_out) *n_live_out = 0;
        return SEMA_OK;
    }
    // SEMA_TRACE=1: phase times of this call on stderr (host clock; every phase below ends synchronised or is host work)
    static const bool trace = getenv("SEMA_TRACE") != nullptr;
    timespec t_prev;
    clock_gettime(CLOCK_MONOTONIC, &t_prev);
    auto phase = [&](const char *name) {
        if (!trace) return;
        timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[sema compact] %-28s %9.3f ms\n", name, (t.tv_sec - t_prev.tv_sec) * 1e3 + (t.tv_nsec - t_prev.tv_nsec) * 1e-6);
        t_prev = t;
    };
    // ---- plan on the device (k1_ingest.cuh): which rows stay (the caller's keep flags, else the validity bytes),
    // their new positions by prefix sums, the ascending gather list src[new] = old, the compacted validity bytes and
    // the old -> new map the caller's chunk table needs.  Nothing has moved yet: a failure here is an ordinary error.
    const uint32_t tiles = (uint32_t)((n + COMPACT_TILE - 1) / COMPACT_TILE);
    const uint64_t C = 1u << 16;                         // rows per gather chunk
    sema_index::CompactScratch &d = s->cs;      // grown on demand, kept between calls, freed with the handle
    uint32_t *tile_count = nullptr, *tile_off = nullptr;
    {
        auto grow = [&](auto **p, size_t *cap, size_t bytes) { return ensure(reinterpret_cast<void **>(p), cap, bytes ? bytes : 1); };
        int arc = SEMA_OK;
        if (keep) arc = grow(&d.flags, &d.flags_cap, n);
        if (!arc) arc = grow(&d.new_valid, &d.new_valid_cap, n);
        if (!arc) arc = grow(&d.src, &d.src_cap, n * sizeof(uint32_t));
        if (!arc) arc = grow(&d.tiles, &d.tiles_cap, 2 * (size_t)tiles * sizeof(uint32_t));
        if (!arc) arc = grow(&d.head, &d.head_cap, 2 * sizeof(unsigned long long));
        if (!arc && new_row_of_old) arc = grow(&d.map, &d.map_cap, n * sizeof(unsigned long long));
        if (arc) return arc;                     // ensure() has set the error text; nothing has moved
        tile_count = d.tiles;
        tile_off = d.tiles + tiles;
    }
    cudaError_t e = cudaSuccess;
    phase("scratch allocation");
    const uint8_t *flags = s->valid;
    if (keep) {
        CK(cudaMemcpyAsync(d.flags, keep, n, cudaMemcpyHostToDevice, s->stream));
        flags = d.flags;
    }
    unsigned long long head[2] = {0ull, (unsigned long long)n};
    CK(cudaMemcpyAsync(d.head, head, sizeof head, cudaMemcpyHostToDevice, s->stream));
    compact_count_kernel<<<tiles, INGEST_THREADS, 0, s->stream>>>(flags, n, tile_count, d.head + 1);
    compact_scan_tiles_kernel<<<1, INGEST_THREADS, 0, s->stream>>>(tile_count, tiles, tile_off, d.head);
    compact_emit_kernel<<<tiles, INGEST_THREADS, 0, s->stream>>>(flags, s->valid, n, tile_off, d.src, d.new_valid,
                                                                 new_row_of_old ? d.map : nullptr);
    CK(cudaGetLastError());
    s->launches += 3;
    CK(cudaMemcpyAsync(head, d.head, sizeof head, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const uint64_t live = head[0], first_moved = head[1] < live ? head[1] : live;   // rows before the first dropped row stay put
    phase("plan kernels");
    if (new_row_of_old) CK(cudaMemcpy(new_row_of_old, d.map, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    phase("old->new map to the host");
    if (n_live_out) *n_live_out = live;
    if (live == n) return SEMA_OK;  // nothing to drop
    // first source row of every gather chunk (one strided copy): decides bounce / direct below
    const uint64_t n_chunks = (live - first_moved + C - 1) / C;
    std::vector<uint32_t> chunk_src;
    try {
        chunk_src.resize(n_chunks ? n_chunks : 1);
    } catch (const std::exception &) {
        return fail(SEMA_ERR_NOMEM, "host allocation failed");
    }
    if (n_chunks)
        CK(cudaMemcpy2D(chunk_src.data(), sizeof(uint32_t), d.src + first_moved, C * sizeof(uint32_t), sizeof(uint32_t), n_chunks,
                        cudaMemcpyDeviceToHost));
    bool any_bounce = false;
    for (uint64_t ci = 0, at = first_moved; ci < n_chunks; ++ci, at += C)
        any_bounce = any_bounce || chunk_src[ci] < at + ((live - at) < C ? (live - at) : C);
    if (any_bounce) {       // the bounce buffer (one chunk of rows) only when some chunk's sources overlap its destination
        const int arc = ensure(reinterpret_cast<void **>(&d.tmp), &d.tmp_cap, C * s->ld * sizeof(float));
        if (arc) return arc;                         // nothing has moved yet
    }
    phase("chunk heads to the host");
    // ---- ordered gather, chunk by chunk in ascending order, everything queued on the query stream with ONE
    // synchronise at the end.  src[i] >= i, so a chunk's destination [at, at + m) never overlaps a LATER chunk's
    // sources.  It may overlap its OWN sources while fewer than m rows have been dropped before it: such a chunk goes
    // through a bounce buffer (gather, then copy into place: 2 reads + 2 writes per row).  As soon as
    // src[at] >= at + m — from the first 65 536 dropped rows on — chunks are gathered straight into place (1 read +
    // 1 write per row).
    // From the first chunk on, rows move in place: an error half-way would leave X out of step with the
    // validity bytes, the row count and the caller's chunk table.  It is reported, the scratch
    // buffers are released, and the handle refuses further work instead of answering from mixed rows.
    const char *what = "gather kernel";
    uint64_t ci = 0;
    for (uint64_t at = first_moved; at < live && e == cudaSuccess; at += C, ++ci) {
        const uint64_t m = (live - at) < C ? (live - at) : C;
        const bool direct = chunk_src[ci] >= at + m;       // every source of this chunk lies beyond its destination
        uint64_t blocks = (m * 32 + INGEST_THREADS - 1) / INGEST_THREADS;
        if (blocks > (uint64_t)s->num_sms * 16) blocks = (uint64_t)s->num_sms * 16;
        gather_rows_kernel<<<(unsigned)blocks, INGEST_THREADS, 0, s->stream>>>(s->X, s->ld, d.src + at, m, direct ? s->X + at * s->ld : d.tmp);
        what = "gather kernel";
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        s->launches++;
        if (!direct) {
            what = "moving a chunk into place";          // stream order keeps tmp safe: the next gather runs after this copy
            e = cudaMemcpyAsync(s->X + at * s->ld, d.tmp, m * s->ld * sizeof(float), cudaMemcpyDeviceToDevice, s->stream);
        }
    }
    if (e == cudaSuccess && live) {
        what = "writing the validity bytes";
        e = cudaMemcpyAsync(s->valid, d.new_valid, live, cudaMemcpyDeviceToDevice, s->stream);
    }
    if (e == cudaSuccess) { what = "final synchronise"; e = cudaStreamSynchronize(s->stream); }
    phase("gather + validity bytes");
    if (e != cudaSuccess) {
        cudaGetLastError();
        s->poisoned = true;
        return fail(SEMA_ERR_CUDA, "compaction failed while %s: %s; the index is now unusable (destroy and rebuild it)", what,
                    cudaGetErrorString(e));
    }
    s->n_rows = live;
    s->n_visible = live;
    if (s->planes_rows > first_moved) s->planes_rows = first_moved;   // K3 planes: re-tile from the first moved row
    return SEMA_OK;
}

namespace {
constexpr uint32_t SEMA_FILE_VERSION = 1;
constexpr uint32_t SEMA_FILE_ENDIAN_TAG = 0x01020304u;   // reads back differently on a machine of the other byte order
struct SemaFileHeader {
    char magic[8];
    uint32_t dim;
    int32_t metric;
    uint64_t n_rows;
    uint32_t version;       // SEMA_FILE_VERSION (0 in files written before the field existed = version 1)
    uint32_t endian_tag;    // SEMA_FILE_ENDIAN_TAG (0 in those older files)
    unsigned char pad[32];
};
static_assert(sizeof(SemaFileHeader) == 64, "header is 64 bytes");
}  // namespace

int sema_index_save(sema_index *s, const char *path)
{
    if (!s || !path) return fail(SEMA_ERR_INVALID, "null argument");
    int rc = sema_index_flush(s);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(SEMA_ERR_INVALID, "cannot open %s for writing", path);
    SemaFileHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "SEMAIDX1", 8);
    h.dim = s->dim;
    h.metric = s->metric;
    h.n_rows = s->n_rows;
    h.version = SEMA_FILE_VERSION;
    h.endian_tag = SEMA_FILE_ENDIAN_TAG;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    const uint64_t n = s->n_rows;
    const uint64_t C = 1u << 16;
    uint8_t *vbuf = static_cast<uint8_t *>(malloc(C));     // validity bytes travel in chunks: nothing scales with n_rows
    ok = ok && vbuf != nullptr;
    for (uint64_t at = 0; ok && at < n; at += C) {
        const uint64_t m = (n - at) < C ? (n - at) : C;
        cudaError_t e = cudaMemcpy(vbuf, s->valid + at, m, cudaMemcpyDeviceToHost);
        ok = e == cudaSuccess && fwrite(vbuf, 1, m, f) == m;
    }
    free(vbuf);
    float *pin = nullptr;
    if (ok && n) ok = cudaHostAlloc(&pin, C * s->dim * sizeof(float), cudaHostAllocDefault) == cudaSuccess;
    for (uint64_t at = 0; ok && at < n; at += C) {
        const uint64_t m = (n - at) < C ? (n - at) : C;
        cudaError_t e = cudaMemcpy2D(pin, s->dim * sizeof(float), s->X + at * s->ld, s->ld * sizeof(float),
                                     s->dim * sizeof(float), m, cudaMemcpyDeviceToHost);
        ok = e == cudaSuccess && fwrite(pin, sizeof(float), m * s->dim, f) == m * s->dim;
    }
    if (pin) cudaFreeHost(pin);
    ok = (fclose(f) == 0) && ok;
    cudaGetLastError();
    return ok ? SEMA_OK : fail(SEMA_ERR_CUDA, "writing %s failed", path);
}

int sema_index_load(const char *path, int device, uint64_t capacity_rows, sema_index **out)
{
    if (!path || !out) return fail(SEMA_ERR_INVALID, "null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(SEMA_ERR_INVALID, "cannot open %s", path);
    SemaFileHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "SEMAIDX1", 8) != 0) {
        fclose(f);
        return fail(SEMA_ERR_INVALID, "%s is not a sema index file", path);
    }
    // Nothing in the header is trusted before it has been checked against the file itself: a corrupt
    // or truncated file must come back as an error code, never as an allocation failure thrown
    // across the C boundary.
    const uint32_t version = h.version ? h.version : 1;   // files written before the field existed carry 0
    if (version != SEMA_FILE_VERSION || (h.endian_tag != 0 && h.endian_tag != SEMA_FILE_ENDIAN_TAG)) {
        fclose(f);
        return fail(SEMA_ERR_INVALID, "%s: unsupported file version %u or byte order", path, version);
    }
    if (h.dim == 0 || h.dim > SEMA_MAX_DIM || (h.metric != SEMA_METRIC_COSINE && h.metric != SEMA_METRIC_L2) ||
        h.n_rows > 0xfffffffeull) {
        fclose(f);
        return fail(SEMA_ERR_INVALID, "%s: corrupt header (dim %u, metric %d, rows %llu)", path, h.dim, h.metric,
                    (unsigned long long)h.n_rows);
    }
    long long fsize = -1;
    if (fseek(f, 0, SEEK_END) == 0) fsize = ftell(f);
    const unsigned long long want = 64ull + h.n_rows + h.n_rows * (unsigned long long)h.dim * 4ull;
    if (fsize < 0 || (unsigned long long)fsize != want || fseek(f, (long)sizeof h, SEEK_SET) != 0) {
        fclose(f);
        return fail(SEMA_ERR_INVALID, "%s: size %lld does not match its header (%llu rows x %u: %llu bytes)", path, fsize,
                    (unsigned long long)h.n_rows, h.dim, want);
    }
    if (capacity_rows < h.n_rows) capacity_rows = h.n_rows;
    sema_index *s = nullptr;
    int rc = sema_index_create(device, h.dim, capacity_rows, h.metric, &s);
    if (rc) { fclose(f); return rc; }
    const uint64_t n = h.n_rows, C = 1u << 16;
    // the validity bytes are read chunk by chunk next to their rows: no allocation scales with n_rows
    uint8_t *vbuf = static_cast<uint8_t *>(malloc(C));
    float *pin = nullptr;
    bool ok = vbuf != nullptr;
    if (!ok) rc = fail(SEMA_ERR_NOMEM, "host allocation failed");
    if (ok && n) {
        ok = cudaHostAlloc(&pin, C * h.dim * sizeof(float), cudaHostAllocDefault) == cudaSuccess;
        if (!ok) { cudaGetLastError(); rc = fail(SEMA_ERR_NOMEM, "pinned staging buffer"); }
    }
    for (uint64_t at = 0; ok && at < n; at += C) {
        const uint64_t m = (n - at) < C ? (n - at) : C;
        ok = fseek(f, (long)(sizeof h + at), SEEK_SET) == 0 && fread(vbuf, 1, m, f) == m &&
             fseek(f, (long)(sizeof h + n + at * h.dim * sizeof(float)), SEEK_SET) == 0 &&
             fread(pin, sizeof(float), m * h.dim, f) == m * h.dim;
        if (ok) {
            rc = sema_index_append(s, pin, m, vbuf, /*normalize=*/0, nullptr);   // rows are stored normalised
            ok = rc == SEMA_OK;
        }
    }
    if (pin) cudaFreeHost(pin);
    free(vbuf);
    fclose(f);
    if (!ok) {
        sema_index_destroy(s);
        return rc ? rc : fail(SEMA_ERR_INVALID, "%s is truncated", path);
    }
    *out = s;
    return SEMA_OK;
}

int sema_index_set_normalize_queries(sema_index *s, int on)
{
    if (!s) return -1;
    if (on >= 0) s->normalize_queries = on ? 1 : 0;
    return s->normalize_queries;
}

int sema_index_set_row_base(sema_index *s, uint64_t row_base)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (row_base + s->capacity > 0xfffffffeull) return fail(SEMA_ERR_INVALID, "row_base + capacity must be < 2^32-1");
    s->row_base = (uint32_t)row_base;
    return SEMA_OK;
}

int sema_index_set_stream(sema_index *s, void *cuda_stream, int external)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    s->stream = external ? reinterpret_cast<cudaStream_t>(cuda_stream) : s->own_stream;
    return SEMA_OK;
}

uint64_t sema_index_size(const sema_index *s) { return s ? s->n_rows : 0; }
uint64_t sema_index_visible(sema_index *s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    poll_ingest(s, false);
    return s->n_visible;
}
uint64_t sema_index_capacity(const sema_index *s) { return s ? s->capacity : 0; }
uint32_t sema_index_dim(const sema_index *s) { return s ? s->dim : 0; }
int sema_index_device(const sema_index *s) { return s ? s->device : -1; }
uint64_t sema_index_last_snapshot(const sema_index *s) { return s ? s->last_snapshot : 0; }
uint64_t sema_index_launch_count(const sema_index *s) { return s ? s->launches : 0; }

int sema_index_set_scan_variant(sema_index *s, int variant)
{
    if (!s) return -1;
#ifdef SEMA_K3_PROBES
    if (variant >= 1000000) { s->k3_debug = variant - 1000000; return variant; }   // probe builds: wide probe masks
#endif
    if (variant >= 1300) return -1;
    if (variant >= 1200) { s->k3_mix_w = variant - 1200; return variant; }    // 1200 + w: rows of a 4-cluster partition = (0.70 + w/100) x rows of a 2-cluster partition (0 = built-in 1.05)
    if (variant >= 1102) return -1;
    if (variant >= 1100) { s->k3_mixed = variant - 1100; return variant; }   // 1100 / 1101 = a K3 stage as one launch / as two concurrent launches (clusters of 4 + clusters of 2, default)
    if (variant >= 1000) return -1;
    if (variant >= 903) return -1;
    if (variant >= 900) { s->stream_mode = variant - 900; return variant; }  // query streams: 900 = one persistent launch when a scan is long enough (default), 901 = whenever the shape allows, 902 = one launch per query
    if (variant >= 800) { s->k3_prefetch = variant - 800; return variant; }  // 800 + d = K3 producer prefetches into L2 d stages ahead (0 = off)
    if (variant >= 702) return -1;
    if (variant >= 700) { s->k3_pair = variant - 700; return variant; }      // 700 = single-CTA kernel for the single-pass stage (default), 701 = CTA pairs (tcgen05 cta_group::2)
    if (variant >= 600) { s->chain = variant - 600; return variant; }        // 600 / 601 = query streams unchained / chained (PDL)
    if (variant >= 500) { s->host_path = variant - 500; return variant; }    // 500 / 501 = host searches staged through H2D + D2H / query by kernel parameter + mapped results
    if (variant >= 400) { s->k3_kc16 = variant - 400; return variant; }      // 400 = lists of 32, 401 = lists of 16 (k <= 10, single pass)
    if (variant >= 300) {                                                    // 300 = default, 308 = epilogue without its group early-out (same results)
#ifndef SEMA_K3_PROBES
        if (variant != 300 && variant != 308) return -1;                     // the work-skipping timing probes exist only in probe builds
#endif
        s->k3_debug = variant - 300;
        return variant;
    }
    if (variant >= 200) { s->k3_qt = variant - 200; return variant; }        // 200 = auto, 201 = one query tile per CTA
    if (variant >= 100) { s->k3_cluster = variant - 100; return variant; }   // 100 = auto, 101/102/104 = K3 cluster size
    if (variant >= 0) s->variant = variant;
    return s->variant;
}

int sema_index_set_batch_precision(sema_index *s, int prec)
{
    if (!s) return -1;
    if (prec == 0 || prec == 1) s->k3_prec = prec;
    return s->k3_prec;
}

int sema_index_batch_precision_active(const sema_index *s)
{
    if (!s || !s->planes || s->planes_rows == 0) return -1;
    return s->planes_fmt;
}

int sema_index_read_rows(sema_index *s, uint64_t first_row, uint64_t n, float *out)
{
    if (!s || (n && !out)) return fail(SEMA_ERR_INVALID, "null argument");
    int rc = sema_index_flush(s);
    if (rc) return rc;
    if (first_row + n > s->n_rows) return fail(SEMA_ERR_INVALID, "rows [%llu, %llu) outside size %llu",
                                               (unsigned long long)first_row, (unsigned long long)(first_row + n), (unsigned long long)s->n_rows);
    if (n == 0) return SEMA_OK;
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy2D(out, s->dim * sizeof(float), s->X + first_row * s->ld, s->ld * sizeof(float),
                    s->dim * sizeof(float), n, cudaMemcpyDeviceToHost));
    return SEMA_OK;
}

}  // extern "C"
