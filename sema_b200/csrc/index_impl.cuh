// index_impl.cuh — internals shared by the translation units behind include/sema_b200.h.
//
// struct sema_index owns the HBM layout (row-major fp32, row stride ld = round_up(dim,4) floats,
// base 256-byte aligned by cudaMalloc, so every row is 16-byte aligned), the streams and the small
// staging buffers.  api_core.cu: lifecycle, ingest (K1), tombstones, compaction, disk cache,
// properties.  api_pool.cu: K0 (mean pooling, the step before the path).  api_search.cu: K2 / K4 dispatch and the single-query entry points.  api_batch.cu:
// K3 (planes, cluster launch, re-scoring, fallback) and the batched entry points.  api_shard.cu:
// the fused peer exchange.  No CPU fallback exists: every compute entry point needs a CUDA device.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include <deque>
#include <exception>
#include <new>
#include <vector>

#include "../../include/sema_b200.h"
#include "common.cuh"
#include "growbuf.cuh"

namespace sema_impl {

int fail(int code, const char *fmt, ...);   // sets the thread-local error text, returns code

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return ::sema_impl::fail(SEMA_ERR_CUDA, "%s failed: %s (%s:%d)", #call,       \
                                     cudaGetErrorString(e_), __FILE__, __LINE__);         \
    } while (0)

// NVTX range around every kernel-launching step (K0-K4, plane split, tombstones, compaction): a timeline tool
// (nsys, ncu --nvtx) shows the path's phases by name; without a tool attached a push / pop is a null-pointer test.
struct NvtxScope {
    explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
    NvtxScope(const NvtxScope &) = delete;
    NvtxScope &operator=(const NvtxScope &) = delete;
};
#define SEMA_NVTX_CAT2(a, b) a##b
#define SEMA_NVTX_CAT(a, b) SEMA_NVTX_CAT2(a, b)
#define SEMA_NVTX(name) ::sema_impl::NvtxScope SEMA_NVTX_CAT(nvtx_scope_, __LINE__)(name)

struct Pending {
    cudaEvent_t ev;
    uint64_t rows_after;
};

constexpr int K_PASS = 128;  // keys one fused pass can select (WarpTopK<4>)
constexpr int MAX_BLOCKS_PER_SM = 8;

// scan_query flags
constexpr unsigned SCAN_CHAINED = 1u;      // directly follows another scan of the same call: may overlap it (PDL)
constexpr unsigned SCAN_HOST_QUERY = 2u;   // q is a HOST pointer; query by kernel parameter, results + flag to res_map

// mapped result slots of the host-query path; one slot: [n_found u32, pad][ids u64 k][scores f32 k] ... [flag u64 @ 2048]
constexpr int RES_SLOTS = SEMA_MAX_INFLIGHT;
constexpr size_t RES_SLOT_BYTES = 4096;
constexpr size_t RES_MAP_BYTES = RES_SLOTS * RES_SLOT_BYTES;
constexpr size_t RES_MAP_FLAG_OFF = 2048;
constexpr int STREAM_FALLBACK = 0x5eaa;   // stream_kernel_launch: the cooperative launch was refused; issue the stream one launch per query
constexpr int STREAM_CTL_WORDS = 8;   // persistent stream kernel: [0,1] work counter (u64), [2] done, [3] fault, [4,5] tickets

}  // namespace sema_impl

struct sema_index {
    int device = 0;
    uint32_t dim = 0, ld = 0;
    uint64_t capacity = 0;
    int metric = 0;
    int num_sms = 0;
    float *X = nullptr;
    uint8_t *valid = nullptr;
    // growable index (sema_index_create_growable): X, valid and the K3 planes are address-space
    // reservations of the maximum size, backed by physical memory as rows arrive
    bool growable = false;
    sema_impl::GrowBuf gX, gValid, gPlanes;
    uint64_t n_rows = 0;     // appended (enqueued)
    uint64_t n_visible = 0;  // ingest completed on the device
    uint64_t last_snapshot = 0;
    uint32_t row_base = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, ingest_stream = nullptr;
    float *q_dev = nullptr;           // ld floats
    float *q_pin = nullptr;           // pinned, ld floats
    uint64_t *partials = nullptr;     // num_sms * MAX_BLOCKS_PER_SM * 128 keys
    unsigned int *ticket = nullptr;   // [0] arrival ticket, [2], [3] tile-claim counters of the TMA scan (alternating)
    uint64_t scan_seq = 0;            // TMA scans launched (selects the claim counter)
    unsigned int *stream_ctl = nullptr;   // control words of the persistent stream kernel (STREAM_CTL_WORDS), zeroed before every launch
    int stream_mode = 0;              // query streams: 0 = one persistent launch when a scan is long enough, 1 = whenever the shape allows, 2 = always one launch per query
    unsigned char *res_map = nullptr; // mapped pinned host memory the host-query path writes results to (RES_MAP_BYTES)
    unsigned char *res_map_dev = nullptr;   // its device address
    uint64_t host_seq = 0;            // tickets issued so far; ticket t uses slot t % RES_SLOTS and stores t to its flag
    struct Slot {                     // one submitted host search (sema_index_search_submit)
        uint64_t ticket = 0;          // 0 = free
        uint32_t k = 0;
        bool sync_done = false;       // shape outside the fast path: the search ran synchronously at submit
        uint32_t nf = 0;
        std::vector<uint64_t> ids;
        std::vector<float> sc;
    } slots[sema_impl::RES_SLOTS];
    int chain = 1;                    // 0 = query streams never chain consecutive scans with PDL (tuning / comparison)
    int host_path = 1;                // 0 = host searches always stage through q_dev / res_dev (tuning / comparison)
    uint64_t *keys_dev = nullptr;     // SEMA_MAX_K keys (multi-pass scratch)
    unsigned char *res_dev = nullptr; // [n_found u32, pad][ids u64 K][scores f32 K]
    unsigned char *res_pin = nullptr;
    // batch buffers, grown on demand
    float *Q_dev = nullptr;
    uint64_t *bids_dev = nullptr;
    float *bsc_dev = nullptr;
    uint32_t *bnf_dev = nullptr;
    size_t batch_cap_q = 0, batch_cap_res = 0, batch_cap_nf = 0;
    uint64_t *tomb_dev = nullptr;
    size_t tomb_cap = 0;
    // compaction scratch, kept between calls (cudaMalloc / cudaFree cost 1 - 20 ms per call on the bench boxes, more
    // than the compaction itself): keep flags, compacted validity bytes, gather list, tile counts + offsets,
    // old -> new map, [live rows, first dropped row], bounce buffer of one gather chunk
    struct CompactScratch {
        uint8_t *flags = nullptr, *new_valid = nullptr;
        uint32_t *src = nullptr, *tiles = nullptr;
        unsigned long long *map = nullptr, *head = nullptr;
        float *tmp = nullptr;
        size_t flags_cap = 0, new_valid_cap = 0, src_cap = 0, tiles_cap = 0, map_cap = 0, head_cap = 0, tmp_cap = 0;
    } cs;
    std::deque<sema_impl::Pending> pending;
    int variant = 0;
    uint64_t launches = 0;
    // ---- K3 (batched tensor-core path) state
    float *max_norm2 = nullptr;         // device: [max, min] squared row norm over the rows K1 has seen
    unsigned char *planes = nullptr;    // pre-tiled bf16 hi/lo planes, built lazily
    uint64_t planes_rows = 0;           // rows [0, planes_rows) are reflected in the planes
    bool planes_failed = false;         // allocation failed once: stay on the K2 loop
    bool poisoned = false;              // a compaction failed after rows had started to move: every later call is refused
    float *Qpad_dev = nullptr;
    uint32_t *cand_rows = nullptr;
    float *cand_thr = nullptr;
    float *cand_sc = nullptr;
    size_t cand_sc_cap = 0;
    uint32_t *flags_dev = nullptr, *flags_pin = nullptr;
    size_t qpad_cap = 0, cand_cap = 0, thr_cap = 0, flags_cap = 0;
    int batch_mode = 0;                 // 0 auto, 1 always the K2 loop, 2 K3 bf16x3 whenever the shape allows, 3 K3 single bf16 pass
    int k3_cluster = 0;                 // 0 auto, else forced cluster size (1, 2, 4) — tuning
    int k3_debug = 0;                   // timing experiments only (wrong results): see k3::Params::debug
    int k3_kc16 = 1;                    // single-pass stage with k <= 10 keeps 16 candidates per list (0 = 32) — tuning
    int k3_qt = 0;                      // 0 auto, 1 = one query tile per CTA even in the single-pass mode — tuning
    int k3_pair = 0;                    // single-pass stage: 0 = the single-CTA kernel (two query tiles per CTA, clusters of 2: 5.64 ms on config 3), 1 = CTA pairs (cta_group::2: 6.22 ms) — tuning
    int k3_mixed = 1;                   // a K3 stage as two concurrent launches, clusters of 4 + clusters of 2 on the SMs left over (0 = one launch) — tuning
    int k3_mix_w = 0;                   // mixed launch: row weight of a 4-cluster partition in percent above 1.00 (0 = built-in default) — tuning
    cudaStream_t aux_stream = nullptr;  // second stream of the mixed launch, with its fork / join events
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int k3_prefetch = 0;                // L2 prefetch distance of K3's producer in pipeline stages (0 = none) — tuning
    int k3_prec = 0;                    // plane format preference: 0 = automatic (fp16 when every |x_i| <= 1024, else bf16), 1 = bf16 always
    int planes_fmt = 0;                 // format the planes are in now (k3::FMT_BF16 / FMT_FP16)
    int planes_prec = 0;                // the preference they were built under
    CUtensorMap planes_tmap;            // the planes as a 3-D tensor for the CTA-pair kernel's cta_group::2 loads (k3_pair.cuh)
    const void *tmap_base = nullptr;    // planes pointer the map was encoded for (nullptr = none yet)
    bool l2_norms_constant = false;     // L2 metric: row norms nearly constant, so K3's dot-product selection ranks like the distance
    int normalize_queries = 0;          // apply K1 to host queries before scanning
    unsigned char *qscratch = nullptr;  // [valid byte x MAXQ pad][float max_norm2 scratch][K0 text counters: +8 query stream, +16 ingest stream]
    uint64_t k3_queries = 0, k3_fallbacks = 0, k3_cascaded = 0;
    float *sub_q = nullptr;             // cascade: the queries the single-pass stage could not prove, and their results
    uint64_t *sub_ids = nullptr;
    float *sub_sc = nullptr;
    uint32_t *sub_nf = nullptr;
    size_t sub_q_cap = 0, sub_ids_cap = 0, sub_sc_cap = 0, sub_nf_cap = 0;
    uint32_t *sub_idx = nullptr;        // cascade: batch positions of those queries
    size_t sub_idx_cap = 0;
    float *q_aligned = nullptr;         // 16-byte aligned copy of caller queries that are not
    size_t q_aligned_cap = 0;
};

namespace sema {
struct Exchange;   // k2_scan.cuh
}

namespace sema_impl {

// api_core.cu
int poll_ingest(sema_index *s, bool wait);            // advance n_visible past completed ingests
int ensure(void **p, size_t *cap, size_t need);       // grow a device scratch buffer
int normalize_queries_dev(sema_index *s, float *q, uint64_t stride, uint32_t nq);   // K1 on queries, in place
// api_pool.cu: kernel K0 (mean pooling + normalisation) for n texts on `stream`; out row stride out_ld floats
int launch_pool(sema_index *s, cudaStream_t stream, const float *tokens_dev, const float *mask_dev, uint64_t n,
                uint32_t seq_len, int skip_masked, float *out_dev, uint64_t out_ld);
// api_search.cu: best k (any k <= SEMA_MAX_K) for one device-resident query; out_keys or res_* may be null
int scan_query(sema_index *s, const float *q_dev, uint32_t n, uint32_t k, uint64_t *out_keys, uint64_t *res_ids,
               float *res_scores, uint32_t *res_nfound, const sema::Exchange *x = nullptr, unsigned flags = 0,
               uint64_t host_ticket = 0);
// Host-query fast path (single fused pass on the TMA kernel): the query rides in the kernel
// parameters and the last block writes the result block + a completion flag straight into mapped
// host memory, so a search costs one launch and no copies.  host_query_ok: does this handle /
// k take that path?  host_query_run: launch (x = optional shard exchange) and wait; fills the outputs.
bool host_query_ok(const sema_index *s, uint32_t k);
// launch for ticket `ticket` (results + flag go to slot ticket % RES_SLOTS); does not wait
int host_query_launch(sema_index *s, const float *q_host, uint32_t n, uint32_t k, const sema::Exchange *x, uint64_t ticket);
// wait for the ticket's flag and copy the slot out
int host_query_wait(sema_index *s, uint64_t ticket, uint32_t k, uint64_t *row_ids, float *scores, uint32_t *n_found);
int host_query_run(sema_index *s, const float *q_host, uint32_t n, uint32_t k, const sema::Exchange *x,
                   uint64_t *row_ids, float *scores, uint32_t *n_found);
// the general host-buffer path: query staged through pinned memory + H2D, result D2H, stream sync
int staged_query_run(sema_index *s, const float *q_host, uint32_t n, uint32_t k, const sema::Exchange *x,
                     uint64_t *row_ids, float *scores, uint32_t *n_found);
// ticket bookkeeping shared by the index-level and the shard-group submit / collect entry points
int slot_claim(sema_index *s, uint32_t k, uint64_t *ticket);                    // next ticket; fails when its slot is still uncollected
int slot_collect(sema_index *s, uint64_t ticket, uint64_t *row_ids, float *scores, uint32_t *n_found);
// api_batch.cu: NaN-poison the listed local rows (device array, on the query stream) in K3's bf16 planes; no-op without planes
int k3_poison_rows(sema_index *s, const uint64_t *rows_dev, uint64_t n);
// api_batch.cu: nq device-resident queries (nq x dim dense); K3 when the shape allows, else K2 per query
int batch_core(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
               uint32_t *nf_d);
// api_search.cu: the whole stream as ONE persistent launch (k2_stream.cuh); x->seq = sequence number of query 0
bool stream_kernel_ok(const sema_index *s, uint32_t nq, uint32_t n, uint32_t k);
int stream_kernel_launch(sema_index *s, const float *Q, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
                         uint32_t *nf_d, const sema::Exchange *x);
// api_batch.cu: a stream of nq single-query scans (K2): one persistent launch, or one launch per query with
// consecutive launches chained (PDL); x / seq: optional fused shard exchange, *seq advanced once per query
int scan_stream(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
                uint32_t *nf_d, sema::Exchange *x, uint64_t *seq);

}  // namespace sema_impl
