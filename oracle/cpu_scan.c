/*
 * oracle/cpu_scan.c — CPU restatement of Sema's vector-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sema_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker and as the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (akshitsinha/sema) ships no tests, golden
 * vectors or fixtures for this path and cannot be built in this environment
 * (no cargo/rustc; the arithmetic lives in the un-vendored crates
 * lancedb 0.23.1 -> lance 1.0.1 -> lance-index 1.0.1 / lance-linalg 1.0.1,
 * Cargo.lock:4258-4259, 3755-3756, 4028-4029, 4133-4134).  This file restates
 * the published behaviour of that engine at the reference's own call sites:
 *
 *   - normalise          src/semantic/embeddings.rs:83-88
 *   - mean_pool          src/semantic/embeddings.rs:61-91
 *   - nullable vector column / row packing
 *                        src/storage/lance_indexer.rs:41-45, 59-76
 *   - flat exact KNN     src/storage/lance_indexer.rs:121-126
 *       (no distance_type(), no create_index() anywhere in the crate =>
 *        LanceDB default: brute-force scan, squared-L2 `_distance`,
 *        ascending, `limit` rows, null vectors skipped)
 *
 * Ordering rule of the oracle: best first; exact ties broken by the lower row
 * id (the upstream TopK leaves tie order unspecified; BASELINE.json tolerates
 * ties within 1e-5).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SEMA_ORACLE_METRIC_DOT 0 /* score = q.x, descending                    */
#define SEMA_ORACLE_METRIC_L2  1 /* _distance = sum (q-x)^2, ascending (LanceDB default) */

/* ------------------------------------------------------------------------- */
/* Normalise tail of mean_pool — src/semantic/embeddings.rs:83-88.            */
/*   let norm: f32 = pooled.iter().map(|x| x * x).sum::<f32>().sqrt();         */
/*   if norm > 0.0 { for val in &mut pooled { *val /= norm; } }                */
/* Sequential f32 sum, f32 sqrt, per-element divide, zero vector untouched.   */
/* ------------------------------------------------------------------------- */
void sema_oracle_normalize(float *rows, uint64_t n, uint32_t d)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) {
        float *x = rows + (uint64_t)r * d;
        volatile float acc = 0.0f; /* volatile: keep the sequential order, no reassociation */
        for (uint32_t j = 0; j < d; ++j) {
            float sq = x[j] * x[j];
            acc = acc + sq;
        }
        float norm = sqrtf(acc);
        if (norm > 0.0f) {
            for (uint32_t j = 0; j < d; ++j) x[j] = x[j] / norm;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* mean_pool — src/semantic/embeddings.rs:61-91 (the step that produces every  */
/* stored vector and every query vector):                                     */
/*   for i in 0..seq_len { mask_sum += mask[i];                                */
/*       for j in 0..hidden { pooled[j] += token_embeddings[[0,i,j]] * mask[i] } } */
/*   if mask_sum > 0.0 { pooled /= mask_sum }   then the normalise tail above. */
/* Sequential f32 sums over i (per column) and over j (norm), multiply then    */
/* add (no FMA: -ffp-contract=off).  n texts are independent.                  */
/* ------------------------------------------------------------------------- */
void sema_oracle_mean_pool(const float *tokens, const float *mask, uint64_t n, uint32_t seq,
                           uint32_t hidden, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)n; ++t) {
        const float *tok = tokens + (uint64_t)t * seq * hidden;
        const float *m = mask + (uint64_t)t * seq;
        float *pooled = out + (uint64_t)t * hidden;
        for (uint32_t j = 0; j < hidden; ++j) pooled[j] = 0.0f;
        volatile float mask_sum = 0.0f;
        for (uint32_t i = 0; i < seq; ++i) {
            const float mv = m[i];
            mask_sum = mask_sum + mv;
            for (uint32_t j = 0; j < hidden; ++j) {
                float p = tok[(uint64_t)i * hidden + j] * mv;
                pooled[j] = pooled[j] + p;
            }
        }
        const float ms = mask_sum;
        if (ms > 0.0f)
            for (uint32_t j = 0; j < hidden; ++j) pooled[j] = pooled[j] / ms;
    }
    sema_oracle_normalize(out, n, hidden);
}

/* ------------------------------------------------------------------------- */
/* Deterministic synthetic embeddings: bit-identical on the */
/* CPU and on the GPU because only integer ops and an exact int->f32 convert  */
/* are used.  value(seed,row,col) = (sum of four 16-bit fields of a           */
/* splitmix64 hash) - 131070, an integer in [-131070, 131070].                */
/* ------------------------------------------------------------------------- */
static inline uint64_t synth_hash(uint64_t seed, uint64_t row, uint64_t col)
{
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + row * 0xBF58476D1CE4E5B9ull +
                 col * 0x94D049BB133111EBull + 1ull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

static inline float synth_value(uint64_t seed, uint64_t row, uint64_t col)
{
    uint64_t h = synth_hash(seed, row, col);
    int32_t s = (int32_t)(h & 0xffff) + (int32_t)((h >> 16) & 0xffff) +
                (int32_t)((h >> 32) & 0xffff) + (int32_t)(h >> 48);
    return (float)(s - 131070);
}

void sema_oracle_synth(float *out, uint64_t seed, uint64_t row0, uint64_t n, uint32_t d)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) {
        float *x = out + (uint64_t)r * d;
        for (uint32_t j = 0; j < d; ++j) x[j] = synth_value(seed, row0 + (uint64_t)r, j);
    }
}

/* ------------------------------------------------------------------------- */
/* Row kernels.  UPSTREAM lance-linalg computes f32 distances with SIMD lanes  */
/* (lane-chunked partial sums, not a sequential sum); restated here as 16      */
/* independent partial sums folded pairwise at the end.                        */
/* ------------------------------------------------------------------------- */
#define LANES 16

static inline float fold16(const float *p)
{
    float a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = p[i] + p[i + 8];
    for (int i = 0; i < 4; ++i) b[i] = a[i] + a[i + 4];
    return (b[0] + b[2]) + (b[1] + b[3]);
}

static inline float row_dot(const float *x, const float *q, uint32_t d)
{
    float p[LANES] = {0};
    uint32_t j = 0;
    for (; j + LANES <= d; j += LANES)
        for (int l = 0; l < LANES; ++l) p[l] += x[j + l] * q[j + l];
    float s = fold16(p);
    for (; j < d; ++j) s += x[j] * q[j];
    return s;
}

static inline float row_l2sq(const float *x, const float *q, uint32_t d)
{
    float p[LANES] = {0};
    uint32_t j = 0;
    for (; j + LANES <= d; j += LANES)
        for (int l = 0; l < LANES; ++l) {
            float t = q[j + l] - x[j + l];
            p[l] += t * t;
        }
    float s = fold16(p);
    for (; j < d; ++j) {
        float t = q[j] - x[j];
        s += t * t;
    }
    return s;
}

/* Ranking key: larger is better.  DOT: the score.  L2: minus the distance.   */
typedef struct {
    float key;
    uint64_t id;
} cand_t;

/* a ranks strictly ahead of b */
static inline int better(cand_t a, cand_t b)
{
    return a.key > b.key || (a.key == b.key && a.id < b.id);
}

/* Bounded list kept sorted best-first (k is small: 10..1024). */
static inline void topk_push(cand_t *list, uint32_t *len, uint32_t k, cand_t c)
{
    if (*len == k) {
        if (!better(c, list[k - 1])) return;
    } else {
        ++*len;
    }
    uint32_t i = *len - 1;
    while (i > 0 && better(c, list[i - 1])) {
        list[i] = list[i - 1];
        --i;
    }
    list[i] = c;
}

/*
 * Flat exact k-NN — src/storage/lance_indexer.rs:121-126
 *   table.query().nearest_to(q)?.limit(k).execute()
 * X: n rows, row stride ld floats (ld >= d).  valid: NULL or one byte per row
 * (0 = null vector, skipped — lance_indexer.rs:41-45, 66-70).  Rows whose key
 * is NaN are skipped as well.  Output: ids (row index + id_base) and, for
 * DOT, the score; for L2, the squared distance.  Returns n_found <= k.
 */
uint32_t sema_oracle_scan(const float *X, uint64_t n, uint32_t d, uint64_t ld,
                          const uint8_t *valid, const float *q, uint32_t k,
                          int metric, uint64_t id_base, uint64_t *ids_out,
                          float *scores_out)
{
    if (k == 0 || n == 0) return 0;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    cand_t *lists = (cand_t *)malloc((size_t)nt * k * sizeof(cand_t));
    uint32_t *lens = (uint32_t *)calloc((size_t)nt, sizeof(uint32_t));

#pragma omp parallel num_threads(nt)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        cand_t *list = lists + (size_t)t * k;
        uint32_t len = 0;
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)n; ++r) {
            if (valid && !valid[r]) continue;
            const float *x = X + (uint64_t)r * ld;
            float key = (metric == SEMA_ORACLE_METRIC_L2) ? -row_l2sq(x, q, d) : row_dot(x, q, d);
            if (key != key) continue; /* NaN never ranks */
            key += 0.0f;              /* -0 -> +0 */
            cand_t c = {key, (uint64_t)r};
            topk_push(list, &len, k, c);
        }
        lens[t] = len;
    }

    cand_t *fin = (cand_t *)malloc((size_t)k * sizeof(cand_t));
    uint32_t flen = 0;
    for (int t = 0; t < nt; ++t)
        for (uint32_t i = 0; i < lens[t]; ++i) topk_push(fin, &flen, k, lists[(size_t)t * k + i]);

    for (uint32_t i = 0; i < flen; ++i) {
        ids_out[i] = fin[i].id + id_base;
        scores_out[i] = (metric == SEMA_ORACLE_METRIC_L2) ? -fin[i].key + 0.0f : fin[i].key;
    }
    free(fin);
    free(lists);
    free(lens);
    return flen;
}

/* nq independent queries (the same call site issued nq times). */
void sema_oracle_scan_batch(const float *X, uint64_t n, uint32_t d, uint64_t ld,
                            const uint8_t *valid, const float *Q, uint32_t nq, uint32_t k,
                            int metric, uint64_t id_base, uint64_t *ids_out, float *scores_out,
                            uint32_t *n_found)
{
    for (uint32_t i = 0; i < nq; ++i)
        n_found[i] = sema_oracle_scan(X, n, d, ld, valid, Q + (uint64_t)i * d, k, metric, id_base,
                                      ids_out + (uint64_t)i * k, scores_out + (uint64_t)i * k);
}

/*
 * Merge of per-shard ranked lists: global top-k = top-k of
 * the union of per-shard top-k lists.  lists: G x k candidates (key, global id),
 * lens[g] valid entries each.
 */
uint32_t sema_oracle_merge(const float *scores, const uint64_t *ids, const uint32_t *lens,
                           uint32_t G, uint32_t k, int metric, uint64_t *ids_out,
                           float *scores_out)
{
    cand_t *fin = (cand_t *)malloc((size_t)(k ? k : 1) * sizeof(cand_t));
    uint32_t flen = 0;
    for (uint32_t g = 0; g < G; ++g)
        for (uint32_t i = 0; i < lens[g]; ++i) {
            float s = scores[(size_t)g * k + i];
            cand_t c = {(metric == SEMA_ORACLE_METRIC_L2) ? -s : s, ids[(size_t)g * k + i]};
            if (k) topk_push(fin, &flen, k, c);
        }
    for (uint32_t i = 0; i < flen; ++i) {
        ids_out[i] = fin[i].id;
        scores_out[i] = (metric == SEMA_ORACLE_METRIC_L2) ? -fin[i].key + 0.0f : fin[i].key;
    }
    free(fin);
    return flen;
}

int sema_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void sema_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
