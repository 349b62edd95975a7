/* abi_smoke.c — uses the drop-in boundary from plain C, the way the reference's Rust FFI would:
 * include/sema_b200.h (inner seam) and include/sema_store.h (outer seam).  No Python, no torch.
 * Exit codes: 0 ok, 3 no CUDA device (the library refuses to run: there is no CPU fallback),
 * 1 wrong result / unexpected error. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sema_store.h"

#define DIM 384
#define N 5000
#define CHECK(call)                                                                    \
    do {                                                                               \
        int rc_ = (call);                                                              \
        if (rc_ != SEMA_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, sema_last_error());          \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

static unsigned long long s = 88172645463325252ull;
static float rnd(void)
{
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (float)((double)(s >> 11) / 9007199254740992.0) - 0.5f;
}

static int embed(void *user, const char *text, float *out, uint32_t dim)
{
    (void)user;
    if (text[0] == '?') return 1; /* "embedding failed" */
    unsigned h = 2166136261u;
    for (const char *p = text; *p; ++p) h = (h ^ (unsigned char)*p) * 16777619u;
    for (uint32_t j = 0; j < dim; ++j) {
        h = h * 1664525u + 1013904223u;
        out[j] = (float)(h >> 8) / 16777216.0f - 0.5f;
    }
    return 0;
}

int main(void)
{
    if (sema_device_count() == 0) {
        sema_index *none = NULL;
        int rc = sema_index_create(0, DIM, 16, SEMA_METRIC_COSINE, &none);
        printf("no CUDA device: sema_index_create -> %d (%s)\n", rc, sema_last_error());
        return (rc == SEMA_ERR_CUDA && none == NULL) ? 3 : 1;
    }
    /* ---- inner seam ---- */
    float *rows = malloc(sizeof(float) * N * DIM);
    for (size_t i = 0; i < (size_t)N * DIM; ++i) rows[i] = rnd();
    sema_index *idx = NULL;
    CHECK(sema_index_create(0, DIM, N, SEMA_METRIC_COSINE, &idx));
    uint64_t first = 99;
    CHECK(sema_index_append(idx, rows, N, NULL, /*normalize=*/1, &first));
    if (first != 0 || sema_index_size(idx) != N) return 1;
    uint64_t ids[10];
    float sc[10];
    uint32_t nf = 0;
    CHECK(sema_index_set_normalize_queries(idx, 1) == 1 ? SEMA_OK : SEMA_ERR_INVALID);
    CHECK(sema_index_search(idx, rows + (size_t)1234 * DIM, 10, ids, sc, &nf)); /* a stored row finds itself */
    if (nf != 10 || ids[0] != 1234 || fabsf(sc[0] - 1.0f) > 1e-5f) {
        fprintf(stderr, "unexpected top hit %llu %.6f (n_found %u)\n", (unsigned long long)ids[0], sc[0], nf);
        return 1;
    }
    for (int i = 1; i < 10; ++i)
        if (sc[i] > sc[i - 1]) return 1;
    /* the host-query path (query in the kernel parameters, results through mapped host memory):
     * an already normalised stored row, several calls in a row */
    static float qn[DIM];
    CHECK(sema_index_read_rows(idx, 77, 1, qn));
    CHECK(sema_index_set_normalize_queries(idx, 0) == 0 ? SEMA_OK : SEMA_ERR_INVALID);
    for (int rep = 0; rep < 3; ++rep) {
        CHECK(sema_index_search(idx, qn, 10, ids, sc, &nf));
        if (nf != 10 || ids[0] != 77 || fabsf(sc[0] - 1.0f) > 1e-5f) return 1;
    }
    /* asynchronous form: two searches in flight, collected in order */
    {
        uint64_t t0 = 0, t1 = 0;
        static float q2[DIM];
        CHECK(sema_index_read_rows(idx, 4321, 1, q2));
        CHECK(sema_index_search_submit(idx, qn, 10, &t0));
        CHECK(sema_index_search_submit(idx, q2, 10, &t1));
        CHECK(sema_index_search_collect(idx, t0, ids, sc, &nf));
        if (nf != 10 || ids[0] != 77) return 1;
        CHECK(sema_index_search_collect(idx, t1, ids, sc, &nf));
        if (nf != 10 || ids[0] != 4321) return 1;
        if (sema_index_search_collect(idx, t1, ids, sc, &nf) != SEMA_ERR_INVALID) return 1; /* already collected */
    }
    /* mean_pool (kernel K0, src/semantic/embeddings.rs:61-91): 3 tokens, the last one padding */
    {
        static float tok[3 * DIM], pooled[DIM], want[DIM];
        const float mask[3] = {1.0f, 1.0f, 0.0f};
        for (int i = 0; i < 3 * DIM; ++i) tok[i] = rnd();
        CHECK(sema_mean_pool(idx, tok, mask, 1, 3, /*skip_masked=*/1, pooled));
        volatile float ss = 0.0f;
        for (int j = 0; j < DIM; ++j) {
            volatile float acc = 0.0f;
            for (int i = 0; i < 3; ++i) { volatile float pr = tok[i * DIM + j] * mask[i]; acc = acc + pr; }
            want[j] = acc / 2.0f;
        }
        for (int j = 0; j < DIM; ++j) { volatile float sq = want[j] * want[j]; ss = ss + sq; }
        const float norm = sqrtf(ss);
        for (int j = 0; j < DIM; ++j)
            if (pooled[j] != want[j] / norm) {
                fprintf(stderr, "mean_pool[%d] = %.9g, expected %.9g\n", j, pooled[j], want[j] / norm);
                return 1;
            }
    }
    uint64_t dead = 1234;
    CHECK(sema_index_tombstone(idx, &dead, 1));
    CHECK(sema_index_search(idx, rows + (size_t)1234 * DIM, 10, ids, sc, &nf));
    if (ids[0] == 1234) return 1;
    CHECK(sema_index_destroy(idx));

    /* ---- outer seam: StorageManager::search(&str, limit) -> Vec<(Chunk, f32)> ---- */
    sema_store *st = NULL;
    CHECK(sema_store_create(0, DIM, 64, /*normalize=*/1, &st));
    CHECK(sema_store_set_embedder(st, embed, NULL));
    const char *cid[] = {"a.md:0", "a.md:1", "b.md:0", "b.md:1"};
    const char *path[] = {"a.md", "a.md", "b.md", "b.md"};
    const uint64_t sl[] = {1, 20, 1, 30}, el[] = {19, 40, 29, 50};
    const char *content[] = {"vector index", "exact scan", "? unembeddable", "top k merge"};
    CHECK(sema_store_index_chunks_embed(st, 4, cid, path, sl, el, content));
    sema_hit hits[8];
    CHECK(sema_store_search(st, "  exact scan \n", 8, hits, &nf)); /* trimmed; identical text -> cosine 1 */
    const char *id = NULL, *fp = NULL, *ct = NULL;
    uint64_t a = 0, b = 0;
    CHECK(sema_store_chunk(st, hits[0].row, &id, &fp, &a, &b, &ct));
    if (nf != 3 || strcmp(id, "a.md:1") != 0 || fabsf(hits[0].score - 1.0f) > 1e-5f || a != 20 || b != 40) {
        fprintf(stderr, "store search: n=%u top=%s score=%.6f\n", nf, id, hits[0].score);
        return 1;
    }
    CHECK(sema_store_search(st, "? unembeddable", 8, hits, &nf)); /* embedding fails -> LIKE fallback */
    if (nf != 1 || hits[0].row != 2 || hits[0].score != 1.0f) return 1;
    if (sema_store_search(st, "'keyword", 8, hits, &nf) != SEMA_ERR_UNSUPPORTED) return 1;
    sema_search_result grouped[SEMA_SEARCH_RESULTS_LIMIT];
    uint32_t ng = 0;
    CHECK(sema_store_execute_search(st, "exact scan", grouped, SEMA_SEARCH_RESULTS_LIMIT, &ng));
    if (ng != 2 || grouped[0].total_matches_in_file + grouped[1].total_matches_in_file != 3) return 1;
    uint64_t removed = 0;
    CHECK(sema_store_remove_file_chunks(st, "a.md", &removed));
    if (removed != 2) return 1;
    CHECK(sema_store_search(st, "exact scan", 8, hits, &nf));
    if (nf != 1 || hits[0].row != 3) return 1;
    CHECK(sema_store_destroy(st));
    free(rows);
    printf("abi_smoke ok\n");
    return 0;
}
