/* abi_driver.c — a plain C99 caller of libfvdb_b200.so through include/fvdb.h alone: what the
 * Rust shim of INTEGRATION.md does, without any Python in between.
 *
 *   gcc -std=c99 -O1 -ffp-contract=off -I include tests/c/abi_driver.c -L fabstir_vectordb_b200 -lfvdb_b200 \
 *       -Wl,-rpath,$PWD/fabstir_vectordb_b200 -lm -o /tmp/abi_driver && /tmp/abi_driver
 *
 * Without a GPU the library must refuse to exist (FVDB_ERR_NO_DEVICE: there is no CPU fallback) — exit 0,
 * "no-device".  With one: centroids, IVF + recent-tier inserts, a tombstone, a hybrid search, the 3x
 * post-filter, a cosine handle; every result is compared bit for bit with the scalar loops below, which
 * restate euclidean_distance_scalar / cosine_similarity_scalar (src/core/vector_ops.rs:39-57). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fvdb.h"

#define D 8
#define N_IVF 200
#define N_FLAT 40
#define NLIST 4
#define NQ 5
#define K 6

static float l2(const float *a, const float *b) {
    float acc = 0.0f;
    for (int i = 0; i < D; ++i) { float t = a[i] - b[i]; acc = acc + t * t; }
    return sqrtf(acc);
}
static float dotp(const float *a, const float *b) {
    float acc = 0.0f;
    for (int i = 0; i < D; ++i) acc = acc + a[i] * b[i];
    return acc;
}
static float cosine(const float *a, const float *b) {
    float d = dotp(a, b), na = sqrtf(dotp(a, a)), nb = sqrtf(dotp(b, b));
    return (na == 0.0f || nb == 0.0f) ? 0.0f : d / (na * nb);
}
static unsigned lcg(unsigned *s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }
static float unit(unsigned *s) { return (float)(lcg(s) % 2001) / 1000.0f - 1.0f; }

#define CHECK(call) do { int rc_ = (call); if (rc_ != FVDB_OK) { \
    fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, fvdb_last_error(h)); return 1; } } while (0)

int main(void) {
    fvdb_index *h = NULL;
    int rc = fvdb_create(0, D, FVDB_METRIC_L2, 32, &h);
    if (rc == FVDB_ERR_NO_DEVICE) {
        printf("no-device: %s\n", fvdb_last_error(NULL));
        return 0;
    }
    if (rc != FVDB_OK) { fprintf(stderr, "fvdb_create -> %d (%s)\n", rc, fvdb_last_error(NULL)); return 1; }
    if (fvdb_abi_version() != FVDB_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }

    static float x[(N_IVF + N_FLAT) * D], q[NQ * D], cent[NLIST * D];
    static uint32_t ids[N_IVF + N_FLAT], lists[N_IVF];
    unsigned seed = 12345u;
    for (int i = 0; i < (N_IVF + N_FLAT) * D; ++i) x[i] = unit(&seed);
    for (int i = 0; i < NQ * D; ++i) q[i] = unit(&seed);
    for (int i = 0; i < N_IVF + N_FLAT; ++i) ids[i] = (uint32_t)i;
    memcpy(cent, x, sizeof(cent));                       /* the first rows as centroids */

    CHECK(fvdb_ivf_set_centroids(h, cent, NLIST));
    CHECK(fvdb_ivf_add(h, x, ids, N_IVF, lists));
    CHECK(fvdb_flat_add(h, x + N_IVF * D, ids + N_IVF, N_FLAT));
    for (int i = 0; i < N_IVF; ++i) {                    /* find_nearest_centroid: strict '<', lowest id wins */
        uint32_t best = 0; float bd = INFINITY;
        for (uint32_t c = 0; c < NLIST; ++c) { float d = l2(x + i * D, cent + c * D); if (d < bd) { bd = d; best = c; } }
        if (lists[i] != best) { fprintf(stderr, "row %d assigned to %u, expected %u\n", i, lists[i], best); return 1; }
    }
    if (fvdb_ivf_add(h, x, ids, 1, NULL) != FVDB_ERR_DUPLICATE) { fprintf(stderr, "duplicate id not rejected\n"); return 1; }
    uint32_t dead = 7;
    CHECK(fvdb_set_deleted(h, &dead, 1, 1));

    /* hybrid search, every list probed: equals a brute-force scan of both tiers minus the tombstone,
     * sorted by (distance, recent tier first on ties — none here —, id) */
    uint32_t out_ids[NQ * K], out_cnt[NQ];
    float out_dist[NQ * K];
    CHECK(fvdb_search(h, q, NQ, K, NLIST, FVDB_TIER_BOTH, NULL, 0, out_ids, out_dist, out_cnt));
    for (int qi = 0; qi < NQ; ++qi) {
        float best_d[K]; uint32_t best_i[K]; int n = 0;
        for (int r = 0; r < N_IVF + N_FLAT; ++r) {
            if ((uint32_t)r == dead) continue;
            float d = l2(q + qi * D, x + r * D);
            int p = n < K ? n : K;
            while (p > 0 && (best_d[p - 1] > d)) { if (p < K) { best_d[p] = best_d[p - 1]; best_i[p] = best_i[p - 1]; } --p; }
            if (p < K) { best_d[p] = d; best_i[p] = (uint32_t)r; if (n < K) ++n; }
        }
        if (out_cnt[qi] != (uint32_t)K) { fprintf(stderr, "query %d: %u results\n", qi, out_cnt[qi]); return 1; }
        for (int j = 0; j < K; ++j)
            if (out_ids[qi * K + j] != best_i[j] || memcmp(&out_dist[qi * K + j], &best_d[j], 4) != 0) {
                fprintf(stderr, "query %d rank %d: got (%u, %.9g) expected (%u, %.9g)\n", qi, j, out_ids[qi * K + j],
                        out_dist[qi * K + j], best_i[j], best_d[j]);
                return 1;
            }
    }

    /* 3x post-filter (HybridIndex::search_with_filter): keep even row ids among the 3k nearest, truncate to k */
    uint64_t keep[(N_IVF + N_FLAT + 63) / 64];
    memset(keep, 0, sizeof(keep));
    for (int r = 0; r < N_IVF + N_FLAT; r += 2) keep[r >> 6] |= 1ull << (r & 63);
    uint32_t p_ids[NQ * 2], p_cnt[NQ], w_ids[NQ * 6], w_cnt[NQ];
    float p_dist[NQ * 2], w_dist[NQ * 6];
    CHECK(fvdb_search_postfilter(h, q, NQ, 2, NLIST, FVDB_TIER_BOTH, keep, N_IVF + N_FLAT, p_ids, p_dist, p_cnt));
    CHECK(fvdb_search(h, q, NQ, 6, NLIST, FVDB_TIER_BOTH, NULL, 0, w_ids, w_dist, w_cnt));
    for (int qi = 0; qi < NQ; ++qi) {
        uint32_t n = 0;
        for (uint32_t j = 0; j < w_cnt[qi] && n < 2; ++j)
            if (!(w_ids[qi * 6 + j] & 1u)) {
                if (n >= p_cnt[qi] || p_ids[qi * 2 + n] != w_ids[qi * 6 + j]) { fprintf(stderr, "post-filter mismatch, query %d\n", qi); return 1; }
                ++n;
            }
        if (n != p_cnt[qi]) { fprintf(stderr, "post-filter count mismatch, query %d\n", qi); return 1; }
    }
    fvdb_stats st;
    CHECK(fvdb_get_stats(h, &st));
    if (st.ivf_rows != N_IVF || st.flat_rows != N_FLAT || st.deleted_rows != 1 || st.nlist != NLIST) { fprintf(stderr, "stats\n"); return 1; }
    fvdb_destroy(h);

    /* a cosine handle: the flat tier scored exhaustively, best (largest) first */
    CHECK(fvdb_create(0, D, FVDB_METRIC_COS, 32, &h));
    CHECK(fvdb_flat_add(h, x, ids, N_IVF));
    if (fvdb_ivf_set_centroids(h, cent, NLIST) != FVDB_ERR_INVALID_CONFIG) { fprintf(stderr, "IVF on a cosine handle not rejected\n"); return 1; }
    CHECK(fvdb_search(h, q, NQ, 1, 0, FVDB_TIER_RECENT, NULL, 0, out_ids, out_dist, out_cnt));
    for (int qi = 0; qi < NQ; ++qi) {
        float bs = -INFINITY; uint32_t bi = 0;
        for (int r = 0; r < N_IVF; ++r) { float s = cosine(q + qi * D, x + r * D); if (s > bs) { bs = s; bi = (uint32_t)r; } }
        if (out_ids[qi] != bi || memcmp(&out_dist[qi], &bs, 4) != 0) { fprintf(stderr, "cosine query %d: (%u, %.9g) vs (%u, %.9g)\n", qi, out_ids[qi], out_dist[qi], bi, bs); return 1; }
    }
    fvdb_destroy(h);
    printf("abi-driver ok\n");
    return 0;
}
