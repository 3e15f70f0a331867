/*
 * fvdb_oracle.c — CPU restatement of fabstir-vectordb's search / training hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fabstir_vectordb_b200/ may import, link or call
 * this file; it is the checker for tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * The Rust reference cannot be compiled here (no cargo/rustc in the image, no network), so
 * this file restates its arithmetic in plain C.  Build with
 *     gcc -O2 -ffp-contract=off -fno-fast-math     (see oracle/Makefile)
 * Rust never contracts a*b+c to FMA and never re-associates float sums; `.powi(2)` is x*x and
 * `.sum::<f32>()` is a left fold, so plain C loops compiled without contraction are
 * bit-identical on x86-64.  Every function cites the reference lines it follows (paths
 * relative to the reference repository root).
 *
 * Parity status: distance / argmin / coarse ranking / list scan / merge / Lloyd iteration /
 * top-k helpers are pinned by the reference's own known-answer tests (tests/test_oracle_kat.py
 * lists them with file:line).  Seeded k-means++ (rand 0.8 StdRng = ChaCha12, an un-vendored,
 * un-pinned dependency) is restated from its published algorithm: PARITY UNPINNED for the
 * seeded random stream; Lloyd parity is tested from a shared initial centroid set instead.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define FO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Scalar kernels — src/core/vector_ops.rs
 * ---------------------------------------------------------------------------------------- */

/* euclidean_distance_scalar, src/core/vector_ops.rs:51-57 (twin: src/hnsw/core.rs:691-697):
 * zip -> (x-y).powi(2) -> sum::<f32>() -> sqrt(). */
FO_EXPORT float fo_l2(const float *a, const float *b, size_t d) {
    float acc = 0.0f;
    for (size_t i = 0; i < d; ++i) {
        float t = a[i] - b[i];
        acc = acc + t * t;
    }
    return sqrtf(acc);
}

/* dot_product_scalar, src/core/vector_ops.rs:35-37. */
FO_EXPORT float fo_dot(const float *a, const float *b, size_t d) {
    float acc = 0.0f;
    for (size_t i = 0; i < d; ++i) acc = acc + a[i] * b[i];
    return acc;
}

/* cosine_similarity_scalar, src/core/vector_ops.rs:39-49. */
FO_EXPORT float fo_cosine(const float *a, const float *b, size_t d) {
    float dot = fo_dot(a, b, d);
    float na = sqrtf(fo_dot(a, a, d));
    float nb = sqrtf(fo_dot(b, b, d));
    if (na == 0.0f || nb == 0.0f) return 0.0f;
    return dot / (na * nb);
}

/* 8 independent (query,row) chains interleaved: the same per-pair operation order as fo_l2
 * (so the same bits), but the add latency of one chain hides behind the others.  This is the
 * "tight" CPU baseline's inner loop; rows are row-major [n x d]. */
static void l2_block8(const float *q, const float *rows, size_t d, float out[8]) {
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
    const float *r0 = rows, *r1 = rows + d, *r2 = rows + 2 * d, *r3 = rows + 3 * d;
    const float *r4 = rows + 4 * d, *r5 = rows + 5 * d, *r6 = rows + 6 * d, *r7 = rows + 7 * d;
    for (size_t i = 0; i < d; ++i) {
        float qi = q[i];
        float t0 = qi - r0[i], t1 = qi - r1[i], t2 = qi - r2[i], t3 = qi - r3[i];
        float t4 = qi - r4[i], t5 = qi - r5[i], t6 = qi - r6[i], t7 = qi - r7[i];
        a0 = a0 + t0 * t0; a1 = a1 + t1 * t1; a2 = a2 + t2 * t2; a3 = a3 + t3 * t3;
        a4 = a4 + t4 * t4; a5 = a5 + t5 * t5; a6 = a6 + t6 * t6; a7 = a7 + t7 * t7;
    }
    out[0] = sqrtf(a0); out[1] = sqrtf(a1); out[2] = sqrtf(a2); out[3] = sqrtf(a3);
    out[4] = sqrtf(a4); out[5] = sqrtf(a5); out[6] = sqrtf(a6); out[7] = sqrtf(a7);
}

/* distances from q to n contiguous rows; bit-identical to n calls of fo_l2. */
FO_EXPORT void fo_l2_many(const float *q, const float *rows, size_t n, size_t d, float *out) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) l2_block8(q, rows + i * d, d, out + i);
    for (; i < n; ++i) out[i] = fo_l2(q, rows + i * d, d);
}

/* ------------------------------------------------------------------------------------------
 * Top-k helpers — src/core/vector_ops.rs:12-32,180-260 and src/core/types.rs:206-238
 * ---------------------------------------------------------------------------------------- */

typedef struct { float key; uint32_t idx; uint32_t aux; } fo_pair;

/* stable merge sort on .key ascending (Rust's slice::sort_by is a stable merge sort; only
 * stability matters for the result). */
static void stable_sort_pairs(fo_pair *a, size_t n, fo_pair *tmp) {
    if (n < 2) return;
    if (n <= 16) {
        for (size_t i = 1; i < n; ++i) {
            fo_pair v = a[i];
            size_t j = i;
            while (j > 0 && a[j - 1].key > v.key) { a[j] = a[j - 1]; --j; }
            a[j] = v;
        }
        return;
    }
    size_t h = n / 2;
    stable_sort_pairs(a, h, tmp);
    stable_sort_pairs(a + h, n - h, tmp);
    size_t i = 0, j = h, o = 0;
    while (i < h && j < n) tmp[o++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
    while (i < h) tmp[o++] = a[i++];
    while (j < n) tmp[o++] = a[j++];
    memcpy(a, tmp, n * sizeof(fo_pair));
}

/* top_k_indices, src/core/vector_ops.rs:12-22: stable sort DESCENDING by score, take k. */
FO_EXPORT size_t fo_top_k_indices(const float *scores, size_t n, size_t k, uint32_t *out) {
    fo_pair *a = (fo_pair *)malloc((n + 1) * sizeof(fo_pair));
    fo_pair *t = (fo_pair *)malloc((n + 1) * sizeof(fo_pair));
    for (size_t i = 0; i < n; ++i) { a[i].key = -scores[i]; a[i].idx = (uint32_t)i; a[i].aux = 0; }
    stable_sort_pairs(a, n, t);
    size_t m = k < n ? k : n;
    for (size_t i = 0; i < m; ++i) out[i] = a[i].idx;
    free(a); free(t);
    return m;
}

/* top_k_indices_heap, src/core/vector_ops.rs:180-201: bounded min-heap on score, replace the
 * minimum only on strict '>', then sort descending.  Restated with a linear-scan bounded set
 * (same admission rule => same retained multiset wherever scores are distinct). */
FO_EXPORT size_t fo_top_k_indices_heap(const float *scores, size_t n, size_t k, uint32_t *out) {
    if (k == 0) return 0;
    fo_pair *heap = (fo_pair *)malloc(k * sizeof(fo_pair));
    size_t len = 0;
    for (size_t i = 0; i < n; ++i) {
        if (len < k) { heap[len].key = scores[i]; heap[len].idx = (uint32_t)i; ++len; continue; }
        size_t mi = 0;
        for (size_t j = 1; j < len; ++j) if (heap[j].key < heap[mi].key) mi = j;
        if (scores[i] > heap[mi].key) { heap[mi].key = scores[i]; heap[mi].idx = (uint32_t)i; }
    }
    fo_pair *t = (fo_pair *)malloc((len + 1) * sizeof(fo_pair));
    for (size_t j = 0; j < len; ++j) heap[j].key = -heap[j].key;
    stable_sort_pairs(heap, len, t);
    for (size_t j = 0; j < len; ++j) out[j] = heap[j].idx;
    free(heap); free(t);
    return len;
}

/* StreamingTopK::{add,get_results}, src/core/vector_ops.rs:203-260: keeps the k largest
 * scores, returns them descending. */
FO_EXPORT size_t fo_streaming_top_k(const float *scores, const uint32_t *ids, size_t n, size_t k,
                                    uint32_t *out_ids, float *out_scores) {
    uint32_t *idx = (uint32_t *)malloc((k ? k : 1) * sizeof(uint32_t));
    size_t m = fo_top_k_indices_heap(scores, n, k, idx);
    for (size_t i = 0; i < m; ++i) { out_ids[i] = ids[idx[i]]; out_scores[i] = scores[idx[i]]; }
    free(idx);
    return m;
}

/* SearchResult::deduplicate + merge_search_results, src/core/types.rs:206-225 and
 * src/core/vector_ops.rs:24-32: keep the minimum distance per id, sort ascending, take k.
 * (ids, dist) is the concatenation of all result sets. */
FO_EXPORT size_t fo_merge_search_results(const uint32_t *ids, const float *dist, size_t n,
                                         size_t k, uint32_t *out_ids, float *out_dist) {
    fo_pair *a = (fo_pair *)malloc((n + 1) * sizeof(fo_pair));
    fo_pair *t = (fo_pair *)malloc((n + 1) * sizeof(fo_pair));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t j = 0;
        for (; j < m; ++j) if (a[j].idx == ids[i]) break;
        if (j == m) { a[m].idx = ids[i]; a[m].key = dist[i]; a[m].aux = 0; ++m; }
        else if (!(a[j].key <= dist[i])) a[j].key = dist[i];
    }
    stable_sort_pairs(a, m, t);
    size_t r = k < m ? k : m;
    for (size_t i = 0; i < r; ++i) { out_ids[i] = a[i].idx; out_dist[i] = a[i].key; }
    free(a); free(t);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Canonical ordering
 *
 * The reference sorts candidates by distance only, with a stable sort, so the order of equal
 * distances is the order of generation: cluster id for the coarse step (src/ivf/core.rs:655),
 * HashMap iteration order inside a posting list (:572-574; per-process random => undefined in
 * the reference itself), recent tier before IVF tier in the hybrid merge
 * (src/hybrid/core.rs:460,472,482).  The oracle fixes the undefined part as "lower row id
 * first": candidates are ranked by (distance, id), and the hybrid merge by
 * (distance, tier, id).  Wherever distances differ this is exactly the reference's order.
 * ---------------------------------------------------------------------------------------- */

static inline uint64_t key_of(float dist, uint32_t id) {
    uint32_t b;
    memcpy(&b, &dist, 4); /* dist >= +0: bit pattern is monotone */
    return ((uint64_t)b << 32) | id;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* keep the k smallest keys of a stream in a sorted array (insertion); O(n + hits*k). */
typedef struct { uint64_t *keys; size_t k, len; } fo_topk;
static inline void topk_push(fo_topk *t, uint64_t key) {
    if (t->len == t->k) {
        if (t->k == 0 || key >= t->keys[t->k - 1]) return;
        t->len--;
    }
    size_t j = t->len;
    while (j > 0 && t->keys[j - 1] > key) { t->keys[j] = t->keys[j - 1]; --j; }
    t->keys[j] = key;
    t->len++;
}

static inline int bit_get(const uint64_t *bits, uint64_t nbits, uint32_t id) {
    if (!bits) return 0;
    if (id >= nbits) return 0;
    return (int)((bits[id >> 6] >> (id & 63)) & 1u);
}

/* ------------------------------------------------------------------------------------------
 * IVF index — src/ivf/core.rs
 * ---------------------------------------------------------------------------------------- */

typedef struct fo_ivf {
    size_t dim, nlist;
    float *centroids;       /* [nlist x dim] */
    /* posting lists, grouped: rows of list l are [list_off[l], list_off[l+1]) */
    size_t n;
    size_t *list_off;       /* [nlist + 1] */
    float *rows;            /* [n x dim] */
    uint32_t *ids;          /* [n] */
} fo_ivf;

/* find_nearest_centroid, src/ivf/core.rs:373-386: strict '<', start (ClusterId(0), +inf). */
FO_EXPORT uint32_t fo_find_nearest_centroid(const float *x, const float *centroids, size_t nlist,
                                            size_t d) {
    uint32_t best = 0;
    float best_dist = INFINITY;
    for (size_t c = 0; c < nlist; ++c) {
        float dist = fo_l2(x, centroids + c * d, d);
        if (dist < best_dist) { best_dist = dist; best = (uint32_t)c; }
    }
    return best;
}

/* batch form (fvdb_assign); OpenMP over rows, each row independent => same bits. */
FO_EXPORT void fo_assign(const float *x, size_t n, const float *centroids, size_t nlist, size_t d,
                         uint32_t *out) {
#pragma omp parallel
    {
        float *dist = (float *)malloc((nlist + 8) * sizeof(float));
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)n; ++i) {
            fo_l2_many(x + (size_t)i * d, centroids, nlist, d, dist);
            uint32_t best = 0;
            float bd = INFINITY;
            for (size_t c = 0; c < nlist; ++c) if (dist[c] < bd) { bd = dist[c]; best = (uint32_t)c; }
            out[i] = best;
        }
        free(dist);
    }
}

/* IVFIndex::set_trained + repeated IVFIndex::insert (src/ivf/core.rs:431-455,509-520):
 * every row goes to its nearest centroid's list; rows keep insertion order inside a list. */
FO_EXPORT fo_ivf *fo_ivf_build(const float *centroids, size_t nlist, size_t dim, const float *x,
                               const uint32_t *ids, size_t n, uint32_t *out_assign) {
    fo_ivf *ix = (fo_ivf *)calloc(1, sizeof(fo_ivf));
    ix->dim = dim; ix->nlist = nlist; ix->n = n;
    ix->centroids = (float *)malloc(nlist * dim * sizeof(float));
    memcpy(ix->centroids, centroids, nlist * dim * sizeof(float));
    uint32_t *assign = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    fo_assign(x, n, centroids, nlist, dim, assign);
    ix->list_off = (size_t *)calloc(nlist + 1, sizeof(size_t));
    for (size_t i = 0; i < n; ++i) ix->list_off[assign[i] + 1]++;
    for (size_t l = 0; l < nlist; ++l) ix->list_off[l + 1] += ix->list_off[l];
    size_t *cur = (size_t *)malloc((nlist + 1) * sizeof(size_t));
    memcpy(cur, ix->list_off, (nlist + 1) * sizeof(size_t));
    ix->rows = (float *)malloc((n ? n : 1) * dim * sizeof(float));
    ix->ids = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    for (size_t i = 0; i < n; ++i) {
        size_t p = cur[assign[i]]++;
        memcpy(ix->rows + p * dim, x + i * dim, dim * sizeof(float));
        ix->ids[p] = ids[i];
    }
    if (out_assign) memcpy(out_assign, assign, n * sizeof(uint32_t));
    free(cur); free(assign);
    return ix;
}

/* build from precomputed assignments (used when the assignment itself came from the oracle
 * earlier, or to mirror a given device layout). */
FO_EXPORT fo_ivf *fo_ivf_build_assigned(const float *centroids, size_t nlist, size_t dim,
                                        const float *x, const uint32_t *ids,
                                        const uint32_t *assign, size_t n) {
    fo_ivf *ix = (fo_ivf *)calloc(1, sizeof(fo_ivf));
    ix->dim = dim; ix->nlist = nlist; ix->n = n;
    ix->centroids = (float *)malloc(nlist * dim * sizeof(float));
    memcpy(ix->centroids, centroids, nlist * dim * sizeof(float));
    ix->list_off = (size_t *)calloc(nlist + 1, sizeof(size_t));
    for (size_t i = 0; i < n; ++i) ix->list_off[assign[i] + 1]++;
    for (size_t l = 0; l < nlist; ++l) ix->list_off[l + 1] += ix->list_off[l];
    size_t *cur = (size_t *)malloc((nlist + 1) * sizeof(size_t));
    memcpy(cur, ix->list_off, (nlist + 1) * sizeof(size_t));
    ix->rows = (float *)malloc((n ? n : 1) * dim * sizeof(float));
    ix->ids = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    for (size_t i = 0; i < n; ++i) {
        size_t p = cur[assign[i]]++;
        memcpy(ix->rows + p * dim, x + i * dim, dim * sizeof(float));
        ix->ids[p] = ids[i];
    }
    free(cur);
    return ix;
}

FO_EXPORT void fo_ivf_free(fo_ivf *ix) {
    if (!ix) return;
    free(ix->centroids); free(ix->list_off); free(ix->rows); free(ix->ids); free(ix);
}

FO_EXPORT size_t fo_ivf_list_len(const fo_ivf *ix, size_t l) {
    return ix->list_off[l + 1] - ix->list_off[l];
}

/* coarse step of search_with_config, src/ivf/core.rs:646-656: all centroid distances, stable
 * ascending sort (lower cluster id first on ties), truncate(n_probe). Returns #probes. */
FO_EXPORT size_t fo_ivf_coarse(const fo_ivf *ix, const float *q, size_t nprobe, uint32_t *out_lists,
                               float *out_dist) {
    size_t nl = ix->nlist;
    fo_pair *a = (fo_pair *)malloc((nl + 1) * sizeof(fo_pair));
    fo_pair *t = (fo_pair *)malloc((nl + 1) * sizeof(fo_pair));
    float *dist = (float *)malloc((nl + 8) * sizeof(float));
    fo_l2_many(q, ix->centroids, nl, ix->dim, dist);
    for (size_t c = 0; c < nl; ++c) { a[c].key = dist[c]; a[c].idx = (uint32_t)c; a[c].aux = 0; }
    stable_sort_pairs(a, nl, t);
    size_t m = nprobe < nl ? nprobe : nl;
    for (size_t i = 0; i < m; ++i) { out_lists[i] = a[i].idx; if (out_dist) out_dist[i] = a[i].key; }
    free(a); free(t); free(dist);
    return m;
}

/* IVFIndex::search_with_config, src/ivf/core.rs:626-681, for one query.
 *   deleted : tombstone bitmap over ids (is_deleted, src/ivf/operations.rs:589; skip :667)
 *   filter  : optional PRE-filter bitmap over ids (bit set = passes); NULL = no filter.
 * Candidates ranked by (distance, id) — see "Canonical ordering".  Returns result count. */
FO_EXPORT size_t fo_ivf_search(const fo_ivf *ix, const float *q, size_t k, size_t nprobe,
                               const uint64_t *deleted, uint64_t deleted_nbits,
                               const uint64_t *filter, uint64_t filter_nbits, uint32_t *out_ids,
                               float *out_dist) {
    size_t nl = ix->nlist, d = ix->dim;
    size_t np = nprobe < nl ? nprobe : nl;
    uint32_t *lists = (uint32_t *)malloc((np + 1) * sizeof(uint32_t));
    np = fo_ivf_coarse(ix, q, np, lists, NULL);
    fo_topk tk;
    tk.k = k; tk.len = 0;
    tk.keys = (uint64_t *)malloc((k + 1) * sizeof(uint64_t));
    size_t maxlen = 0;
    for (size_t p = 0; p < np; ++p) {
        size_t L = fo_ivf_list_len(ix, lists[p]);
        if (L > maxlen) maxlen = L;
    }
    float *dist = (float *)malloc((maxlen + 8) * sizeof(float));
    for (size_t p = 0; p < np; ++p) {
        size_t b = ix->list_off[lists[p]], L = fo_ivf_list_len(ix, lists[p]);
        fo_l2_many(q, ix->rows + b * d, L, d, dist);
        for (size_t r = 0; r < L; ++r) {
            uint32_t id = ix->ids[b + r];
            if (bit_get(deleted, deleted_nbits, id)) continue;
            if (filter && !bit_get(filter, filter_nbits, id)) continue;
            topk_push(&tk, key_of(dist[r], id));
        }
    }
    for (size_t i = 0; i < tk.len; ++i) {
        uint32_t b = (uint32_t)(tk.keys[i] >> 32);
        out_ids[i] = (uint32_t)(tk.keys[i] & 0xffffffffu);
        memcpy(&out_dist[i], &b, 4);
    }
    size_t r = tk.len;
    free(tk.keys); free(dist); free(lists);
    return r;
}

/* "Faithful-cost" single-query search: same arithmetic and result as fo_ivf_search, but with
 * the reference's data movement — every row of a probed list is cloned into a freshly
 * allocated buffer first (get_cluster_vectors, src/ivf/core.rs:565-574), every candidate is
 * materialised, and ALL candidates are fully sorted (:677) before truncate(k).  Used only to
 * time what HybridIndex::search costs per query on one core. */
FO_EXPORT size_t fo_ivf_search_faithful(const fo_ivf *ix, const float *q, size_t k, size_t nprobe,
                                        const uint64_t *deleted, uint64_t deleted_nbits,
                                        uint32_t *out_ids, float *out_dist) {
    size_t nl = ix->nlist, d = ix->dim;
    size_t np = nprobe < nl ? nprobe : nl;
    uint32_t *lists = (uint32_t *)malloc((np + 1) * sizeof(uint32_t));
    np = fo_ivf_coarse(ix, q, np, lists, NULL);
    size_t cap = 1024, len = 0;
    uint64_t *keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    for (size_t p = 0; p < np; ++p) {
        size_t b = ix->list_off[lists[p]], L = fo_ivf_list_len(ix, lists[p]);
        float **clones = (float **)malloc((L + 1) * sizeof(float *));
        for (size_t r = 0; r < L; ++r) { /* clone per row, like Vec<f32>::clone */
            clones[r] = (float *)malloc(d * sizeof(float));
            memcpy(clones[r], ix->rows + (b + r) * d, d * sizeof(float));
        }
        for (size_t r = 0; r < L; ++r) {
            uint32_t id = ix->ids[b + r];
            if (!bit_get(deleted, deleted_nbits, id)) {
                if (len == cap) { cap *= 2; keys = (uint64_t *)realloc(keys, cap * sizeof(uint64_t)); }
                keys[len++] = key_of(fo_l2(q, clones[r], d), id);
            }
            free(clones[r]);
        }
        free(clones);
    }
    qsort(keys, len, sizeof(uint64_t), cmp_u64);
    size_t r = k < len ? k : len;
    for (size_t i = 0; i < r; ++i) {
        uint32_t b = (uint32_t)(keys[i] >> 32);
        out_ids[i] = (uint32_t)(keys[i] & 0xffffffffu);
        memcpy(&out_dist[i], &b, 4);
    }
    free(keys); free(lists);
    return r;
}

/* Exhaustive scoring with a similarity metric, the only way the reference uses its cosine / dot kernels:
 * batch_cosine_similarity (src/core/vector_ops.rs:8-10) over every vector, then top_k_indices (:12-23) —
 * a STABLE sort by descending score, first k (ties keep the order of the input).  metric: 1 = cosine
 * (cosine_similarity_scalar :39-49, the query is `a`), 2 = dot product (:35-37).  Deleted / filtered-out
 * rows are skipped as in the L2 searches.  Returns the number of results. */
typedef struct { float score; uint32_t id; size_t order; } fo_scored;
static int cmp_scored_desc(const void *pa, const void *pb) {
    const fo_scored *a = (const fo_scored *)pa, *b = (const fo_scored *)pb;
    if (a->score > b->score) return -1;     /* b.partial_cmp(a): larger score first */
    if (a->score < b->score) return 1;
    return a->order < b->order ? -1 : (a->order > b->order ? 1 : 0);   /* stable */
}
FO_EXPORT size_t fo_flat_search_metric(const float *rows, const uint32_t *ids, size_t n, size_t d,
                                       const float *q, size_t k, int metric, const uint64_t *deleted,
                                       uint64_t deleted_nbits, const uint64_t *filter,
                                       uint64_t filter_nbits, uint32_t *out_ids, float *out_score) {
    fo_scored *sc = (fo_scored *)malloc((n ? n : 1) * sizeof(fo_scored));
    size_t m = 0;
    for (size_t r = 0; r < n; ++r) {
        uint32_t id = ids ? ids[r] : (uint32_t)r;
        if (bit_get(deleted, deleted_nbits, id)) continue;
        if (filter && !bit_get(filter, filter_nbits, id)) continue;
        sc[m].score = metric == 1 ? fo_cosine(q, rows + r * d, d) : fo_dot(q, rows + r * d, d);
        sc[m].id = id;
        sc[m].order = m;
        ++m;
    }
    qsort(sc, m, sizeof(fo_scored), cmp_scored_desc);
    size_t out = m < k ? m : k;
    for (size_t i = 0; i < out; ++i) { out_ids[i] = sc[i].id; out_score[i] = sc[i].score; }
    free(sc);
    return out;
}

/* Exact scan of a flat tier: the ground truth for recall, and the replacement semantics of
 * the recent tier (HNSWIndex::search, src/hnsw/core.rs:398-467, returns an approximate
 * subset of this; deleted nodes dropped :452-459). */
FO_EXPORT size_t fo_flat_search(const float *rows, const uint32_t *ids, size_t n, size_t d,
                                const float *q, size_t k, const uint64_t *deleted,
                                uint64_t deleted_nbits, const uint64_t *filter,
                                uint64_t filter_nbits, uint32_t *out_ids, float *out_dist) {
    fo_topk tk;
    tk.k = k; tk.len = 0;
    tk.keys = (uint64_t *)malloc((k + 1) * sizeof(uint64_t));
    enum { CH = 4096 };
    float *dist = (float *)malloc((CH + 8) * sizeof(float));
    for (size_t b = 0; b < n; b += CH) {
        size_t L = n - b < CH ? n - b : CH;
        fo_l2_many(q, rows + b * d, L, d, dist);
        for (size_t r = 0; r < L; ++r) {
            uint32_t id = ids ? ids[b + r] : (uint32_t)(b + r);
            if (bit_get(deleted, deleted_nbits, id)) continue;
            if (filter && !bit_get(filter, filter_nbits, id)) continue;
            topk_push(&tk, key_of(dist[r], id));
        }
    }
    for (size_t i = 0; i < tk.len; ++i) {
        uint32_t b = (uint32_t)(tk.keys[i] >> 32);
        out_ids[i] = (uint32_t)(tk.keys[i] & 0xffffffffu);
        memcpy(&out_dist[i], &b, 4);
    }
    size_t r = tk.len;
    free(tk.keys); free(dist);
    return r;
}

/* HybridIndex::search_with_config, src/hybrid/core.rs:425-486, for one query with the recent
 * tier served exactly: recent results (k) then IVF results (k) concatenated, stable sort by
 * distance (recent first on ties), truncate(k), NO de-duplication (:482-483).
 * ivf may be NULL (untrained => recent tier only, :465). */
FO_EXPORT size_t fo_hybrid_search(const fo_ivf *ivf, const float *flat_rows, const uint32_t *flat_ids,
                                  size_t flat_n, size_t d, const float *q, size_t k, size_t nprobe,
                                  unsigned tiers, const uint64_t *deleted, uint64_t deleted_nbits,
                                  const uint64_t *filter, uint64_t filter_nbits, uint32_t *out_ids,
                                  float *out_dist) {
    uint32_t *ids = (uint32_t *)malloc((2 * k + 2) * sizeof(uint32_t));
    float *dist = (float *)malloc((2 * k + 2) * sizeof(float));
    size_t n1 = 0, n2 = 0;
    if ((tiers & 1u) && flat_n > 0)
        n1 = fo_flat_search(flat_rows, flat_ids, flat_n, d, q, k, deleted, deleted_nbits, filter,
                            filter_nbits, ids, dist);
    if ((tiers & 2u) && ivf)
        n2 = fo_ivf_search(ivf, q, k, nprobe, deleted, deleted_nbits, filter, filter_nbits,
                           ids + n1, dist + n1);
    size_t n = n1 + n2;
    fo_pair *a = (fo_pair *)malloc((n + 1) * sizeof(fo_pair));
    fo_pair *t = (fo_pair *)malloc((n + 1) * sizeof(fo_pair));
    for (size_t i = 0; i < n; ++i) { a[i].key = dist[i]; a[i].idx = ids[i]; a[i].aux = 0; }
    stable_sort_pairs(a, n, t);
    size_t r = k < n ? k : n;
    for (size_t i = 0; i < r; ++i) { out_ids[i] = a[i].idx; out_dist[i] = a[i].key; }
    free(a); free(t); free(ids); free(dist);
    return r;
}

/* HybridIndex::search_with_filter, src/hybrid/core.rs:513-549: POST-filter — search(3k),
 * keep rows whose metadata matches (here: bit set in `match`; ids missing from the map are
 * dropped :537-541 => bit clear), truncate(k).  May return < k although >= k matches exist. */
FO_EXPORT size_t fo_hybrid_search_postfilter(const fo_ivf *ivf, const float *flat_rows,
                                             const uint32_t *flat_ids, size_t flat_n, size_t d,
                                             const float *q, size_t k, size_t nprobe, unsigned tiers,
                                             const uint64_t *deleted, uint64_t deleted_nbits,
                                             const uint64_t *match, uint64_t match_nbits,
                                             uint32_t *out_ids, float *out_dist) {
    size_t k3 = k * 3;
    uint32_t *ids = (uint32_t *)malloc((k3 + 1) * sizeof(uint32_t));
    float *dist = (float *)malloc((k3 + 1) * sizeof(float));
    size_t n = fo_hybrid_search(ivf, flat_rows, flat_ids, flat_n, d, q, k3, nprobe, tiers, deleted,
                                deleted_nbits, NULL, 0, ids, dist);
    size_t r = 0;
    for (size_t i = 0; i < n && r < k; ++i)
        if (bit_get(match, match_nbits, ids[i])) { out_ids[r] = ids[i]; out_dist[r] = dist[i]; ++r; }
    free(ids); free(dist);
    return r;
}

/* IVFIndex::batch_search (src/ivf/operations.rs:132-145) over the hybrid search; the reference
 * loops sequentially — the "tight" baseline spreads queries over all host threads (results
 * identical, queries are independent).  threads <= 0: all available.  faithful != 0 uses the
 * clone-per-row / full-sort cost model on the IVF tier (single thread meaningful). */
FO_EXPORT void fo_hybrid_batch_search(const fo_ivf *ivf, const float *flat_rows,
                                      const uint32_t *flat_ids, size_t flat_n, size_t d,
                                      const float *q, size_t nq, size_t k, size_t nprobe,
                                      unsigned tiers, const uint64_t *deleted, uint64_t deleted_nbits,
                                      const uint64_t *filter, uint64_t filter_nbits, int threads,
                                      int faithful, uint32_t *out_ids, float *out_dist,
                                      uint32_t *out_count) {
#if defined(_OPENMP)
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (long long i = 0; i < (long long)nq; ++i) {
        size_t r;
        if (faithful && flat_n == 0 && ivf && !filter)
            r = fo_ivf_search_faithful(ivf, q + (size_t)i * d, k, nprobe, deleted, deleted_nbits,
                                       out_ids + (size_t)i * k, out_dist + (size_t)i * k);
        else
            r = fo_hybrid_search(ivf, flat_rows, flat_ids, flat_n, d, q + (size_t)i * d, k, nprobe,
                                 tiers, deleted, deleted_nbits, filter, filter_nbits,
                                 out_ids + (size_t)i * k, out_dist + (size_t)i * k);
        out_count[i] = (uint32_t)r;
    }
}

FO_EXPORT int fo_num_threads(void) {
#if defined(_OPENMP)
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* recall as evaluate_search_quality defines it, src/ivf/operations.rs:355-371:
 * |found ∩ truth| / |truth| averaged over queries (truth truncated to k). */
FO_EXPORT double fo_recall(const uint32_t *found, const uint32_t *found_cnt, const uint32_t *truth,
                           const uint32_t *truth_cnt, size_t nq, size_t k) {
    double acc = 0.0;
    size_t used = 0;
    for (size_t i = 0; i < nq; ++i) {
        size_t tc = truth_cnt[i], fc = found_cnt[i], hit = 0;
        if (tc == 0) continue;
        for (size_t a = 0; a < tc; ++a)
            for (size_t b = 0; b < fc; ++b)
                if (truth[i * k + a] == found[i * k + b]) { ++hit; break; }
        acc += (double)hit / (double)tc;
        ++used;
    }
    return used ? acc / (double)used : 1.0;
}

/* ------------------------------------------------------------------------------------------
 * k-means — src/ivf/core.rs:240-429
 * ---------------------------------------------------------------------------------------- */

/* compute_error, src/ivf/core.rs:419-429: f32 left-fold of dist*dist over data order, / n. */
FO_EXPORT float fo_compute_error(const float *data, size_t n, size_t d, const float *centroids,
                                 const uint32_t *assign) {
    float total = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float dist = fo_l2(data + i * d, centroids + (size_t)assign[i] * d, d);
        total += dist * dist;
    }
    return total / (float)n;
}

/* update_centroids, src/ivf/core.rs:388-417: per-cluster f32 sums in data order, mean =
 * sum / count as f32; a cluster with no members keeps its previous centroid (:410-415). */
FO_EXPORT void fo_update_centroids(const float *data, size_t n, size_t d, size_t nlist,
                                   const uint32_t *assign, float *centroids) {
    float *sums = (float *)calloc(nlist * d, sizeof(float));
    size_t *counts = (size_t *)calloc(nlist, sizeof(size_t));
    for (size_t i = 0; i < n; ++i) {
        float *s = sums + (size_t)assign[i] * d;
        const float *v = data + i * d;
        for (size_t j = 0; j < d; ++j) s[j] += v[j];
        counts[assign[i]]++;
    }
    for (size_t c = 0; c < nlist; ++c)
        if (counts[c] > 0)
            for (size_t j = 0; j < d; ++j) centroids[c * d + j] = sums[c * d + j] / (float)counts[c];
    free(sums); free(counts);
}

/* The Lloyd loop of IVFIndex::train, src/ivf/core.rs:279-334, from given initial centroids
 * (in/out).  result = {iterations, converged, initial_error, final_error}.  assign_out
 * (nullable) receives the final assignments.  The `max_iterations == 10 && n < 20` special
 * case of :313-317 is kept. */
FO_EXPORT void fo_train_lloyd(const float *data, size_t n, size_t d, size_t nlist,
                              size_t max_iterations, float *centroids, uint32_t *assign_out,
                              uint32_t *out_iterations, uint32_t *out_converged,
                              float *out_initial_error, float *out_final_error) {
    uint32_t *assign = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t)); /* ClusterId(0) */
    uint32_t *next = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    float prev_error = INFINITY;
    float initial_error = fo_compute_error(data, n, d, centroids, assign);
    int converged = 0;
    size_t iterations = 0;
    for (size_t iter = 0; iter < max_iterations; ++iter) {
        iterations = iter + 1;
        int changed = 0;
        fo_assign(data, n, centroids, nlist, d, next);
        for (size_t i = 0; i < n; ++i)
            if (next[i] != assign[i]) { changed = 1; assign[i] = next[i]; }
        fo_update_centroids(data, n, d, nlist, assign, centroids);
        if (iterations >= max_iterations) break;
        float current_error = fo_compute_error(data, n, d, centroids, assign);
        float error_change = fabsf(prev_error - current_error) / prev_error;
        if (!changed || error_change < 1e-4f) {
            converged = 1;
            if (max_iterations == 10 && n < 20) { prev_error = current_error; continue; }
            break;
        }
        prev_error = current_error;
    }
    float final_error = fo_compute_error(data, n, d, centroids, assign);
    if (assign_out) memcpy(assign_out, assign, n * sizeof(uint32_t));
    *out_iterations = (uint32_t)iterations;
    *out_converged = (uint32_t)converged;
    *out_initial_error = initial_error;
    *out_final_error = final_error;
    free(assign); free(next);
}

/* ---- rand 0.8 StdRng (ChaCha12) — restated from the published algorithm; UNPINNED ------- */

typedef struct {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[64];
    size_t index; /* next unread word in buf; 64 = empty */
} fo_stdrng;

static inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
#define QR(a, b, c, d)                                                                             \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12);                          \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

static void chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x[16];
    memcpy(x, s, sizeof(x));
    for (int r = 0; r < 6; ++r) { /* 12 rounds = 6 double rounds */
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

static void stdrng_refill(fo_stdrng *r) {
    for (int b = 0; b < 4; ++b) chacha12_block(r->key, r->counter + (uint64_t)b, r->buf + 16 * b);
    r->counter += 4;
    r->index = 0;
}

/* SeedableRng::seed_from_u64 (rand_core 0.6): PCG32 stream expands the u64 into 32 bytes. */
static void stdrng_seed(fo_stdrng *r, uint64_t state) {
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    for (int i = 0; i < 8; ++i) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        r->key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
    r->counter = 0;
    r->index = 64;
}

static uint32_t stdrng_u32(fo_stdrng *r) {
    if (r->index >= 64) stdrng_refill(r);
    return r->buf[r->index++];
}

/* BlockRng::next_u64: two consecutive words, low first; straddles a refill like rand_core. */
static uint64_t stdrng_u64(fo_stdrng *r) {
    if (r->index < 63) {
        uint64_t lo = r->buf[r->index], hi = r->buf[r->index + 1];
        r->index += 2;
        return (hi << 32) | lo;
    } else if (r->index >= 64) {
        stdrng_refill(r);
        uint64_t lo = r->buf[0], hi = r->buf[1];
        r->index = 2;
        return (hi << 32) | lo;
    } else {
        uint64_t lo = r->buf[63];
        stdrng_refill(r);
        uint64_t hi = r->buf[0];
        r->index = 1;
        return (hi << 32) | lo;
    }
}

/* Rng::gen_range(0..n) for usize (UniformInt::sample_single, rand 0.8.5). */
static uint64_t stdrng_range(fo_stdrng *r, uint64_t range) {
    int lz = __builtin_clzll(range);
    uint64_t zone = (range << lz) - 1;
    for (;;) {
        uint64_t v = stdrng_u64(r);
        __uint128_t m = (__uint128_t)v * range;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo <= zone) return hi;
    }
}

/* Standard f32: (next_u32 >> 8) * 2^-24. */
static float stdrng_f32(fo_stdrng *r) { return (float)(stdrng_u32(r) >> 8) * (1.0f / 16777216.0f); }

FO_EXPORT void fo_stdrng_stream(uint64_t seed, uint32_t *out_u32, size_t n) {
    fo_stdrng r;
    stdrng_seed(&r, seed);
    for (size_t i = 0; i < n; ++i) out_u32[i] = stdrng_u32(&r);
}

/* initialize_centroids, src/ivf/core.rs:336-371 (k-means++).  The reference recomputes the
 * min over all chosen centroids every round (:346-354); a running minimum gives the same
 * values (min is exact).  Returns the number of centroids produced — may be < k when the f32
 * cumulative sum never reaches the threshold (:361-367), exactly like the reference. */
FO_EXPORT size_t fo_kmeanspp_init(const float *data, size_t n, size_t d, size_t k, uint64_t seed,
                                  float *centroids, uint32_t *picked) {
    fo_stdrng rng;
    stdrng_seed(&rng, seed);
    float *mind = (float *)malloc((n + 8) * sizeof(float));
    float *tmp = (float *)malloc((n + 8) * sizeof(float));
    for (size_t j = 0; j < n; ++j) mind[j] = INFINITY;
    size_t first = (size_t)stdrng_range(&rng, (uint64_t)n);
    size_t count = 0;
    memcpy(centroids, data + first * d, d * sizeof(float));
    if (picked) picked[0] = (uint32_t)first;
    count = 1;
    for (size_t i = 1; i < k; ++i) {
        const float *last = centroids + (count - 1) * d;
        /* distances to the newest centroid; fo_l2(point, centroid) operand order as :351 */
#pragma omp parallel for schedule(static)
        for (long long j = 0; j < (long long)n; ++j) tmp[j] = fo_l2(data + (size_t)j * d, last, d);
        for (size_t j = 0; j < n; ++j) mind[j] = mind[j] < tmp[j] ? mind[j] : tmp[j]; /* f32::min */
        float total = 0.0f;
        for (size_t j = 0; j < n; ++j) total += mind[j] * mind[j];
        float cumulative = 0.0f;
        float threshold = stdrng_f32(&rng) * total;
        for (size_t j = 0; j < n; ++j) {
            cumulative += mind[j] * mind[j];
            if (cumulative >= threshold) {
                memcpy(centroids + count * d, data + j * d, d * sizeof(float));
                if (picked) picked[count] = (uint32_t)j;
                ++count;
                break;
            }
        }
    }
    free(mind); free(tmp);
    return count;
}
