/*
 * oracle.h — CPU restatement of takurot/Pyrope's vector-scan hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pyrope_b200/ may include, link,
 * load or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * reported CPU baseline — never as the product path.
 *
 * Parity status: the reference is C# (.NET) and cannot be built or run in this
 * image (no dotnet/mono).  The restatement is pinned against (i) every numeric
 * and functional case in the reference's own unit tests for this path
 * (tests/Pyrope.GarnetServer.Tests/Vector/ *.cs files, re-expressed in
 * tests/test_oracle_reference_cases.py) and (ii) published known-answer values
 * of System.Random.  Bit-level agreement with a live .NET run is UNPINNED
 * (the reference pins none itself: see SURVEY.md §8c).
 *
 * Conventions chosen where the reference leaves them to the runtime
 * (System.Numerics.Vector<float>): W = 8 lanes (x64 AVX2 default), separate
 * mul and add (no FMA contraction; build with -ffp-contract=off), horizontal
 * sum of a lane vector = ((v0+v1)+(v2+v3)) + ((v4+v5)+(v6+v7)) (vdpps per
 * 128-bit half, then low+high).
 *
 * All ids are int64 ordinals (the reference's string ids stay in the host shim).
 * Scores are "higher is better": L2 -> -||q-x||^2, IP -> q.x, Cosine -> cos.
 */
#ifndef PYROPE_ORACLE_H
#define PYROPE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_L2 = 0, ORC_IP = 1, ORC_COSINE = 2 }; /* VectorMetric, IVectorIndex.cs:5-10 */

/* ---- System.Random(seed) (.NET BCL, Net5CompatSeedImpl; not under /root/reference) ---- */
typedef struct orc_random orc_random;
orc_random *orc_random_new(int32_t seed);
void orc_random_free(orc_random *r);
int32_t orc_random_next(orc_random *r);        /* Random.Next()        */
double orc_random_next_double(orc_random *r);  /* Random.NextDouble()  */
/* Pyrope.Benchmarks/Program.cs:251-263 GenerateRandomVectors: out[i] = (float)rng.NextDouble() */
void orc_random_fill(int32_t seed, int64_t n, float *out);

/* ---- VectorMath.cs ---- */
float orc_dot(const float *a, const float *b, int n);          /* DotProduct        :8-37   */
float orc_l2sq(const float *a, const float *b, int n);         /* L2Squared         :39-70  */
float orc_norm(const float *v, int n);                         /* ComputeNorm       :72-100 */
float orc_cosine(const float *q, const float *v, float qn, float vn, int n); /* :102-109 */
float orc_dot_unsafe(const float *a, const float *b, int n);   /* DotProductUnsafe  :128-186 */
float orc_l2sq_unsafe(const float *a, const float *b, int n);  /* L2SquaredUnsafe   :188-253 */

/* ---- KMeansUtils.cs ---- */
/* FindNearestCentroid :70-93.  cnorms may be NULL unless metric == COSINE. */
int orc_find_nearest_centroid(const float *vec, const float *centroids, const float *cnorms,
                              int k, int dim, int metric);
/* Train :10-68.  data is n x dim row-major with leading dimension ld (sub-vector views).
 * Returns the number of centroids written (min(k, n), k<=0 -> 1). iters_out may be NULL. */
int orc_kmeans_train(const float *data, int64_t n, int dim, int64_t ld, int k, int metric,
                     int max_iter, int32_t seed, float *centroids_out, int *iters_out);

/* ---- ProductQuantizer.cs ---- */
typedef struct orc_pq orc_pq;
orc_pq *orc_pq_new(int dim, int m, int k);                      /* ctor :15-26 (NULL on bad args) */
void orc_pq_free(orc_pq *pq);
void orc_pq_train(orc_pq *pq, const float *data, int64_t n);    /* Train :28-58 */
int orc_pq_ksub(const orc_pq *pq, int m);                       /* trained codewords in subspace m */
void orc_pq_get_codebook(const orc_pq *pq, float *out);         /* [m][K][subDim], zero padded */
void orc_pq_set_codebook(orc_pq *pq, const float *cb, const int *ksub);
int orc_pq_encode(const orc_pq *pq, const float *vec, uint8_t *code);      /* Encode :60-80 */
void orc_pq_distance_table(const orc_pq *pq, const float *query, float *table); /* [m][K] :98-120 */

/* ---- BruteForceVectorIndex.cs (FLAT) ---- */
typedef struct orc_flat orc_flat;
orc_flat *orc_flat_new(int dim, int metric);
void orc_flat_free(orc_flat *ix);
/* return 0 ok, -1 duplicate id (InvalidOperationException :143) */
int orc_flat_add(orc_flat *ix, int64_t id, const float *vec);
void orc_flat_upsert(orc_flat *ix, int64_t id, const float *vec);
int orc_flat_delete(orc_flat *ix, int64_t id);
int orc_flat_count(const orc_flat *ix);                          /* GetStats :119-131 */
void orc_flat_add_batch(orc_flat *ix, int64_t n, const int64_t *ids, const float *X);
/* Search :275-379.  max_scans < 0 means "no option".  Returns result count, or -2 if topk<=0. */
int orc_flat_search(const orc_flat *ix, const float *q, int topk, int64_t max_scans,
                    int64_t *ids_out, float *scores_out);

/* ---- IvfFlatVectorIndex.cs ---- */
typedef struct orc_ivfflat orc_ivfflat;
orc_ivfflat *orc_ivfflat_new(int dim, int metric, int nlist);
void orc_ivfflat_free(orc_ivfflat *ix);
void orc_ivfflat_set_nprobe(orc_ivfflat *ix, int nprobe);        /* CombineNProbe :14 */
void orc_ivfflat_add(orc_ivfflat *ix, int64_t id, const float *vec);   /* Add/Upsert :39-60 */
void orc_ivfflat_add_batch(orc_ivfflat *ix, int64_t n, const int64_t *ids, const float *X);
int orc_ivfflat_delete(orc_ivfflat *ix, int64_t id);             /* :62-83 */
void orc_ivfflat_build(orc_ivfflat *ix);                         /* :85-145 */
int orc_ivfflat_is_built(const orc_ivfflat *ix);
int orc_ivfflat_ncentroids(const orc_ivfflat *ix);
/* test infrastructure: adopt centroids and inverted lists (offsets [nlist+1], ids / vecs list-major) built elsewhere */
void orc_ivfflat_adopt(orc_ivfflat *ix, int nlist, const float *centroids, const int64_t *offs, const int64_t *ids,
                       const float *vecs);
void orc_ivfflat_get_centroids(const orc_ivfflat *ix, float *out);
int orc_ivfflat_count(const orc_ivfflat *ix);                    /* GetStats :300-312 */
int orc_ivfflat_list_size(const orc_ivfflat *ix, int list);
void orc_ivfflat_get_list(const orc_ivfflat *ix, int list, int64_t *ids_out);
/* Search :147-231.  max_scans<0 / nprobe<0 mean "option absent". */
int orc_ivfflat_search(const orc_ivfflat *ix, const float *q, int topk, int64_t max_scans,
                       int nprobe, int64_t *ids_out, float *scores_out);

/* ---- IvfPqVectorIndex.cs ---- */
typedef struct orc_ivfpq orc_ivfpq;
orc_ivfpq *orc_ivfpq_new(int dim, int metric, int m, int k, int nlist);
void orc_ivfpq_free(orc_ivfpq *ix);
void orc_ivfpq_add(orc_ivfpq *ix, int64_t id, const float *vec);      /* :36-47 */
void orc_ivfpq_add_batch(orc_ivfpq *ix, int64_t n, const int64_t *ids, const float *X);
int orc_ivfpq_delete(orc_ivfpq *ix, int64_t id);                      /* :48-53 (buffer only) */
void orc_ivfpq_build(orc_ivfpq *ix);                                  /* :55-116 */
int orc_ivfpq_is_built(const orc_ivfpq *ix);
int orc_ivfpq_ncentroids(const orc_ivfpq *ix);
void orc_ivfpq_get_centroids(const orc_ivfpq *ix, float *out);
const orc_pq *orc_ivfpq_pq(const orc_ivfpq *ix);
int orc_ivfpq_list_size(const orc_ivfpq *ix, int list);
void orc_ivfpq_get_list(const orc_ivfpq *ix, int list, int64_t *ids_out, uint8_t *codes_out);
/* Adopt externally produced state (frozen codebooks + list layout), for large-config baselines:
 * centroids [nlist][dim], codebook [m][K][subDim], list_offsets [nlist+1], ids/codes list-major. */
void orc_ivfpq_adopt(orc_ivfpq *ix, int nlist, const float *centroids, const float *codebook,
                     const int64_t *list_offsets, const int64_t *ids, const uint8_t *codes);
/* Search :118-212.  nprobe<0 -> default 1.  MaxScans is ignored by the reference. */
int orc_ivfpq_search(const orc_ivfpq *ix, const float *q, int topk, int nprobe,
                     int64_t *ids_out, float *scores_out);

/* ---- DeltaVectorIndex.cs:95-121: merge head/tail result lists (head wins on id) ---- */
int orc_delta_merge(const int64_t *head_ids, const float *head_scores, int nh,
                    const int64_t *tail_ids, const float *tail_scores, int nt, int topk,
                    int64_t *ids_out, float *scores_out);

/* ---- batched drivers: one query per thread (the reference's concurrency model) ---- */
/* kind: 0 FLAT, 1 IVF_FLAT, 2 IVF_PQ.  outputs are [nq][topk], counts [nq]. */
void orc_search_batch(int kind, const void *ix, const float *Q, int64_t nq, int topk,
                      int64_t max_scans, int nprobe, int nthreads, int64_t *ids_out,
                      float *scores_out, int32_t *counts_out);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
