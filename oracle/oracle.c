/*
 * oracle.c — CPU restatement of takurot/Pyrope's vector-scan hot path (see oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY — never linked into or called from the product library.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/Pyrope.GarnetServer/Vector unless stated otherwise).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define W 8 /* System.Numerics.Vector<float>.Count on x64/AVX2 */

/* ============================================================================================
 * System.Random(int seed) — .NET BCL Net5CompatSeedImpl / CompatPrng (Knuth subtractive).
 * Not in /root/reference (BCL dependency, runtime net8.0+); restated from the published
 * algorithm.  Known answers checked in tests: Random(0).Next()=1559595546,
 * Random(1).Next()=534011718, Random(42).Next()=1434747710,
 * Random(42).NextDouble()=0.6681064659115423.
 * Call sites: KMeansUtils.cs:16,20; Pyrope.Benchmarks/Program.cs:253; reference tests.
 * ============================================================================================ */
struct orc_random {
    int32_t sa[56];
    int inext, inextp;
};

#define DN_INT_MAX 2147483647

static void dn_random_init(orc_random *r, int32_t seed) {
    int32_t subtraction = (seed == INT32_MIN) ? DN_INT_MAX : (seed < 0 ? -seed : seed);
    int32_t mj = 161803398 - subtraction;
    int32_t mk = 1;
    int ii = 0;
    memset(r->sa, 0, sizeof r->sa);
    r->sa[55] = mj;
    for (int i = 1; i < 55; i++) {
        if ((ii += 21) >= 55) ii -= 55;
        r->sa[ii] = mk;
        mk = mj - mk;
        if (mk < 0) mk += DN_INT_MAX;
        mj = r->sa[ii];
    }
    for (int k = 1; k < 5; k++) {
        for (int i = 1; i < 56; i++) {
            int n = i + 30;
            if (n >= 55) n -= 55;
            /* C# int arithmetic wraps; do it in uint32 to stay defined in C */
            r->sa[i] = (int32_t)((uint32_t)r->sa[i] - (uint32_t)r->sa[1 + n]);
            if (r->sa[i] < 0) r->sa[i] += DN_INT_MAX;
        }
    }
    r->inext = 0;
    r->inextp = 21;
}

static inline int32_t dn_random_sample(orc_random *r) {
    int a = r->inext, b = r->inextp;
    if (++a >= 56) a = 1;
    if (++b >= 56) b = 1;
    int32_t v = (int32_t)((uint32_t)r->sa[a] - (uint32_t)r->sa[b]);
    if (v == DN_INT_MAX) v--;
    if (v < 0) v += DN_INT_MAX;
    r->sa[a] = v;
    r->inext = a;
    r->inextp = b;
    return v;
}

orc_random *orc_random_new(int32_t seed) {
    orc_random *r = (orc_random *)malloc(sizeof *r);
    dn_random_init(r, seed);
    return r;
}
void orc_random_free(orc_random *r) { free(r); }
int32_t orc_random_next(orc_random *r) { return dn_random_sample(r); }
double orc_random_next_double(orc_random *r) { return dn_random_sample(r) * (1.0 / DN_INT_MAX); }

void orc_random_fill(int32_t seed, int64_t n, float *out) {
    orc_random r;
    dn_random_init(&r, seed);
    for (int64_t i = 0; i < n; i++) out[i] = (float)(dn_random_sample(&r) * (1.0 / DN_INT_MAX));
}

/* ============================================================================================
 * VectorMath.cs
 * ============================================================================================ */

/* Vector.Dot(v, Vector<float>.One): vdpps per 128-bit half then low+high. */
static inline float hsum8(const float *v) {
    float lo = (v[0] + v[1]) + (v[2] + v[3]);
    float hi = (v[4] + v[5]) + (v[6] + v[7]);
    return lo + hi;
}

/* VectorMath.cs:8-37 DotProduct — single vector accumulator, step W, scalar tail. */
float orc_dot(const float *a, const float *b, int n) {
    int i = 0;
    float sum = 0.0f;
    if (n >= W) {
        float acc[W] = {0};
        int end = n - W;
        for (; i <= end; i += W)
            for (int j = 0; j < W; j++) {
                float p = a[i + j] * b[i + j];
                acc[j] = acc[j] + p;
            }
        sum += hsum8(acc);
    }
    for (; i < n; i++) {
        float p = a[i] * b[i];
        sum += p;
    }
    return sum;
}

/* VectorMath.cs:39-70 L2Squared */
float orc_l2sq(const float *a, const float *b, int n) {
    int i = 0;
    float sum = 0.0f;
    if (n >= W) {
        float acc[W] = {0};
        int end = n - W;
        for (; i <= end; i += W)
            for (int j = 0; j < W; j++) {
                float d = a[i + j] - b[i + j];
                float p = d * d;
                acc[j] = acc[j] + p;
            }
        sum += hsum8(acc);
    }
    for (; i < n; i++) {
        float d = a[i] - b[i];
        float p = d * d;
        sum += p;
    }
    return sum;
}

/* VectorMath.cs:72-100 ComputeNorm */
float orc_norm(const float *v, int n) {
    int i = 0;
    float sum = 0.0f;
    if (n >= W) {
        float acc[W] = {0};
        int end = n - W;
        for (; i <= end; i += W)
            for (int j = 0; j < W; j++) {
                float p = v[i + j] * v[i + j];
                acc[j] = acc[j] + p;
            }
        sum += hsum8(acc);
    }
    for (; i < n; i++) {
        float p = v[i] * v[i];
        sum += p;
    }
    return sqrtf(sum); /* MathF.Sqrt: correctly rounded */
}

/* VectorMath.cs:102-109 Cosine(query, vector, queryNorm, vectorNorm) */
float orc_cosine(const float *q, const float *v, float qn, float vn, int n) {
    if (qn < 1e-6f || vn < 1e-6f) return 0.0f;
    float dot = orc_dot(q, v, n);
    float den = qn * vn;
    return dot / den;
}

/* VectorMath.cs:128-186 DotProductUnsafe — 4 accumulators, remainder accumulator, scalar tail */
float orc_dot_unsafe(const float *a, const float *b, int n) {
    int i = 0;
    float sum = 0.0f;
    if (n >= W * 4) {
        float a1[W] = {0}, a2[W] = {0}, a3[W] = {0}, a4[W] = {0}, fin[W];
        int end = n - W * 4;
        while (i <= end) {
            for (int j = 0; j < W; j++) {
                float p1 = a[i + j] * b[i + j];
                float p2 = a[i + W + j] * b[i + W + j];
                float p3 = a[i + 2 * W + j] * b[i + 2 * W + j];
                float p4 = a[i + 3 * W + j] * b[i + 3 * W + j];
                a1[j] = a1[j] + p1;
                a2[j] = a2[j] + p2;
                a3[j] = a3[j] + p3;
                a4[j] = a4[j] + p4;
            }
            i += W * 4;
        }
        for (int j = 0; j < W; j++) fin[j] = ((a1[j] + a2[j]) + a3[j]) + a4[j];
        sum += hsum8(fin);
    }
    if (i <= n - W) {
        float acc[W] = {0};
        while (i <= n - W) {
            for (int j = 0; j < W; j++) {
                float p = a[i + j] * b[i + j];
                acc[j] = acc[j] + p;
            }
            i += W;
        }
        sum += hsum8(acc);
    }
    for (; i < n; i++) {
        float p = a[i] * b[i];
        sum += p;
    }
    return sum;
}

/* VectorMath.cs:188-253 L2SquaredUnsafe */
float orc_l2sq_unsafe(const float *a, const float *b, int n) {
    int i = 0;
    float sum = 0.0f;
    if (n >= W * 4) {
        float a1[W] = {0}, a2[W] = {0}, a3[W] = {0}, a4[W] = {0}, fin[W];
        int end = n - W * 4;
        while (i <= end) {
            for (int j = 0; j < W; j++) {
                float d1 = a[i + j] - b[i + j];
                float d2 = a[i + W + j] - b[i + W + j];
                float d3 = a[i + 2 * W + j] - b[i + 2 * W + j];
                float d4 = a[i + 3 * W + j] - b[i + 3 * W + j];
                float p1 = d1 * d1, p2 = d2 * d2, p3 = d3 * d3, p4 = d4 * d4;
                a1[j] = a1[j] + p1;
                a2[j] = a2[j] + p2;
                a3[j] = a3[j] + p3;
                a4[j] = a4[j] + p4;
            }
            i += W * 4;
        }
        for (int j = 0; j < W; j++) fin[j] = ((a1[j] + a2[j]) + a3[j]) + a4[j];
        sum += hsum8(fin);
    }
    if (i <= n - W) {
        float acc[W] = {0};
        while (i <= n - W) {
            for (int j = 0; j < W; j++) {
                float d = a[i + j] - b[i + j];
                float p = d * d;
                acc[j] = acc[j] + p;
            }
            i += W;
        }
        sum += hsum8(acc);
    }
    for (; i < n; i++) {
        float d = a[i] - b[i];
        float p = d * d;
        sum += p;
    }
    return sum;
}

/* ============================================================================================
 * .NET BCL containers the reference leans on (restated; not under /root/reference):
 *   PriorityQueue<TElement,float>  — quaternary min-heap
 *   List<T>.Sort(Comparison)       — ArraySortHelper introsort (unstable)
 *   Dictionary<K,V>                — insertion order with LIFO reuse of removed slots
 * ============================================================================================ */
typedef struct {
    int64_t id;
    float score;
} res_t;

typedef struct {
    res_t *nodes;
    int size, cap;
} pq4_t;

static void pq4_init(pq4_t *h, int cap) {
    h->cap = cap < 4 ? 4 : cap;
    h->size = 0;
    h->nodes = (res_t *)malloc(sizeof(res_t) * (size_t)h->cap);
}
static void pq4_free(pq4_t *h) { free(h->nodes); }

static void pq4_enqueue(pq4_t *h, res_t node) {
    if (h->size == h->cap) {
        h->cap *= 2;
        h->nodes = (res_t *)realloc(h->nodes, sizeof(res_t) * (size_t)h->cap);
    }
    int idx = h->size++;
    while (idx > 0) { /* MoveUpDefaultComparer */
        int parent = (idx - 1) >> 2;
        if (node.score < h->nodes[parent].score) {
            h->nodes[idx] = h->nodes[parent];
            idx = parent;
        } else
            break;
    }
    h->nodes[idx] = node;
}

static res_t pq4_dequeue(pq4_t *h) {
    res_t root = h->nodes[0];
    int size = --h->size;
    if (size > 0) { /* RemoveRootNode -> MoveDownDefaultComparer(lastNode, 0) */
        res_t node = h->nodes[size];
        int idx = 0, i;
        while ((i = 4 * idx + 1) < size) {
            res_t minc = h->nodes[i];
            int mini = i;
            int ub = i + 4 < size ? i + 4 : size;
            while (++i < ub) {
                if (h->nodes[i].score < minc.score) {
                    minc = h->nodes[i];
                    mini = i;
                }
            }
            if (node.score <= minc.score) break;
            h->nodes[idx] = minc;
            idx = mini;
        }
        h->nodes[idx] = node;
    }
    return root;
}

/* comparison (a,b) => b.Score.CompareTo(a.Score): <0 when a sorts first (a.score > b.score) */
static inline int cmp_desc(const res_t *a, const res_t *b) {
    return (b->score < a->score) ? -1 : (b->score > a->score) ? 1 : 0;
}
static inline void rswap(res_t *k, int i, int j) {
    res_t t = k[i];
    k[i] = k[j];
    k[j] = t;
}
static inline void swap_if_greater(res_t *k, int i, int j) {
    if (cmp_desc(&k[i], &k[j]) > 0) rswap(k, i, j);
}
static void insertion_sort(res_t *k, int n) {
    for (int i = 0; i < n - 1; i++) {
        res_t t = k[i + 1];
        int j = i;
        while (j >= 0 && cmp_desc(&t, &k[j]) < 0) {
            k[j + 1] = k[j];
            j--;
        }
        k[j + 1] = t;
    }
}
static void down_heap(res_t *k, int i, int n) {
    res_t d = k[i - 1];
    while (i <= n / 2) {
        int child = 2 * i;
        if (child < n && cmp_desc(&k[child - 1], &k[child]) < 0) child++;
        if (!(cmp_desc(&d, &k[child - 1]) < 0)) break;
        k[i - 1] = k[child - 1];
        i = child;
    }
    k[i - 1] = d;
}
static void heap_sort(res_t *k, int n) {
    for (int i = n >> 1; i >= 1; i--) down_heap(k, i, n);
    for (int i = n; i > 1; i--) {
        rswap(k, 0, i - 1);
        down_heap(k, 1, i - 1);
    }
}
static int pick_pivot_and_partition(res_t *k, int n) {
    int hi = n - 1, middle = hi >> 1;
    swap_if_greater(k, 0, middle);
    swap_if_greater(k, 0, hi);
    swap_if_greater(k, middle, hi);
    res_t pivot = k[middle];
    rswap(k, middle, hi - 1);
    int left = 0, right = hi - 1;
    while (left < right) {
        while (cmp_desc(&k[++left], &pivot) < 0) {
        }
        while (cmp_desc(&pivot, &k[--right]) < 0) {
        }
        if (left >= right) break;
        rswap(k, left, right);
    }
    if (left != hi - 1) rswap(k, left, hi - 1);
    return left;
}
static void intro_sort(res_t *k, int n, int depth) {
    int part = n;
    while (part > 1) {
        if (part <= 16) {
            if (part == 2) {
                swap_if_greater(k, 0, 1);
                return;
            }
            if (part == 3) {
                swap_if_greater(k, 0, 1);
                swap_if_greater(k, 0, 2);
                swap_if_greater(k, 1, 2);
                return;
            }
            insertion_sort(k, part);
            return;
        }
        if (depth == 0) {
            heap_sort(k, part);
            return;
        }
        depth--;
        int p = pick_pivot_and_partition(k, part);
        intro_sort(k + p + 1, part - (p + 1), depth);
        part = p;
    }
}
static void dn_sort_desc(res_t *k, int n) {
    if (n > 1) {
        int lg = 0;
        unsigned v = (unsigned)n;
        while (v >>= 1) lg++;
        intro_sort(k, n, 2 * (lg + 1));
    }
}

/* drain a heap the way every Search does: Dequeue all (ascending), Sort descending */
static int drain_sorted(pq4_t *h, int64_t *ids_out, float *scores_out) {
    int n = h->size;
    res_t *tmp = (res_t *)malloc(sizeof(res_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) tmp[i] = pq4_dequeue(h);
    dn_sort_desc(tmp, n);
    for (int i = 0; i < n; i++) {
        ids_out[i] = tmp[i].id;
        scores_out[i] = tmp[i].score;
    }
    free(tmp);
    return n;
}

/* ---- int64 -> int32 open-addressing map (lookup only; ordering handled by ndict) ---- */
typedef struct {
    int64_t *keys;
    int32_t *vals; /* -1 empty, -2 tombstone */
    int cap, used, filled;
} imap_t;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static void imap_init(imap_t *m, int cap) {
    int c = 16;
    while (c < cap * 2) c <<= 1;
    m->cap = c;
    m->used = m->filled = 0;
    m->keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)c);
    m->vals = (int32_t *)malloc(sizeof(int32_t) * (size_t)c);
    for (int i = 0; i < c; i++) m->vals[i] = -1;
}
static void imap_free(imap_t *m) {
    free(m->keys);
    free(m->vals);
}
static int imap_get(const imap_t *m, int64_t key) {
    int i = (int)(mix64((uint64_t)key) & (uint64_t)(m->cap - 1));
    for (;;) {
        if (m->vals[i] == -1) return -1;
        if (m->vals[i] >= 0 && m->keys[i] == key) return m->vals[i];
        i = (i + 1) & (m->cap - 1);
    }
}
static void imap_put(imap_t *m, int64_t key, int32_t val);
static void imap_grow(imap_t *m) {
    imap_t n;
    imap_init(&n, m->used * 2 + 8);
    for (int i = 0; i < m->cap; i++)
        if (m->vals[i] >= 0) imap_put(&n, m->keys[i], m->vals[i]);
    imap_free(m);
    *m = n;
}
static void imap_put(imap_t *m, int64_t key, int32_t val) {
    if ((m->filled + 1) * 2 > m->cap) imap_grow(m);
    int i = (int)(mix64((uint64_t)key) & (uint64_t)(m->cap - 1));
    int first_tomb = -1;
    for (;;) {
        if (m->vals[i] == -1) break;
        if (m->vals[i] == -2) {
            if (first_tomb < 0) first_tomb = i;
        } else if (m->keys[i] == key) {
            m->vals[i] = val;
            return;
        }
        i = (i + 1) & (m->cap - 1);
    }
    if (first_tomb >= 0)
        i = first_tomb;
    else
        m->filled++;
    m->keys[i] = key;
    m->vals[i] = val;
    m->used++;
}
static int imap_del(imap_t *m, int64_t key) {
    int i = (int)(mix64((uint64_t)key) & (uint64_t)(m->cap - 1));
    for (;;) {
        if (m->vals[i] == -1) return 0;
        if (m->vals[i] >= 0 && m->keys[i] == key) {
            m->vals[i] = -2;
            m->used--;
            return 1;
        }
        i = (i + 1) & (m->cap - 1);
    }
}

/* ---- Dictionary<id, float[]> with .NET enumeration order: entries in slot order, removed
 *      slots pushed on a LIFO free list and reused by the next insert. ---- */
typedef struct {
    int dim;
    int count;     /* high-water slot count (_count) */
    int cap;
    int free_head; /* -1 none */
    int nfree;
    int64_t *ids;
    float *vecs;   /* [cap][dim] */
    float *norms;
    int32_t *next; /* >= -1 live ; <= -2 : free, encoded -3 - next_free (StartOfFreeList) */
    imap_t map;
} ndict_t;

static void ndict_init(ndict_t *d, int dim) {
    memset(d, 0, sizeof *d);
    d->dim = dim;
    d->cap = 16;
    d->free_head = -1;
    d->ids = (int64_t *)malloc(sizeof(int64_t) * 16);
    d->vecs = (float *)malloc(sizeof(float) * 16 * (size_t)dim);
    d->norms = (float *)malloc(sizeof(float) * 16);
    d->next = (int32_t *)malloc(sizeof(int32_t) * 16);
    imap_init(&d->map, 16);
}
static void ndict_free(ndict_t *d) {
    free(d->ids);
    free(d->vecs);
    free(d->norms);
    free(d->next);
    imap_free(&d->map);
}
static void ndict_clear(ndict_t *d) {
    int dim = d->dim;
    ndict_free(d);
    ndict_init(d, dim);
}
static int ndict_live(const ndict_t *d) { return d->count - d->nfree; }
static int ndict_is_live(const ndict_t *d, int slot) { return d->next[slot] >= -1; }
/* d[id] = (vec, norm): overwrite keeps the slot (and so the enumeration position) */
static void ndict_set(ndict_t *d, int64_t id, const float *vec, float norm) {
    int slot = imap_get(&d->map, id);
    if (slot < 0) {
        if (d->nfree > 0) {
            slot = d->free_head;
            d->free_head = -3 - d->next[slot];
            d->nfree--;
        } else {
            if (d->count == d->cap) {
                d->cap *= 2;
                d->ids = (int64_t *)realloc(d->ids, sizeof(int64_t) * (size_t)d->cap);
                d->vecs = (float *)realloc(d->vecs, sizeof(float) * (size_t)d->cap * (size_t)d->dim);
                d->norms = (float *)realloc(d->norms, sizeof(float) * (size_t)d->cap);
                d->next = (int32_t *)realloc(d->next, sizeof(int32_t) * (size_t)d->cap);
            }
            slot = d->count++;
        }
        d->ids[slot] = id;
        d->next[slot] = -1;
        imap_put(&d->map, id, slot);
    }
    memcpy(d->vecs + (size_t)slot * (size_t)d->dim, vec, sizeof(float) * (size_t)d->dim);
    d->norms[slot] = norm;
}
static int ndict_remove(ndict_t *d, int64_t id) {
    int slot = imap_get(&d->map, id);
    if (slot < 0) return 0;
    imap_del(&d->map, id);
    d->next[slot] = -3 - d->free_head;
    d->free_head = slot;
    d->nfree++;
    return 1;
}
static int ndict_contains(const ndict_t *d, int64_t id) { return imap_get(&d->map, id) >= 0; }

/* ============================================================================================
 * KMeansUtils.cs
 * ============================================================================================ */

/* KMeansUtils.cs:70-93 FindNearestCentroid: strict '>' from float.MinValue => lowest index wins. */
static int find_nearest_centroid_ld(const float *vec, const float *centroids, int64_t cld,
                                    const float *cnorms, int k, int dim, int metric) {
    int best = 0;
    float best_score = -3.40282347e+38f; /* float.MinValue */
    float vnorm = metric == ORC_COSINE ? orc_norm(vec, dim) : 0.0f;
    for (int i = 0; i < k; i++) {
        const float *c = centroids + (size_t)i * (size_t)cld;
        float score;
        switch (metric) {
        case ORC_L2: score = -orc_l2sq(vec, c, dim); break;
        case ORC_IP: score = orc_dot(vec, c, dim); break;
        case ORC_COSINE: score = orc_cosine(vec, c, vnorm, cnorms[i], dim); break;
        default: score = 0.0f;
        }
        if (score > best_score) {
            best_score = score;
            best = i;
        }
    }
    return best;
}

int orc_find_nearest_centroid(const float *vec, const float *centroids, const float *cnorms, int k,
                              int dim, int metric) {
    return find_nearest_centroid_ld(vec, centroids, dim, cnorms, k, dim, metric);
}

typedef struct {
    int32_t key;
    int32_t idx;
} keyidx_t;
static int cmp_keyidx(const void *a, const void *b) {
    const keyidx_t *x = (const keyidx_t *)a, *y = (const keyidx_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0); /* LINQ OrderBy is stable */
}

/* KMeansUtils.cs:10-68 Train */
int orc_kmeans_train(const float *data, int64_t n, int dim, int64_t ld, int k, int metric,
                     int max_iter, int32_t seed, float *centroids, int *iters_out) {
    if (iters_out) *iters_out = 0;
    if (n == 0) return 0;
    if (k <= 0) k = 1;
    if (k > n) k = (int)n;

    /* :16-20  data.OrderBy(_ => rnd.Next()).Take(k): one key per element in order, stable sort */
    orc_random rnd;
    dn_random_init(&rnd, seed);
    keyidx_t *keys = (keyidx_t *)malloc(sizeof(keyidx_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) {
        keys[i].key = dn_random_sample(&rnd);
        keys[i].idx = (int32_t)i;
    }
    qsort(keys, (size_t)n, sizeof(keyidx_t), cmp_keyidx);
    for (int c = 0; c < k; c++)
        memcpy(centroids + (size_t)c * dim, data + (size_t)keys[c].idx * (size_t)ld,
               sizeof(float) * (size_t)dim);
    free(keys);

    int32_t *assign = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    float *cnorm = (float *)malloc(sizeof(float) * (size_t)k);
    float *newc = (float *)malloc(sizeof(float) * (size_t)dim);
    int64_t *counts = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    int64_t *offs = (int64_t *)malloc(sizeof(int64_t) * ((size_t)k + 1));
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);

    for (int iter = 0; iter < max_iter; iter++) {
        if (iters_out) *iters_out = iter + 1;
        int changed = 0;
        for (int c = 0; c < k; c++)
            cnorm[c] = metric == ORC_COSINE ? orc_norm(centroids + (size_t)c * dim, dim) : 0.0f;

        /* :33-38 Parallel.For assignment (order-free) */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++)
            assign[i] = find_nearest_centroid_ld(data + (size_t)i * (size_t)ld, centroids, dim,
                                                 cnorm, k, dim, metric);

        /* :40-43 clusters[a].Add(data[i]) in data order */
        memset(counts, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t i = 0; i < n; i++) counts[assign[i]]++;
        offs[0] = 0;
        for (int c = 0; c < k; c++) offs[c + 1] = offs[c] + counts[c];
        memset(counts, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t i = 0; i < n; i++) order[offs[assign[i]] + counts[assign[i]]++] = (int32_t)i;

        /* :46-63 update: fp32 running sum in cluster (data) order, then /= Count */
        for (int c = 0; c < k; c++) {
            int64_t cnt = offs[c + 1] - offs[c];
            if (cnt == 0) continue;
            for (int d = 0; d < dim; d++) newc[d] = 0.0f;
            for (int64_t j = offs[c]; j < offs[c + 1]; j++) {
                const float *v = data + (size_t)order[j] * (size_t)ld;
                for (int d = 0; d < dim; d++) newc[d] += v[d];
            }
            float fc = (float)(int32_t)cnt;
            for (int d = 0; d < dim; d++) newc[d] /= fc;
            /* ArraysEqual :95-101: Math.Abs(a[i]-b[i]) > 1e-6 (float promoted to double) */
            int equal = 1;
            float *oldc = centroids + (size_t)c * dim;
            for (int d = 0; d < dim; d++) {
                float df = oldc[d] - newc[d];
                if ((double)fabsf(df) > 1e-6) {
                    equal = 0;
                    break;
                }
            }
            if (!equal) {
                memcpy(oldc, newc, sizeof(float) * (size_t)dim);
                changed = 1;
            }
        }
        if (!changed) break;
    }
    free(assign);
    free(cnorm);
    free(newc);
    free(counts);
    free(offs);
    free(order);
    return k;
}

/* ============================================================================================
 * ProductQuantizer.cs
 * ============================================================================================ */
struct orc_pq {
    int dim, m, sub, k;
    int trained;
    int *ksub;       /* trained codewords per subspace (<= k) */
    float *codebook; /* [m][k][sub], zero padded beyond ksub */
};

orc_pq *orc_pq_new(int dim, int m, int k) {
    if (m <= 0 || dim % m != 0) return NULL; /* :18 */
    if (k > 256) return NULL;                /* :19 */
    orc_pq *pq = (orc_pq *)calloc(1, sizeof *pq);
    pq->dim = dim;
    pq->m = m;
    pq->sub = dim / m;
    pq->k = k;
    pq->ksub = (int *)calloc((size_t)m, sizeof(int));
    pq->codebook = (float *)calloc((size_t)m * (size_t)k * (size_t)pq->sub, sizeof(float));
    return pq;
}
void orc_pq_free(orc_pq *pq) {
    if (!pq) return;
    free(pq->ksub);
    free(pq->codebook);
    free(pq);
}

/* :28-58 Train: per subspace KMeansUtils.Train(sub, K, SubDim, L2, maxIter:10, seed:42+m) */
void orc_pq_train(orc_pq *pq, const float *data, int64_t n) {
    if (n == 0) return;
    for (int mi = 0; mi < pq->m; mi++) {
        float *cb = pq->codebook + (size_t)mi * pq->k * pq->sub;
        memset(cb, 0, sizeof(float) * (size_t)pq->k * pq->sub);
        pq->ksub[mi] = orc_kmeans_train(data + (size_t)mi * pq->sub, n, pq->sub, pq->dim, pq->k,
                                        ORC_L2, 10, 42 + mi, cb, NULL);
    }
    pq->trained = 1;
}
int orc_pq_ksub(const orc_pq *pq, int m) { return pq->ksub[m]; }
void orc_pq_get_codebook(const orc_pq *pq, float *out) {
    memcpy(out, pq->codebook, sizeof(float) * (size_t)pq->m * pq->k * pq->sub);
}
void orc_pq_set_codebook(orc_pq *pq, const float *cb, const int *ksub) {
    memcpy(pq->codebook, cb, sizeof(float) * (size_t)pq->m * pq->k * pq->sub);
    for (int i = 0; i < pq->m; i++) pq->ksub[i] = ksub ? ksub[i] : pq->k;
    pq->trained = 1;
}

/* :60-80 Encode + :122-136 FindNearest (strict '<' from float.MaxValue => lowest k wins) */
int orc_pq_encode(const orc_pq *pq, const float *vec, uint8_t *code) {
    if (!pq->trained) return -1;
    for (int mi = 0; mi < pq->m; mi++) {
        const float *sub = vec + (size_t)mi * pq->sub;
        const float *cb = pq->codebook + (size_t)mi * pq->k * pq->sub;
        float mind = 3.40282347e+38f;
        int best = 0;
        for (int k = 0; k < pq->ksub[mi]; k++) {
            float d = orc_l2sq_unsafe(sub, cb + (size_t)k * pq->sub, pq->sub);
            if (d < mind) {
                mind = d;
                best = k;
            }
        }
        code[mi] = (uint8_t)best;
    }
    return 0;
}

/* :98-120 ComputeDistanceTable: table[m][k] = L2SquaredUnsafe(subQuery_m, codeword_mk) */
void orc_pq_distance_table(const orc_pq *pq, const float *query, float *table) {
    for (int mi = 0; mi < pq->m; mi++) {
        const float *sub = query + (size_t)mi * pq->sub;
        const float *cb = pq->codebook + (size_t)mi * pq->k * pq->sub;
        for (int k = 0; k < pq->k; k++)
            table[(size_t)mi * pq->k + k] =
                k < pq->ksub[mi] ? orc_l2sq_unsafe(sub, cb + (size_t)k * pq->sub, pq->sub) : 0.0f;
    }
}

/* ============================================================================================
 * BruteForceVectorIndex.cs (FLAT).  EnableQuantization (SQ8) is out of scope.
 * ============================================================================================ */
struct orc_flat {
    int dim, metric;
    int count, cap; /* _vectors.Count */
    float *vecs;    /* [cap][dim] */
    float *norms;
    int64_t *ids;
    uint8_t *deleted;
    imap_t idmap;
};

orc_flat *orc_flat_new(int dim, int metric) {
    if (dim <= 0) return NULL; /* :44-47 */
    orc_flat *ix = (orc_flat *)calloc(1, sizeof *ix);
    ix->dim = dim;
    ix->metric = metric;
    ix->cap = 16;
    ix->vecs = (float *)malloc(sizeof(float) * 16 * (size_t)dim);
    ix->norms = (float *)malloc(sizeof(float) * 16);
    ix->ids = (int64_t *)malloc(sizeof(int64_t) * 16);
    ix->deleted = (uint8_t *)malloc(16);
    imap_init(&ix->idmap, 16);
    return ix;
}
void orc_flat_free(orc_flat *ix) {
    if (!ix) return;
    free(ix->vecs);
    free(ix->norms);
    free(ix->ids);
    free(ix->deleted);
    imap_free(&ix->idmap);
    free(ix);
}
/* :162-184 InternalAdd */
static void flat_internal_add(orc_flat *ix, int64_t id, const float *vec, float norm) {
    if (ix->count == ix->cap) {
        ix->cap *= 2;
        ix->vecs = (float *)realloc(ix->vecs, sizeof(float) * (size_t)ix->cap * (size_t)ix->dim);
        ix->norms = (float *)realloc(ix->norms, sizeof(float) * (size_t)ix->cap);
        ix->ids = (int64_t *)realloc(ix->ids, sizeof(int64_t) * (size_t)ix->cap);
        ix->deleted = (uint8_t *)realloc(ix->deleted, (size_t)ix->cap);
    }
    int idx = ix->count++;
    imap_put(&ix->idmap, id, idx);
    ix->ids[idx] = id;
    memcpy(ix->vecs + (size_t)idx * ix->dim, vec, sizeof(float) * (size_t)ix->dim);
    ix->norms[idx] = norm;
    ix->deleted[idx] = 0;
}
/* :133-160 Add */
int orc_flat_add(orc_flat *ix, int64_t id, const float *vec) {
    if (imap_get(&ix->idmap, id) >= 0) return -1;
    float norm = ix->metric == ORC_COSINE ? orc_norm(vec, ix->dim) : 0.0f;
    flat_internal_add(ix, id, vec, norm);
    return 0;
}
void orc_flat_add_batch(orc_flat *ix, int64_t n, const int64_t *ids, const float *X) {
    for (int64_t i = 0; i < n; i++) orc_flat_add(ix, ids ? ids[i] : i, X + (size_t)i * ix->dim);
}
/* :186-229 Upsert: existing id is overwritten in place and un-deleted */
void orc_flat_upsert(orc_flat *ix, int64_t id, const float *vec) {
    float norm = ix->metric == ORC_COSINE ? orc_norm(vec, ix->dim) : 0.0f;
    int idx = imap_get(&ix->idmap, id);
    if (idx >= 0) {
        memcpy(ix->vecs + (size_t)idx * ix->dim, vec, sizeof(float) * (size_t)ix->dim);
        ix->norms[idx] = norm;
        ix->deleted[idx] = 0;
    } else
        flat_internal_add(ix, id, vec, norm);
}
/* :231-254 Delete: tombstone + drop from id map */
int orc_flat_delete(orc_flat *ix, int64_t id) {
    int idx = imap_get(&ix->idmap, id);
    if (idx < 0) return 0;
    ix->deleted[idx] = 1;
    imap_del(&ix->idmap, id);
    return 1;
}
int orc_flat_count(const orc_flat *ix) { return ix->idmap.used; }

/* :275-379 Search (non-quantised branch :337-361) */
int orc_flat_search(const orc_flat *ix, const float *q, int topk, int64_t max_scans,
                    int64_t *ids_out, float *scores_out) {
    if (topk <= 0) return -2; /* ArgumentOutOfRangeException :278 */
    int count = ix->count;
    if (count == 0) return 0;
    int64_t scan_limit = max_scans >= 0 ? (max_scans < count ? max_scans : count) : count;
    /* max_scans <0 encodes "null"; a negative explicit value behaves as <=0 -> empty, which a
       caller expresses as 0 */
    if (scan_limit <= 0) return 0;
    pq4_t heap;
    pq4_init(&heap, topk + 1);
    int64_t scanned = 0;
    float qnorm = ix->metric == ORC_COSINE ? orc_norm(q, ix->dim) : 0.0f;
    for (int i = 0; i < count; i++) {
        if (ix->deleted[i]) continue;
        if (scanned >= scan_limit) break;
        scanned++;
        const float *v = ix->vecs + (size_t)i * ix->dim;
        float score;
        switch (ix->metric) {
        case ORC_L2: score = -orc_l2sq_unsafe(q, v, ix->dim); break;
        case ORC_IP: score = orc_dot_unsafe(q, v, ix->dim); break;
        default: {
            if (qnorm < 1e-6f || ix->norms[i] < 1e-6f)
                score = 0.0f;
            else {
                float den = qnorm * ix->norms[i];
                score = orc_dot_unsafe(q, v, ix->dim) / den;
            }
        }
        }
        res_t r = {ix->ids[i], score};
        pq4_enqueue(&heap, r);
        if (heap.size > topk) pq4_dequeue(&heap);
    }
    int n = drain_sorted(&heap, ids_out, scores_out);
    pq4_free(&heap);
    return n;
}

/* ============================================================================================
 * IvfFlatVectorIndex.cs
 * ============================================================================================ */
typedef struct {
    int64_t *ids;
    float *vecs;
    float *norms;
    int n, cap;
} flist_t;

struct orc_ivfflat {
    int dim, metric, nlist, nprobe;
    ndict_t buffer;
    int built;
    int nc; /* _centroids.Count */
    float *centroids, *cnorms;
    flist_t *lists; /* [nc] (Dictionary<int,List> keyed 0..nc-1 in key order) */
};

static float ivf_score(int metric, const float *q, const float *v, float qn, float vn, int dim) {
    /* IvfFlatVectorIndex.cs:351-360 ComputeScore / IvfPqVectorIndex.cs:214-224 */
    switch (metric) {
    case ORC_L2: return -orc_l2sq(q, v, dim);
    case ORC_IP: return orc_dot(q, v, dim);
    case ORC_COSINE: return orc_cosine(q, v, qn, vn, dim);
    default: return 0.0f;
    }
}

orc_ivfflat *orc_ivfflat_new(int dim, int metric, int nlist) {
    if (dim <= 0) return NULL;
    orc_ivfflat *ix = (orc_ivfflat *)calloc(1, sizeof *ix);
    ix->dim = dim;
    ix->metric = metric;
    ix->nlist = nlist;
    ix->nprobe = 3;
    ndict_init(&ix->buffer, dim);
    return ix;
}
static void ivfflat_free_lists(orc_ivfflat *ix) {
    for (int i = 0; i < ix->nc; i++) {
        free(ix->lists[i].ids);
        free(ix->lists[i].vecs);
        free(ix->lists[i].norms);
    }
    free(ix->lists);
    free(ix->centroids);
    free(ix->cnorms);
    ix->lists = NULL;
    ix->centroids = ix->cnorms = NULL;
    ix->nc = 0;
}
void orc_ivfflat_free(orc_ivfflat *ix) {
    if (!ix) return;
    ivfflat_free_lists(ix);
    ndict_free(&ix->buffer);
    free(ix);
}
void orc_ivfflat_set_nprobe(orc_ivfflat *ix, int nprobe) { ix->nprobe = nprobe; }
void orc_ivfflat_add(orc_ivfflat *ix, int64_t id, const float *vec) {
    float norm = ix->metric == ORC_COSINE ? orc_norm(vec, ix->dim) : 0.0f; /* CreateEntry :343-349 */
    ndict_set(&ix->buffer, id, vec, norm);
}
void orc_ivfflat_add_batch(orc_ivfflat *ix, int64_t n, const int64_t *ids, const float *X) {
    for (int64_t i = 0; i < n; i++) orc_ivfflat_add(ix, ids ? ids[i] : i, X + (size_t)i * ix->dim);
}
static void flist_push(flist_t *l, int dim, int64_t id, const float *v, float norm) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 8;
        l->ids = (int64_t *)realloc(l->ids, sizeof(int64_t) * (size_t)l->cap);
        l->vecs = (float *)realloc(l->vecs, sizeof(float) * (size_t)l->cap * (size_t)dim);
        l->norms = (float *)realloc(l->norms, sizeof(float) * (size_t)l->cap);
    }
    l->ids[l->n] = id;
    memcpy(l->vecs + (size_t)l->n * dim, v, sizeof(float) * (size_t)dim);
    l->norms[l->n] = norm;
    l->n++;
}
/* :62-83 Delete: buffer.Remove + RemoveAll in every list */
int orc_ivfflat_delete(orc_ivfflat *ix, int64_t id) {
    int removed = ndict_remove(&ix->buffer, id);
    if (ix->built) {
        for (int c = 0; c < ix->nc; c++) {
            flist_t *l = &ix->lists[c];
            int w = 0;
            for (int j = 0; j < l->n; j++) {
                if (l->ids[j] == id) {
                    removed = 1;
                    continue;
                }
                if (w != j) {
                    l->ids[w] = l->ids[j];
                    l->norms[w] = l->norms[j];
                    memmove(l->vecs + (size_t)w * ix->dim, l->vecs + (size_t)j * ix->dim,
                            sizeof(float) * (size_t)ix->dim);
                }
                w++;
            }
            l->n = w;
        }
    }
    return removed;
}

/* :85-145 Build */
void orc_ivfflat_build(orc_ivfflat *ix) {
    int dim = ix->dim;
    /* 1. uniqueData: existing list items (list key order, item order), then buffer overrides */
    ndict_t uniq;
    ndict_init(&uniq, dim);
    if (ix->built)
        for (int c = 0; c < ix->nc; c++)
            for (int j = 0; j < ix->lists[c].n; j++)
                ndict_set(&uniq, ix->lists[c].ids[j], ix->lists[c].vecs + (size_t)j * dim,
                          ix->lists[c].norms[j]);
    for (int s = 0; s < ix->buffer.count; s++)
        if (ndict_is_live(&ix->buffer, s))
            ndict_set(&uniq, ix->buffer.ids[s], ix->buffer.vecs + (size_t)s * dim,
                      ix->buffer.norms[s]);
    int n = ndict_live(&uniq); /* no removals in uniq: slots 0..n-1 dense */
    if (n == 0) {
        ndict_free(&uniq);
        return;
    }
    /* 2. Train */
    int k = ix->nlist < n ? ix->nlist : n;
    if (k <= 0) k = 1;
    float *cent = (float *)malloc(sizeof(float) * (size_t)k * (size_t)dim);
    k = orc_kmeans_train(uniq.vecs, n, dim, dim, k, ix->metric, 10, 42, cent, NULL);
    float *cn = (float *)malloc(sizeof(float) * (size_t)k);
    for (int c = 0; c < k; c++) cn[c] = ix->metric == ORC_COSINE ? orc_norm(cent + (size_t)c * dim, dim) : 0.0f;
    /* 3. Assign */
    flist_t *lists = (flist_t *)calloc((size_t)k, sizeof(flist_t));
    int32_t *assign = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++)
        assign[i] = find_nearest_centroid_ld(uniq.vecs + (size_t)i * dim, cent, dim, cn, k, dim, ix->metric);
    for (int i = 0; i < n; i++)
        flist_push(&lists[assign[i]], dim, uniq.ids[i], uniq.vecs + (size_t)i * dim, uniq.norms[i]);
    free(assign);
    /* 4. Commit */
    ivfflat_free_lists(ix);
    ix->centroids = cent;
    ix->cnorms = cn;
    ix->lists = lists;
    ix->nc = k;
    ndict_clear(&ix->buffer);
    ix->built = 1;
    ndict_free(&uniq);
}
int orc_ivfflat_is_built(const orc_ivfflat *ix) { return ix->built; }
int orc_ivfflat_ncentroids(const orc_ivfflat *ix) { return ix->nc; }
void orc_ivfflat_get_centroids(const orc_ivfflat *ix, float *out) {
    memcpy(out, ix->centroids, sizeof(float) * (size_t)ix->nc * (size_t)ix->dim);
}
int orc_ivfflat_count(const orc_ivfflat *ix) {
    int c = ndict_live(&ix->buffer);
    for (int i = 0; i < ix->nc; i++) c += ix->lists[i].n;
    return c;
}
int orc_ivfflat_list_size(const orc_ivfflat *ix, int list) { return ix->lists[list].n; }
void orc_ivfflat_get_list(const orc_ivfflat *ix, int list, int64_t *ids_out) {
    memcpy(ids_out, ix->lists[list].ids, sizeof(int64_t) * (size_t)ix->lists[list].n);
}

/* Test infrastructure: take over centroids and inverted lists built elsewhere (the GPU index under test), so that a
 * search compares the scan alone at sizes where re-running k-means on the CPU would take hours.  vecs is list-major like
 * ids; norms are recomputed the way Add does (:347). */
void orc_ivfflat_adopt(orc_ivfflat *ix, int nlist, const float *centroids, const int64_t *offs, const int64_t *ids,
                       const float *vecs) {
    int dim = ix->dim;
    ivfflat_free_lists(ix);
    ix->nc = nlist;
    ix->centroids = (float *)malloc(sizeof(float) * (size_t)nlist * (size_t)dim);
    memcpy(ix->centroids, centroids, sizeof(float) * (size_t)nlist * (size_t)dim);
    ix->cnorms = (float *)malloc(sizeof(float) * (size_t)nlist);
    for (int c = 0; c < nlist; c++)
        ix->cnorms[c] = ix->metric == ORC_COSINE ? orc_norm(ix->centroids + (size_t)c * dim, dim) : 0.0f;
    ix->lists = (flist_t *)calloc((size_t)nlist, sizeof(flist_t));
    for (int c = 0; c < nlist; c++) {
        int64_t n = offs[c + 1] - offs[c];
        flist_t *l = &ix->lists[c];
        l->n = l->cap = (int)n;
        if (n == 0) continue;
        l->ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
        l->vecs = (float *)malloc(sizeof(float) * (size_t)n * (size_t)dim);
        l->norms = (float *)malloc(sizeof(float) * (size_t)n);
        memcpy(l->ids, ids + offs[c], sizeof(int64_t) * (size_t)n);
        memcpy(l->vecs, vecs + (size_t)offs[c] * (size_t)dim, sizeof(float) * (size_t)n * (size_t)dim);
        for (int64_t i = 0; i < n; i++)
            l->norms[i] = ix->metric == ORC_COSINE ? orc_norm(l->vecs + (size_t)i * dim, dim) : 0.0f;
    }
    ndict_clear(&ix->buffer);
    ix->built = 1;
}

/* rank all centroids: score each (:186-193), List.Sort descending (:196) */
static res_t *rank_centroids(int metric, const float *q, float qn, const float *cent,
                             const float *cnorms, int nc, int dim) {
    res_t *cs = (res_t *)malloc(sizeof(res_t) * (size_t)(nc > 0 ? nc : 1));
    for (int i = 0; i < nc; i++) {
        cs[i].id = i;
        cs[i].score = ivf_score(metric, q, cent + (size_t)i * dim, qn, cnorms[i], dim);
    }
    dn_sort_desc(cs, nc);
    return cs;
}

/* :147-231 Search */
int orc_ivfflat_search(const orc_ivfflat *ix, const float *q, int topk, int64_t max_scans_opt,
                       int nprobe_opt, int64_t *ids_out, float *scores_out) {
    int dim = ix->dim;
    int nprobe = nprobe_opt >= 0 ? nprobe_opt : ix->nprobe;
    int64_t max_scans = max_scans_opt >= 0 ? max_scans_opt : INT32_MAX;
    pq4_t heap;
    pq4_init(&heap, (topk > 0 ? topk : 0) + 1);
    int64_t scanned = 0;
    float qn = ix->metric == ORC_COSINE ? orc_norm(q, dim) : 0.0f;
    /* 1. buffer (exact) :170-180 */
    for (int s = 0; s < ix->buffer.count; s++) {
        if (!ndict_is_live(&ix->buffer, s)) continue;
        if (scanned >= max_scans) break;
        scanned++;
        float score = ivf_score(ix->metric, q, ix->buffer.vecs + (size_t)s * dim, qn,
                                ix->buffer.norms[s], dim);
        res_t r = {ix->buffer.ids[s], score};
        pq4_enqueue(&heap, r);
        if (heap.size > topk) pq4_dequeue(&heap);
    }
    /* 2. index :183-219 */
    if (ix->built && ix->nc > 0 && scanned < max_scans) {
        res_t *cs = rank_centroids(ix->metric, q, qn, ix->centroids, ix->cnorms, ix->nc, dim);
        int probes = nprobe < ix->nc ? nprobe : ix->nc;
        for (int i = 0; i < probes; i++) {
            if (scanned >= max_scans) break;
            const flist_t *l = &ix->lists[cs[i].id];
            for (int j = 0; j < l->n; j++) {
                if (scanned >= max_scans) break;
                if (ndict_contains(&ix->buffer, l->ids[j])) continue; /* seenIds :210 */
                scanned++;
                float score = ivf_score(ix->metric, q, l->vecs + (size_t)j * dim, qn, l->norms[j], dim);
                res_t r = {l->ids[j], score};
                pq4_enqueue(&heap, r);
                if (heap.size > topk) pq4_dequeue(&heap);
            }
        }
        free(cs);
    }
    int n = drain_sorted(&heap, ids_out, scores_out);
    pq4_free(&heap);
    return n;
}

/* ============================================================================================
 * IvfPqVectorIndex.cs
 * ============================================================================================ */
typedef struct {
    int64_t *ids;
    uint8_t *codes;
    int n, cap;
} plist_t;

struct orc_ivfpq {
    int dim, metric, m, k, nlist;
    orc_pq *pq;
    ndict_t buffer;
    int built, nc;
    float *centroids, *cnorms;
    plist_t *lists;
};

orc_ivfpq *orc_ivfpq_new(int dim, int metric, int m, int k, int nlist) {
    orc_pq *pq = orc_pq_new(dim, m, k);
    if (!pq) return NULL;
    orc_ivfpq *ix = (orc_ivfpq *)calloc(1, sizeof *ix);
    ix->dim = dim;
    ix->metric = metric;
    ix->m = m;
    ix->k = k;
    ix->nlist = nlist;
    ix->pq = pq;
    ndict_init(&ix->buffer, dim);
    return ix;
}
static void ivfpq_free_lists(orc_ivfpq *ix) {
    for (int i = 0; i < ix->nc; i++) {
        free(ix->lists[i].ids);
        free(ix->lists[i].codes);
    }
    free(ix->lists);
    free(ix->centroids);
    free(ix->cnorms);
    ix->lists = NULL;
    ix->centroids = ix->cnorms = NULL;
    ix->nc = 0;
}
void orc_ivfpq_free(orc_ivfpq *ix) {
    if (!ix) return;
    ivfpq_free_lists(ix);
    ndict_free(&ix->buffer);
    orc_pq_free(ix->pq);
    free(ix);
}
void orc_ivfpq_add(orc_ivfpq *ix, int64_t id, const float *vec) { ndict_set(&ix->buffer, id, vec, 0.0f); }
void orc_ivfpq_add_batch(orc_ivfpq *ix, int64_t n, const int64_t *ids, const float *X) {
    for (int64_t i = 0; i < n; i++) orc_ivfpq_add(ix, ids ? ids[i] : i, X + (size_t)i * ix->dim);
}
int orc_ivfpq_delete(orc_ivfpq *ix, int64_t id) { return ndict_remove(&ix->buffer, id); }

static void plist_push(plist_t *l, int m, int64_t id, const uint8_t *code) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 8;
        l->ids = (int64_t *)realloc(l->ids, sizeof(int64_t) * (size_t)l->cap);
        l->codes = (uint8_t *)realloc(l->codes, (size_t)l->cap * (size_t)m);
    }
    l->ids[l->n] = id;
    memcpy(l->codes + (size_t)l->n * m, code, (size_t)m);
    l->n++;
}

/* :55-116 Build (buffer only; replaces lists) */
void orc_ivfpq_build(orc_ivfpq *ix) {
    int dim = ix->dim;
    int n = ndict_live(&ix->buffer);
    if (n == 0) return; /* :62, :65 */
    /* buffer.Values.ToList() in enumeration order */
    float *all = (float *)malloc(sizeof(float) * (size_t)n * (size_t)dim);
    int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int w = 0;
    for (int s = 0; s < ix->buffer.count; s++)
        if (ndict_is_live(&ix->buffer, s)) {
            memcpy(all + (size_t)w * dim, ix->buffer.vecs + (size_t)s * dim, sizeof(float) * (size_t)dim);
            ids[w++] = ix->buffer.ids[s];
        }
    int nc = ix->nlist < n ? ix->nlist : n;
    float *cent = (float *)malloc(sizeof(float) * (size_t)(nc > 0 ? nc : 1) * (size_t)dim);
    nc = orc_kmeans_train(all, n, dim, dim, nc, ix->metric, 10, 123, cent, NULL); /* :69 */
    float *cn = (float *)malloc(sizeof(float) * (size_t)nc);
    for (int c = 0; c < nc; c++) cn[c] = ix->metric == ORC_COSINE ? orc_norm(cent + (size_t)c * dim, dim) : 0.0f;
    /* residuals :73-86 */
    float *res = (float *)malloc(sizeof(float) * (size_t)n * (size_t)dim);
    int32_t *assign = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float *v = all + (size_t)i * dim;
        int c = find_nearest_centroid_ld(v, cent, dim, cn, nc, dim, ix->metric);
        assign[i] = c;
        for (int d = 0; d < dim; d++) res[(size_t)i * dim + d] = v[d] - cent[(size_t)c * dim + d];
    }
    orc_pq_train(ix->pq, res, n); /* :89 */
    /* encode + populate :92-107 */
    ivfpq_free_lists(ix);
    ix->lists = (plist_t *)calloc((size_t)nc, sizeof(plist_t));
    ix->nc = nc;
    ix->centroids = cent;
    ix->cnorms = cn;
    uint8_t *codes = (uint8_t *)malloc((size_t)n * (size_t)ix->m);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) orc_pq_encode(ix->pq, res + (size_t)i * dim, codes + (size_t)i * ix->m);
    for (int i = 0; i < n; i++) plist_push(&ix->lists[assign[i]], ix->m, ids[i], codes + (size_t)i * ix->m);
    free(codes);
    free(res);
    free(assign);
    free(all);
    free(ids);
    ndict_clear(&ix->buffer);
    ix->built = 1;
}
int orc_ivfpq_is_built(const orc_ivfpq *ix) { return ix->built; }
int orc_ivfpq_ncentroids(const orc_ivfpq *ix) { return ix->nc; }
void orc_ivfpq_get_centroids(const orc_ivfpq *ix, float *out) {
    memcpy(out, ix->centroids, sizeof(float) * (size_t)ix->nc * (size_t)ix->dim);
}
const orc_pq *orc_ivfpq_pq(const orc_ivfpq *ix) { return ix->pq; }
int orc_ivfpq_list_size(const orc_ivfpq *ix, int list) { return ix->lists[list].n; }
void orc_ivfpq_get_list(const orc_ivfpq *ix, int list, int64_t *ids_out, uint8_t *codes_out) {
    const plist_t *l = &ix->lists[list];
    if (ids_out) memcpy(ids_out, l->ids, sizeof(int64_t) * (size_t)l->n);
    if (codes_out) memcpy(codes_out, l->codes, (size_t)l->n * (size_t)ix->m);
}

void orc_ivfpq_adopt(orc_ivfpq *ix, int nlist, const float *centroids, const float *codebook,
                     const int64_t *offs, const int64_t *ids, const uint8_t *codes) {
    int dim = ix->dim;
    ivfpq_free_lists(ix);
    ix->nc = nlist;
    ix->centroids = (float *)malloc(sizeof(float) * (size_t)nlist * (size_t)dim);
    memcpy(ix->centroids, centroids, sizeof(float) * (size_t)nlist * (size_t)dim);
    ix->cnorms = (float *)malloc(sizeof(float) * (size_t)nlist);
    for (int c = 0; c < nlist; c++)
        ix->cnorms[c] = ix->metric == ORC_COSINE ? orc_norm(ix->centroids + (size_t)c * dim, dim) : 0.0f;
    orc_pq_set_codebook(ix->pq, codebook, NULL);
    ix->lists = (plist_t *)calloc((size_t)nlist, sizeof(plist_t));
    for (int c = 0; c < nlist; c++) {
        int64_t n = offs[c + 1] - offs[c];
        plist_t *l = &ix->lists[c];
        l->n = l->cap = (int)n;
        if (n == 0) continue;
        l->ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
        l->codes = (uint8_t *)malloc((size_t)n * (size_t)ix->m);
        memcpy(l->ids, ids + offs[c], sizeof(int64_t) * (size_t)n);
        memcpy(l->codes, codes + (size_t)offs[c] * (size_t)ix->m, (size_t)n * (size_t)ix->m);
    }
    ix->built = 1;
}

/* :118-212 Search */
int orc_ivfpq_search(const orc_ivfpq *ix, const float *q, int topk, int nprobe_opt,
                     int64_t *ids_out, float *scores_out) {
    int dim = ix->dim, m = ix->m, K = ix->k;
    int nprobe = nprobe_opt >= 0 ? nprobe_opt : 1; /* :125 */
    pq4_t heap;
    pq4_init(&heap, (topk > 0 ? topk : 0) + 1);
    float qn = ix->metric == ORC_COSINE ? orc_norm(q, dim) : 0.0f;
    /* 1. buffer exact :130-136 (ComputeScore with vNorm=-1 => recomputed for Cosine) */
    for (int s = 0; s < ix->buffer.count; s++) {
        if (!ndict_is_live(&ix->buffer, s)) continue;
        const float *v = ix->buffer.vecs + (size_t)s * dim;
        float vn = ix->metric == ORC_COSINE ? orc_norm(v, dim) : 0.0f;
        float score = ivf_score(ix->metric, q, v, qn, vn, dim);
        res_t r = {ix->buffer.ids[s], score};
        pq4_enqueue(&heap, r);
        if (heap.size > topk) pq4_dequeue(&heap);
    }
    if (ix->built) {
        res_t *cs = rank_centroids(ix->metric, q, qn, ix->centroids, ix->cnorms, ix->nc, dim);
        int probes = nprobe < ix->nc ? nprobe : ix->nc;
        float *resq = (float *)malloc(sizeof(float) * (size_t)dim);
        float *table = (float *)malloc(sizeof(float) * (size_t)m * (size_t)K);
        for (int i = 0; i < probes; i++) {
            int c = (int)cs[i].id;
            const plist_t *l = &ix->lists[c];
            if (l->n == 0) continue;
            const float *cen = ix->centroids + (size_t)c * dim;
            for (int d = 0; d < dim; d++) resq[d] = q[d] - cen[d]; /* :161-163 */
            orc_pq_distance_table(ix->pq, resq, table);            /* :166 */
            for (int j = 0; j < l->n; j++) {
                if (ndict_contains(&ix->buffer, l->ids[j])) continue; /* seen :170 */
                const uint8_t *code = l->codes + (size_t)j * m;
                float dist = 0.0f;
                for (int mi = 0; mi < m; mi++) dist += table[(size_t)mi * K + code[mi]]; /* :182-186 */
                res_t r = {l->ids[j], -dist}; /* :194 */
                pq4_enqueue(&heap, r);
                if (heap.size > topk) pq4_dequeue(&heap);
            }
        }
        free(resq);
        free(table);
        free(cs);
    }
    int n = drain_sorted(&heap, ids_out, scores_out);
    pq4_free(&heap);
    return n;
}

/* ============================================================================================
 * DeltaVectorIndex.cs:95-121 — merge (tail first, head overwrites by id), sort desc, Take(k)
 * ============================================================================================ */
int orc_delta_merge(const int64_t *hid, const float *hs, int nh, const int64_t *tid,
                    const float *ts, int nt, int topk, int64_t *ids_out, float *scores_out) {
    int cap = nh + nt;
    res_t *merged = (res_t *)malloc(sizeof(res_t) * (size_t)(cap > 0 ? cap : 1));
    int n = 0;
    /* Dictionary insertion order: tail entries first; a head hit on an existing key overwrites in place */
    for (int i = 0; i < nt; i++) {
        int found = -1;
        for (int j = 0; j < n; j++)
            if (merged[j].id == tid[i]) found = j;
        if (found >= 0)
            merged[found].score = ts[i];
        else {
            merged[n].id = tid[i];
            merged[n++].score = ts[i];
        }
    }
    for (int i = 0; i < nh; i++) {
        int found = -1;
        for (int j = 0; j < n; j++)
            if (merged[j].id == hid[i]) found = j;
        if (found >= 0)
            merged[found].score = hs[i];
        else {
            merged[n].id = hid[i];
            merged[n++].score = hs[i];
        }
    }
    dn_sort_desc(merged, n);
    int out = n < topk ? n : (topk > 0 ? topk : 0);
    for (int i = 0; i < out; i++) {
        ids_out[i] = merged[i].id;
        scores_out[i] = merged[i].score;
    }
    free(merged);
    return out;
}

/* ============================================================================================
 * Batched driver: one query per thread (Garnet session threads under read locks)
 * ============================================================================================ */
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_search_batch(int kind, const void *ix, const float *Q, int64_t nq, int topk,
                      int64_t max_scans, int nprobe, int nthreads, int64_t *ids_out,
                      float *scores_out, int32_t *counts_out) {
    int dim = kind == 0 ? ((const orc_flat *)ix)->dim
                        : kind == 1 ? ((const orc_ivfflat *)ix)->dim : ((const orc_ivfpq *)ix)->dim;
    if (nthreads <= 0) nthreads = orc_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t i = 0; i < nq; i++) {
        const float *q = Q + (size_t)i * dim;
        int64_t *io = ids_out + (size_t)i * topk;
        float *so = scores_out + (size_t)i * topk;
        int n;
        if (kind == 0)
            n = orc_flat_search((const orc_flat *)ix, q, topk, max_scans, io, so);
        else if (kind == 1)
            n = orc_ivfflat_search((const orc_ivfflat *)ix, q, topk, max_scans, nprobe, io, so);
        else
            n = orc_ivfpq_search((const orc_ivfpq *)ix, q, topk, nprobe, io, so);
        counts_out[i] = n;
    }
}
