/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C restatement of the reference's kd-tree build + kNN query algorithm
 * (wendazhou/nbodyhpc, kdtree/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) oracle/_ref/libnbref.so = the reference's own sources compiled unmodified (oracle/Makefile),
 *       on the reference's test fixtures (tests/test.cpp:43-111, tests/test_inserters.cpp:120-121)
 *       and on random inputs: node arrays byte-identical, distances bit-identical, counters equal;
 *   (2) the committed golden vectors under tests/golden/ (generated from (1) by
 *       tests/golden/make_golden.py), so the pin also holds where /root/reference is absent.
 *
 * Everything is float32 and FMA-free (built with -ffp-contract=off and without -mfma), mirroring
 * the reference build (kdtree/CMakeLists.txt:74: -mavx2 only).
 *
 * Tie semantics.  The reference inserts a point only if d2 < current k-th best (strict), so among
 * exact d2 ties at the k-th boundary the first VISITED point wins, and it finally sorts by distance
 * only (kdtree_opt.hpp:13-18, kdtree.cpp:151).  The visiting order inside a leaf depends on the
 * reference's Floyd-Rivest/AVX partition, which is not restated here.  The tree-walking functions
 * below therefore reproduce the reference exactly up to the identity of exactly-tied points; rows
 * are returned in canonical (d2, index) order.  orc_brute_force defines the exact top-k under the
 * total order (d2, index) that the B200 product implements (north star: "exact distance ties
 * ordered by index").
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t dim;    /* -1 = leaf                                   kdtree.hpp:149-163 */
    float split;    /* split coordinate (internal nodes)                              */
    uint32_t left;  /* leaf: first point; internal: left child node index             */
    uint32_t right; /* leaf: one past last point; internal: right child node index    */
} orc_node;

typedef struct {
    uint64_t n_pad;
    uint64_t n_nodes, cap_nodes;
    int leaf_size; /* effective: max(leaf_size, 2*block)           kdtree_impl.hpp:91 */
    int block;
    int periodic;
    float box;
    float *x, *y, *z;
    uint32_t *idx;
    orc_node *nodes;
} orc_tree;

/* ------------------------------------------------------------------------------------------ */
/* Fixtures: Philox4x32-10 (Random123, published algorithm) + u01<float>                        */
/* kdtree_utils.hpp:16-46: key = {seed, 0}, counter = {dim, i, 0, 0}, lane 0,                   */
/* u01(x) = float(x) * 2^-32 + 2^-33 (Random123/uniform.hpp:174-184), times boxsize.           */
/* ------------------------------------------------------------------------------------------ */
static uint32_t philox4x32_10_lane0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                    uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

void orc_philox_points(uint32_t n, uint32_t seed, float boxsize, float *out_aos) {
    const float factor = 1.0f / 4294967296.0f, half = 0.5f * factor;
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t d = 0; d < 3; ++d) {
            uint32_t r = philox4x32_10_lane0(d, i, 0, 0, seed, 0);
            volatile float u = (float)r * factor;
            out_aos[3 * (size_t)i + d] = (u + half) * boxsize;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Metrics                                                                 kdtree.hpp:20-121  */
/* ------------------------------------------------------------------------------------------ */
static inline float min3f(float a, float b, float c) {
    /* std::min({a,b,c}) returns the first minimum; values only matter here */
    float m = a;
    if (b < m) m = b;
    if (c < m) m = c;
    return m;
}

/* L2Distance::operator() kdtree.hpp:22-31 / L2PeriodicDistance::operator() kdtree.hpp:71-84.
 * Operand order follows the leaf kernels (point - query: kdtree_asm_systemv.asm:76-87,
 * kdtree_opt.hpp:139-145); squares make the sign irrelevant. */
static inline float point_d2(const float p[3], const float q[3], int periodic, float box) {
    float r = 0.0f;
    for (int i = 0; i < 3; ++i) {
        float d = p[i] - q[i];
        if (periodic) {
            float dp = d + box, dm = d - box;
            r += min3f(d * d, dp * dp, dm * dm);
        } else {
            r += d * d;
        }
    }
    return r;
}

float orc_point_distance(const float *p, const float *q, float boxsize) {
    return point_d2(p, q, boxsize >= 0, boxsize);
}

/* L2Distance::box_distance kdtree.hpp:34-45 / L2PeriodicDistance::box_distance kdtree.hpp:88-107.
 * box = {lo0, hi0, lo1, hi1, lo2, hi2}. */
static inline float box_d2(const float q[3], const float b[6], int periodic, float box) {
    float r = 0.0f;
    for (int i = 0; i < 3; ++i) {
        if (!periodic) {
            float dl = b[2 * i] - q[i];
            if (!(dl > 0.0f)) dl = 0.0f; /* std::max(x, 0) */
            float dr = q[i] - b[2 * i + 1];
            if (!(dr > 0.0f)) dr = 0.0f;
            r += dl * dl + dr * dr;
        } else if (q[i] < b[2 * i]) {
            float d = b[2 * i] - q[i];
            float dw = q[i] + box - b[2 * i + 1];
            float m = dw < d ? dw : d;
            r += m * m;
        } else if (q[i] > b[2 * i + 1]) {
            float d = q[i] - b[2 * i + 1];
            float dw = b[2 * i] + box - q[i];
            float m = dw < d ? dw : d;
            r += m * m;
        }
    }
    return r;
}

float orc_box_distance(const float *q, const float *box6, float boxsize) {
    return box_d2(q, box6, boxsize >= 0, boxsize);
}

/* ------------------------------------------------------------------------------------------ */
/* Build                                                                                        */
/* ------------------------------------------------------------------------------------------ */
static inline void swap_pts(orc_tree *t, size_t i, size_t j) {
    float f; uint32_t u;
    f = t->x[i]; t->x[i] = t->x[j]; t->x[j] = f;
    f = t->y[i]; t->y[i] = t->y[j]; t->y[j] = f;
    f = t->z[i]; t->z[i] = t->z[j]; t->z[j] = f;
    u = t->idx[i]; t->idx[i] = t->idx[j]; t->idx[j] = u;
}

/* Selection: after the call, position `nth` holds the element of rank nth-lo of [lo,hi) along
 * `key`, everything before it is <= and everything after it is >=.  This is the contract of the
 * reference's selection policies (kdtree_impl.hpp:31-52, kdtree_build_opt.hpp:36-60); which of
 * several equal-coordinate points lands on which side is NOT part of that contract, and the
 * reference's Floyd-Rivest permutation is not restated. */
static void select_nth(orc_tree *t, float *key, int64_t lo, int64_t hi, int64_t nth) {
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2, last = hi - 1;
        /* pivot value = median of three */
        float a = key[lo], b = key[mid], c = key[last];
        float pivot = a < b ? (b < c ? b : (a < c ? c : a)) : (a < c ? a : (b < c ? c : b));
        /* three-way partition: [lo,lt) < pivot, [lt,gt] == pivot, (gt,hi) > pivot */
        int64_t lt = lo, i = lo, gt = last;
        while (i <= gt) {
            if (key[i] < pivot) { swap_pts(t, (size_t)lt, (size_t)i); ++lt; ++i; }
            else if (pivot < key[i]) { swap_pts(t, (size_t)i, (size_t)gt); --gt; }
            else ++i;
        }
        if (nth < lt) hi = lt;
        else if (nth > gt) lo = gt + 1;
        else return;
    }
}

static int push_node(orc_tree *t, orc_node nd, uint32_t *out) {
    if (t->n_nodes == t->cap_nodes) {
        uint64_t cap = t->cap_nodes ? 2 * t->cap_nodes : 64;
        orc_node *p = (orc_node *)realloc(t->nodes, cap * sizeof(orc_node));
        if (!p) return 1;
        t->nodes = p; t->cap_nodes = cap;
    }
    *out = (uint32_t)t->n_nodes;
    t->nodes[t->n_nodes++] = nd;
    return 0;
}

/* KDTreeBuilder::build_node kdtree_impl.hpp:98-146 (+ build_left_right_nonthreaded :148-157):
 * leaf iff count <= leaf_size_; median_offset = ((count/2)/block)*block; split = coordinate of the
 * rank-median_offset element; node pushed BEFORE its children (pre-order, left subtree first);
 * children cycle the dimension. */
static uint32_t build_node(orc_tree *t, int dim, uint32_t left, uint32_t count) {
    uint32_t me = 0;
    if (count <= (uint32_t)t->leaf_size) {
        orc_node leaf = {-1, 0.0f, left, left + count};
        push_node(t, leaf, &me);
        return me;
    }
    uint32_t median = (count / 2 / (uint32_t)t->block) * (uint32_t)t->block;
    float *key = dim == 0 ? t->x : (dim == 1 ? t->y : t->z);
    select_nth(t, key, (int64_t)left, (int64_t)left + count, (int64_t)left + median);
    orc_node nd = {dim, key[left + median], 0u, 0u};
    push_node(t, nd, &me);
    uint32_t l = build_node(t, (dim + 1) % 3, left, median);
    uint32_t r = build_node(t, (dim + 1) % 3, left + median, count - median);
    t->nodes[me].left = l;
    t->nodes[me].right = r;
    return me;
}

/* Pad + AoS->SoA + in-box validation (pybind.cpp:14-56, kdtree.cpp:64-90), argument checks
 * (kdtree.cpp:98-108), then the recursive builder.  boxsize < 0 => open boundaries.
 * status: 0 ok, 1 out-of-box point, 2 bad block size, 3 allocation failure. */
orc_tree *orc_tree_build(const float *xyz_aos, uint64_t n, int leaf_size, int block, float boxsize,
                         int *status) {
    int st = 0;
    if (block <= 0 || block % 8 != 0) { if (status) *status = 2; return NULL; }
    orc_tree *t = (orc_tree *)calloc(1, sizeof(orc_tree));
    uint64_t n_pad = (n + (uint64_t)block - 1) / (uint64_t)block * (uint64_t)block;
    t->n_pad = n_pad;
    t->block = block;
    t->leaf_size = leaf_size > 2 * block ? leaf_size : 2 * block;
    t->periodic = boxsize >= 0;
    t->box = boxsize >= 0 ? boxsize : 0.0f;
    size_t alloc = n_pad ? n_pad : 1;
    t->x = (float *)malloc(alloc * 4); t->y = (float *)malloc(alloc * 4);
    t->z = (float *)malloc(alloc * 4); t->idx = (uint32_t *)malloc(alloc * 4);
    if (!t->x || !t->y || !t->z || !t->idx) st = 3;
    for (uint64_t i = 0; !st && i < n_pad; ++i) {
        t->idx[i] = (uint32_t)i; /* iota over the PADDED length (pybind.cpp:27) */
        if (i < n) {
            t->x[i] = xyz_aos[3 * i]; t->y[i] = xyz_aos[3 * i + 1]; t->z[i] = xyz_aos[3 * i + 2];
            if (t->periodic) {
                for (int d = 0; d < 3; ++d) {
                    float v = xyz_aos[3 * i + d];
                    if (!(v >= 0.0f && v <= boxsize)) st = 1;
                }
            }
        } else {
            t->x[i] = t->y[i] = t->z[i] = FLT_MAX;
        }
    }
    if (!st) build_node(t, 0, 0, (uint32_t)n_pad);
    if (status) *status = st;
    if (st) {
        free(t->x); free(t->y); free(t->z); free(t->idx); free(t->nodes); free(t);
        return NULL;
    }
    return t;
}

void orc_tree_free(orc_tree *t) {
    if (!t) return;
    free(t->x); free(t->y); free(t->z); free(t->idx); free(t->nodes); free(t);
}

uint64_t orc_tree_num_points(const orc_tree *t) { return t->n_pad; }
uint64_t orc_tree_num_nodes(const orc_tree *t) { return t->n_nodes; }
void orc_tree_copy_nodes(const orc_tree *t, void *nodes16) {
    memcpy(nodes16, t->nodes, t->n_nodes * sizeof(orc_node));
}
void orc_tree_copy_points(const orc_tree *t, float *x, float *y, float *z, uint32_t *idx) {
    memcpy(x, t->x, t->n_pad * 4); memcpy(y, t->y, t->n_pad * 4);
    memcpy(z, t->z, t->n_pad * 4); memcpy(idx, t->idx, t->n_pad * 4);
}

/* Tree topology without any point data: number of nodes the reference creates for n_pad points
 * (it depends on counts only, kdtree_impl.hpp:101-110). */
static uint64_t count_nodes(uint64_t count, uint64_t leaf, uint64_t block) {
    if (count <= leaf) return 1;
    uint64_t m = count / 2 / block * block;
    return 1 + count_nodes(m, leaf, block) + count_nodes(count - m, leaf, block);
}
uint64_t orc_expected_num_nodes(uint64_t n, int leaf_size, int block) {
    uint64_t n_pad = (n + (uint64_t)block - 1) / (uint64_t)block * (uint64_t)block;
    uint64_t leaf = (uint64_t)(leaf_size > 2 * block ? leaf_size : 2 * block);
    return count_nodes(n_pad, leaf, (uint64_t)block);
}

/* ------------------------------------------------------------------------------------------ */
/* Top-k queue.  Restates the CONTRACT of TournamentTree (tournament_tree.hpp:42-105): k slots   */
/* initialised to {FLT_MAX, 0xFFFFFFFF} (kdtree_impl.hpp:206-210), top() = a slot of maximum    */
/* distance, replace_top() overwrites it.  Kept as a plain array with a tracked maximum.        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float d; uint32_t i; } orc_pair;

typedef struct {
    const orc_tree *t;
    float q[3];
    int k;
    orc_pair *best;
    int top; /* slot holding the current maximum */
    uint64_t nodes_visited, nodes_pruned, points_visited;
} orc_search;

static inline void find_top(orc_search *s) {
    int top = 0;
    for (int j = 1; j < s->k; ++j)
        if (s->best[top].d < s->best[j].d) top = j;
    s->top = top;
}

/* InsertShorterDistanceVanilla kdtree_opt.hpp:20-44 == the asm/AVX leaf kernels' semantics
 * (kdtree_asm_systemv.asm:121-189): insert iff d2 < current top (strict). */
static void process_leaf(orc_search *s, const orc_node *nd) {
    const orc_tree *t = s->t;
    for (uint32_t p = nd->left; p < nd->right; ++p) {
        float pt[3] = {t->x[p], t->y[p], t->z[p]};
        float d = point_d2(pt, s->q, t->periodic, t->box);
        if (d >= s->best[s->top].d) continue;
        s->best[s->top].d = d;
        s->best[s->top].i = t->idx[p];
        find_top(s);
    }
    s->points_visited += nd->right - nd->left; /* kdtree_impl.hpp:219 */
}

/* KDTreeQuery::compute kdtree_impl.hpp:226-268 */
static void compute(orc_search *s, const orc_node *nd, const float bounds[6]) {
    s->nodes_visited += 1;
    if (nd->dim == -1) { process_leaf(s, nd); return; }
    const orc_node *closer = s->t->nodes + nd->left, *further = s->t->nodes + nd->right;
    int close_b = 2 * nd->dim + 1, far_b = 2 * nd->dim;
    if (s->q[nd->dim] > nd->split) {
        const orc_node *tmp = closer; closer = further; further = tmp;
        int ti = close_b; close_b = far_b; far_b = ti;
    }
    {
        float cb[6]; memcpy(cb, bounds, sizeof cb);
        cb[close_b] = nd->split;
        float d = box_d2(s->q, cb, s->t->periodic, s->t->box);
        if (d < s->best[s->top].d) compute(s, closer, cb);
        else s->nodes_pruned += 1;
    }
    float fb[6]; memcpy(fb, bounds, sizeof fb);
    fb[far_b] = nd->split;
    float d = box_d2(s->q, fb, s->t->periodic, s->t->box);
    if (s->best[s->top].d < d) { s->nodes_pruned += 1; return; }
    compute(s, further, fb);
}

static int cmp_pair(const void *a, const void *b) {
    const orc_pair *l = (const orc_pair *)a, *r = (const orc_pair *)b;
    if (l->d < r->d) return -1;
    if (l->d > r->d) return 1;
    return (l->i > r->i) - (l->i < r->i);
}

static void finish_row(orc_pair *best, int k, int squared, float *out_d, uint32_t *out_i) {
    qsort(best, (size_t)k, sizeof(orc_pair), cmp_pair);   /* kdtree.cpp:149-151, canonical ties */
    for (int j = 0; j < k; ++j) {
        /* postprocess, kdtree.cpp:154-156; squared: the metric's value before it (kdtree.hpp:22-31,71-84) */
        out_d[j] = squared ? best[j].d : sqrtf(best[j].d);
        out_i[j] = best[j].i;
    }
}

typedef struct {
    const orc_tree *t;
    const float *q;
    uint64_t begin, end;
    int k, brute;
    float *out_d;
    uint32_t *out_i;
    uint64_t stats[3];
} orc_job;

/* find_nearest_naive, tests/test.cpp:14-37, under the total order (d2, index). */
static void brute_one(const orc_tree *t, const float q[3], int k, orc_pair *best) {
    for (int j = 0; j < k; ++j) { best[j].d = FLT_MAX; best[j].i = 0xFFFFFFFFu; }
    for (uint64_t p = 0; p < t->n_pad; ++p) {
        float pt[3] = {t->x[p], t->y[p], t->z[p]};
        orc_pair c = {point_d2(pt, q, t->periodic, t->box), t->idx[p]};
        if (!(c.d < FLT_MAX)) continue;              /* never better than an empty slot */
        if (cmp_pair(&c, &best[k - 1]) >= 0) continue;
        int j = k - 1;
        while (j > 0 && cmp_pair(&c, &best[j - 1]) < 0) { best[j] = best[j - 1]; --j; }
        best[j] = c;
    }
}

static void *run_job(void *arg) {
    orc_job *job = (orc_job *)arg;
    const orc_tree *t = job->t;
    orc_pair *best = (orc_pair *)malloc((size_t)job->k * sizeof(orc_pair));
    for (uint64_t i = job->begin; i < job->end; ++i) {
        float *od = job->out_d + i * (uint64_t)job->k;
        uint32_t *oi = job->out_i + i * (uint64_t)job->k;
        if (job->brute & 1) {
            brute_one(t, job->q + 3 * i, job->k, best);
            finish_row(best, job->k, job->brute & 2, od, oi);
            continue;
        }
        orc_search s;
        memset(&s, 0, sizeof s);
        s.t = t; s.k = job->k; s.best = best;
        memcpy(s.q, job->q + 3 * i, sizeof s.q);
        for (int j = 0; j < job->k; ++j) { best[j].d = FLT_MAX; best[j].i = 0xFFFFFFFFu; }
        s.top = 0;
        float bounds[6];
        for (int d = 0; d < 3; ++d) {                 /* initial_box kdtree.hpp:51-61,111-120 */
            bounds[2 * d] = t->periodic ? 0.0f : -FLT_MAX;
            bounds[2 * d + 1] = t->periodic ? t->box : FLT_MAX;
        }
        compute(&s, t->nodes, bounds);
        finish_row(best, job->k, job->brute & 2, od, oi);
        job->stats[0] += s.nodes_visited;
        job->stats[1] += s.nodes_pruned;
        job->stats[2] += s.points_visited;
    }
    free(best);
    return NULL;
}

/* Batched query (PyKDTree::query pybind.cpp:90-189): contiguous chunks of queries per worker like
 * thread_pool::parallelize_loop (thread_pool.hpp:147-183).  brute & 1 => exhaustive scan instead
 * of the tree walk; brute & 2 => rows hold squared distances (no postprocess).  stats (optional) receives the summed KDTreeQueryStatistics counters
 * (kdtree.hpp:124-131).  Returns 0 ok, 1 if k <= 0. */
int orc_tree_query(const orc_tree *t, const float *q_aos, uint64_t m, int k, int workers, int brute,
                   float *out_d, uint32_t *out_i, uint64_t *stats) {
    if (k <= 0) return 1;
    if (workers < 1) workers = 1;
    if ((uint64_t)workers > m) workers = m ? (int)m : 1;
    orc_job *jobs = (orc_job *)calloc((size_t)workers, sizeof(orc_job));
    pthread_t *th = (pthread_t *)calloc((size_t)workers, sizeof(pthread_t));
    uint64_t chunk = m / (uint64_t)workers;
    for (int w = 0; w < workers; ++w) {
        jobs[w].t = t; jobs[w].q = q_aos; jobs[w].k = k; jobs[w].brute = brute;
        jobs[w].out_d = out_d; jobs[w].out_i = out_i;
        jobs[w].begin = (uint64_t)w * chunk;
        jobs[w].end = w == workers - 1 ? m : (uint64_t)(w + 1) * chunk;
        if (workers > 1) pthread_create(&th[w], NULL, run_job, &jobs[w]);
        else run_job(&jobs[w]);
    }
    if (stats) stats[0] = stats[1] = stats[2] = 0;
    for (int w = 0; w < workers; ++w) {
        if (workers > 1) pthread_join(th[w], NULL);
        if (stats) for (int c = 0; c < 3; ++c) stats[c] += jobs[w].stats[c];
    }
    free(jobs); free(th);
    return 0;
}
