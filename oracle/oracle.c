/*
 * oracle.c — CPU restatement of agent-bom's exposure-graph traversals over a CSR.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (agent_bom_b200/)
 * may import, link or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do, and there only as
 * the checker / reported CPU baseline, never as the thing shipped.
 *
 * Every function below is a sequential restatement of one reference function
 * (paths relative to /root/reference/src/agent_bom/):
 *
 *   orc_impact_many        graph/container.py:230-279   UnifiedGraph.impact_of
 *   orc_bfs_many           graph/container.py:367-391   UnifiedGraph.bfs
 *   orc_reachable_many     graph/container.py:411-436   UnifiedGraph.reachable_from
 *   orc_shortest_path      graph/container.py:393-409   UnifiedGraph.shortest_path
 *   orc_traverse           graph/container.py:438-538   UnifiedGraph.traverse_subgraph
 *   orc_distances_many     graph/dependency_reach.py:169-198  _bfs_distances_along
 *   orc_derived_paths      api/routes/graph.py:686-786  _derived_attack_paths
 *                          api/routes/graph.py:488-503  _edge_relationships_for_hops
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks every function
 * against outputs of the unmodified reference (tests/golden/*.json.gz, made by
 * oracle/make_golden.py which imports /root/reference/src) including the
 * reference's own known-answer graphs.
 *
 * Graph format (shared with the device engine):
 *   fwd row u = graph.adjacency[u] in list order, rev row u =
 *   graph.reverse_adjacency[u] in list order (container.py:146-198): every
 *   original edge i contributes entry (row=src, nbr=dst, eid2=2i) to fwd and
 *   (row=dst, nbr=src, eid2=2i) to rev; a bidirectional edge additionally
 *   contributes its reversed copy (row=dst, nbr=src, eid2=2i+1) to fwd and
 *   (row=src, nbr=dst, eid2=2i+1) to rev.  meta byte = rel(5b) | traversable<<5
 *   | first_pair<<6 (unused by the oracle) | reversed_copy<<7.  node_type[u] = entity code, 255 for
 *   an id that appears only as an edge endpoint ("ghost": not in graph.nodes).
 *
 * Sources are independent, so the *_many functions run them on all host
 * threads (OpenMP) — this is also the CPU baseline bench.py reports.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define META_REL 0x1F
#define META_TRAV 0x20
#define META_FIRSTPAIR 0x40
#define META_REVCOPY 0x80
#define GHOST 255
#define NHIST 24

enum { REL_USES = 1, REL_DEPENDS_ON = 2, REL_PROVIDES_TOOL = 3, REL_EXPOSES_CRED = 4, REL_VULNERABLE_TO = 9 };
enum { ET_AGENT = 0, ET_SERVER = 1, ET_PACKAGE = 2, ET_VULN = 8, ET_MISCONF = 9, ET_USER = 13, ET_SERVICE_ACCOUNT = 17 };

typedef struct {
    int32_t n_nodes;
    int64_t n_entries;
    const uint32_t *fwd_off; const int32_t *fwd_nbr; const uint8_t *fwd_meta; const uint32_t *fwd_eid;
    const uint32_t *rev_off; const int32_t *rev_nbr; const uint8_t *rev_meta; const uint32_t *rev_eid;
    const uint8_t *node_type;
} orc_graph;

/* ragged result of a batch of traversals */
typedef struct {
    int64_t n_queries;
    int64_t *off;       /* [n_queries+1] into nodes/aux */
    int32_t *nodes;     /* discovery order */
    int32_t *aux;       /* parent position (bfs) or depth (others) */
    int64_t *eoff;      /* [n_queries+1] into edges (traverse only) */
    uint32_t *edges;    /* eid2 of every passing candidate, scan order, duplicates kept */
    uint32_t *hist;     /* [n_queries*24] (impact only) */
    int32_t *maxd;      /* [n_queries] max depth reached */
    int32_t *flags;     /* [n_queries] bit0 truncated, bit1 source-not-in-nodes */
    int64_t *edge_count;/* [n_queries] reference's edge_count budget counter */
    int64_t n_exp, m_scan, n_disc; /* algorithmic-byte counters, summed over queries */
} orc_result;

/* ---------- small growable buffers ---------- */
typedef struct { int32_t *p; int64_t n, cap; } vec32;
typedef struct { uint32_t *p; int64_t n, cap; } vecu32;
static void v32_push(vec32 *v, int32_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = (int32_t *)realloc(v->p, (size_t)v->cap * 4); }
    v->p[v->n++] = x;
}
static void vu32_push(vecu32 *v, uint32_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = (uint32_t *)realloc(v->p, (size_t)v->cap * 4); }
    v->p[v->n++] = x;
}

/* per-thread scratch: epoch-stamped visited array + output accumulation */
typedef struct {
    int32_t *stamp; int32_t epoch;
    vec32 nodes, aux; vecu32 edges;
    int64_t n_exp, m_scan, n_disc;
} scratch;

static void scratch_init(scratch *s, int32_t n) {
    memset(s, 0, sizeof(*s));
    s->stamp = (int32_t *)calloc((size_t)n + 1, 4);
    s->epoch = 0;
}
static void scratch_free(scratch *s) { free(s->stamp); free(s->nodes.p); free(s->aux.p); free(s->edges.p); }
static inline int seen(scratch *s, int32_t v) { return s->stamp[v] == s->epoch; }
static inline void mark(scratch *s, int32_t v) { s->stamp[v] = s->epoch; }
static inline void unmark(scratch *s, int32_t v) { s->stamp[v] = 0; }

static orc_result *result_new(int64_t nq, int want_hist, int want_edges) {
    orc_result *r = (orc_result *)calloc(1, sizeof(orc_result));
    r->n_queries = nq;
    r->off = (int64_t *)calloc((size_t)nq + 1, 8);
    r->maxd = (int32_t *)calloc((size_t)nq + 1, 4);
    r->flags = (int32_t *)calloc((size_t)nq + 1, 4);
    r->edge_count = (int64_t *)calloc((size_t)nq + 1, 8);
    if (want_hist) r->hist = (uint32_t *)calloc((size_t)nq * NHIST + 1, 4);
    if (want_edges) r->eoff = (int64_t *)calloc((size_t)nq + 1, 8);
    return r;
}

void orc_result_free(orc_result *r) {
    if (!r) return;
    free(r->off); free(r->nodes); free(r->aux); free(r->eoff); free(r->edges);
    free(r->hist); free(r->maxd); free(r->flags); free(r->edge_count); free(r);
}
/* accessors for ctypes */
int64_t orc_result_total(const orc_result *r) { return r->off[r->n_queries]; }
int64_t orc_result_total_edges(const orc_result *r) { return r->eoff ? r->eoff[r->n_queries] : 0; }
const int64_t *orc_result_off(const orc_result *r) { return r->off; }
const int32_t *orc_result_nodes(const orc_result *r) { return r->nodes; }
const int32_t *orc_result_aux(const orc_result *r) { return r->aux; }
const int64_t *orc_result_eoff(const orc_result *r) { return r->eoff; }
const uint32_t *orc_result_edges(const orc_result *r) { return r->edges; }
const uint32_t *orc_result_hist(const orc_result *r) { return r->hist; }
const int32_t *orc_result_maxd(const orc_result *r) { return r->maxd; }
const int32_t *orc_result_flags(const orc_result *r) { return r->flags; }
const int64_t *orc_result_edge_count(const orc_result *r) { return r->edge_count; }
void orc_result_counters(const orc_result *r, int64_t *out3) { out3[0] = r->n_exp; out3[1] = r->m_scan; out3[2] = r->n_disc; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Batch driver: runs `body(q)` for every query on all threads; each query's
 * output (appended to the thread's scratch vectors between `begin` and `end`)
 * is stitched into query order afterwards.
 */
typedef void (*query_fn)(const orc_graph *g, scratch *s, int64_t q, void *ctx, orc_result *r);

static void run_batch(const orc_graph *g, int64_t nq, query_fn fn, void *ctx, orc_result *r, int threads) {
    int nt = 1;
#ifdef _OPENMP
    nt = threads > 0 ? threads : omp_get_max_threads();
#endif
    (void)threads;
    scratch *S = (scratch *)calloc((size_t)nt, sizeof(scratch));
    /* where each query's slice lives: thread, start in thread vec, count */
    int32_t *q_thr = (int32_t *)malloc((size_t)(nq + 1) * 4);
    int64_t *q_start = (int64_t *)malloc((size_t)(nq + 1) * 8);
    int64_t *q_estart = (int64_t *)malloc((size_t)(nq + 1) * 8);
    int64_t *q_cnt = (int64_t *)calloc((size_t)nq + 1, 8);
    int64_t *q_ecnt = (int64_t *)calloc((size_t)nq + 1, 8);
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        scratch *s = &S[t];
        scratch_init(s, g->n_nodes);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (int64_t q = 0; q < nq; q++) {
            s->epoch++;
            int64_t n0 = s->nodes.n, e0 = s->edges.n;
            fn(g, s, q, ctx, r);
            q_thr[q] = t; q_start[q] = n0; q_cnt[q] = s->nodes.n - n0;
            q_estart[q] = e0; q_ecnt[q] = s->edges.n - e0;
        }
    }
    for (int64_t q = 0; q < nq; q++) r->off[q + 1] = r->off[q] + q_cnt[q];
    int64_t total = r->off[nq];
    r->nodes = (int32_t *)malloc((size_t)(total + 1) * 4);
    r->aux = (int32_t *)malloc((size_t)(total + 1) * 4);
    if (r->eoff) {
        for (int64_t q = 0; q < nq; q++) r->eoff[q + 1] = r->eoff[q] + q_ecnt[q];
        r->edges = (uint32_t *)malloc((size_t)(r->eoff[nq] + 1) * 4);
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
    for (int64_t q = 0; q < nq; q++) {
        scratch *s = &S[q_thr[q]];
        if (q_cnt[q]) {
            memcpy(r->nodes + r->off[q], s->nodes.p + q_start[q], (size_t)q_cnt[q] * 4);
            memcpy(r->aux + r->off[q], s->aux.p + q_start[q], (size_t)q_cnt[q] * 4);
        }
        if (r->eoff && q_ecnt[q]) memcpy(r->edges + r->eoff[q], s->edges.p + q_estart[q], (size_t)q_ecnt[q] * 4);
    }
    for (int t = 0; t < nt; t++) {
        r->n_exp += S[t].n_exp; r->m_scan += S[t].m_scan; r->n_disc += S[t].n_disc;
        scratch_free(&S[t]);
    }
    free(S); free(q_thr); free(q_start); free(q_estart); free(q_cnt); free(q_ecnt);
}

/* ------------------------------------------------------------------ */
/* impact_of — container.py:230-279                                    */
/* ------------------------------------------------------------------ */
typedef struct { const int32_t *sources; int32_t max_depth; } impact_ctx;

static void impact_one(const orc_graph *g, scratch *s, int64_t q, void *vctx, orc_result *r) {
    const impact_ctx *c = (const impact_ctx *)vctx;
    int32_t src = c->sources[q];
    /* :247-248  node_id not in self.nodes -> all-zero result */
    if (src < 0 || src >= g->n_nodes || g->node_type[src] == GHOST) { r->flags[q] = 2; return; }
    int64_t base = s->nodes.n;     /* queue == this query's slice of s->nodes; aux = depth */
    /* the source occupies a virtual queue slot; we keep it in a local so the
       emitted slice is exactly visited \ {source} in discovery order */
    mark(s, src);                  /* :251 visited = {node_id} */
    int32_t maxd = 0;
    int64_t head = -1;             /* -1 = the source itself, then indices into the slice */
    for (;;) {
        int32_t cur, depth;
        if (head < 0) { cur = src; depth = 0; }
        else { if (base + head >= s->nodes.n) break; cur = s->nodes.p[base + head]; depth = s->aux.p[base + head]; }
        head++;
        if (depth >= c->max_depth) continue;                 /* :257-258 */
        s->n_exp++;
        uint32_t a = g->rev_off[cur], b = g->rev_off[cur + 1];
        s->m_scan += (int64_t)(b - a);
        for (uint32_t p = a; p < b; p++) {                   /* :260 reverse_adjacency[current] */
            int32_t v = g->rev_nbr[p];                       /* edge.source */
            if (!seen(s, v)) {                               /* :261-264 */
                mark(s, v);
                v32_push(&s->nodes, v); v32_push(&s->aux, depth + 1);
                s->n_disc++;
                if (depth + 1 > maxd) maxd = depth + 1;
            }
        }
    }
    /* :266-271 histogram over ids present in nodes */
    uint32_t *h = r->hist + q * NHIST;
    for (int64_t i = base; i < s->nodes.n; i++) {
        uint8_t t = g->node_type[s->nodes.p[i]];
        if (t < NHIST) h[t]++;
    }
    r->maxd[q] = maxd;
}

orc_result *orc_impact_many(const orc_graph *g, const int32_t *sources, int64_t n, int32_t max_depth, int threads) {
    orc_result *r = result_new(n, 1, 0);
    impact_ctx c = { sources, max_depth };
    run_batch(g, n, impact_one, &c, r, threads);
    return r;
}

/* ------------------------------------------------------------------ */
/* bfs — container.py:367-391 (aux = parent position in the slice, -1 = source) */
/* ------------------------------------------------------------------ */
typedef struct { const int32_t *sources; int32_t max_depth; int32_t traversable_only; } bfs_ctx;

static void bfs_one(const orc_graph *g, scratch *s, int64_t q, void *vctx, orc_result *r) {
    const bfs_ctx *c = (const bfs_ctx *)vctx;
    int32_t src = c->sources[q];
    if (src < 0 || src >= g->n_nodes || g->node_type[src] == GHOST) { r->flags[q] = 2; return; }  /* :374-375 */
    /* Local queue of (node, parent, depth); depth = len(path)-1.  Nodes at
       depth max_depth+1 are marked visited and then dropped (:381-382) — kept
       in a private list here, emitted slice holds depth 1..max_depth only. */
    vec32 qn = {0}, qp = {0}, qd = {0};
    v32_push(&qn, src); v32_push(&qp, -1); v32_push(&qd, 0);
    mark(s, src);
    int64_t base = s->nodes.n;
    /* map from private queue index -> emitted position (for parent links) */
    vec32 emitted = {0};
    int32_t maxd = 0;
    for (int64_t head = 0; head < qn.n; head++) {
        int32_t cur = qn.p[head], depth = qd.p[head];
        if (depth > c->max_depth) { v32_push(&emitted, -2); continue; }          /* :381-382 */
        if (depth >= 1) {                                                        /* :383-384 */
            int32_t par = qp.p[head];
            int32_t parpos = par == 0 ? -1 : emitted.p[par];
            v32_push(&emitted, (int32_t)(s->nodes.n - base));
            v32_push(&s->nodes, cur); v32_push(&s->aux, parpos);
            if (depth > maxd) maxd = depth;
        } else v32_push(&emitted, -1);
        s->n_exp++;
        uint32_t a = g->fwd_off[cur], b = g->fwd_off[cur + 1];
        s->m_scan += (int64_t)(b - a);
        for (uint32_t p = a; p < b; p++) {                                       /* :385 */
            if (c->traversable_only && !(g->fwd_meta[p] & META_TRAV)) continue;  /* :386-387 */
            int32_t v = g->fwd_nbr[p];
            if (!seen(s, v)) {                                                   /* :388-390 */
                mark(s, v);
                v32_push(&qn, v); v32_push(&qp, (int32_t)head); v32_push(&qd, depth + 1);
                s->n_disc++;
            }
        }
    }
    r->maxd[q] = maxd;
    free(qn.p); free(qp.p); free(qd.p); free(emitted.p);
}

orc_result *orc_bfs_many(const orc_graph *g, const int32_t *sources, int64_t n, int32_t max_depth, int32_t traversable_only, int threads) {
    orc_result *r = result_new(n, 0, 0);
    bfs_ctx c = { sources, max_depth, traversable_only };
    run_batch(g, n, bfs_one, &c, r, threads);
    return r;
}

/* ------------------------------------------------------------------ */
/* reachable_from — container.py:411-436 (slice excludes the source; aux = depth) */
/* ------------------------------------------------------------------ */
static void reach_one(const orc_graph *g, scratch *s, int64_t q, void *vctx, orc_result *r) {
    const bfs_ctx *c = (const bfs_ctx *)vctx;
    int32_t src = c->sources[q];
    if (src < 0 || src >= g->n_nodes || g->node_type[src] == GHOST) { r->flags[q] = 2; return; }  /* :420-421 */
    int64_t base = s->nodes.n;
    mark(s, src);
    int32_t maxd = 0;
    int64_t head = -1;
    for (;;) {
        int32_t cur, depth;
        if (head < 0) { cur = src; depth = 0; }
        else { if (base + head >= s->nodes.n) break; cur = s->nodes.p[base + head]; depth = s->aux.p[base + head]; }
        head++;
        if (depth >= c->max_depth) continue;                                     /* :426-427 */
        s->n_exp++;
        uint32_t a = g->fwd_off[cur], b = g->fwd_off[cur + 1];
        s->m_scan += (int64_t)(b - a);
        for (uint32_t p = a; p < b; p++) {
            if (c->traversable_only && !(g->fwd_meta[p] & META_TRAV)) continue;  /* :429-430 */
            int32_t v = g->fwd_nbr[p];
            if (!seen(s, v)) {
                mark(s, v);
                v32_push(&s->nodes, v); v32_push(&s->aux, depth + 1);
                s->n_disc++;
                if (depth + 1 > maxd) maxd = depth + 1;
            }
        }
    }
    r->maxd[q] = maxd;
}

orc_result *orc_reachable_many(const orc_graph *g, const int32_t *sources, int64_t n, int32_t max_depth, int32_t traversable_only, int threads) {
    orc_result *r = result_new(n, 0, 0);
    bfs_ctx c = { sources, max_depth, traversable_only };
    run_batch(g, n, reach_one, &c, r, threads);
    return r;
}

/* ------------------------------------------------------------------ */
/* _bfs_distances_along — dependency_reach.py:169-198 (unbounded depth, rel mask) */
/* slice = every node reached except the start, aux = hop distance      */
/* ------------------------------------------------------------------ */
typedef struct { const int32_t *sources; uint32_t rel_mask; } dist_ctx;

static void dist_one(const orc_graph *g, scratch *s, int64_t q, void *vctx, orc_result *r) {
    const dist_ctx *c = (const dist_ctx *)vctx;
    int32_t src = c->sources[q];
    if (src < 0 || src >= g->n_nodes) { r->flags[q] = 2; return; }
    int64_t base = s->nodes.n;
    mark(s, src);                                                                /* :181 distances = {start: 0} */
    int32_t maxd = 0;
    int64_t head = -1;
    for (;;) {
        int32_t cur, depth;
        if (head < 0) { cur = src; depth = 0; }
        else { if (base + head >= s->nodes.n) break; cur = s->nodes.p[base + head]; depth = s->aux.p[base + head]; }
        head++;
        s->n_exp++;
        uint32_t a = g->fwd_off[cur], b = g->fwd_off[cur + 1];
        s->m_scan += (int64_t)(b - a);
        for (uint32_t p = a; p < b; p++) {                                       /* :186-197 */
            if (!((c->rel_mask >> (g->fwd_meta[p] & META_REL)) & 1u)) continue;
            int32_t v = g->fwd_nbr[p];
            if (seen(s, v)) continue;
            mark(s, v);
            v32_push(&s->nodes, v); v32_push(&s->aux, depth + 1);
            s->n_disc++;
            if (depth + 1 > maxd) maxd = depth + 1;
        }
    }
    r->maxd[q] = maxd;
}

orc_result *orc_distances_many(const orc_graph *g, const int32_t *sources, int64_t n, uint32_t rel_mask, int threads) {
    orc_result *r = result_new(n, 0, 0);
    dist_ctx c = { sources, rel_mask };
    run_batch(g, n, dist_one, &c, r, threads);
    return r;
}

/* ------------------------------------------------------------------ */
/* shortest_path — container.py:393-409.  Returns path length (#nodes) */
/* written to out_path (capacity cap), 0 = None.                        */
/* ------------------------------------------------------------------ */
int64_t orc_shortest_path(const orc_graph *g, int32_t src, int32_t dst, int32_t *out_path, int64_t cap) {
    if (src < 0 || dst < 0 || src >= g->n_nodes || dst >= g->n_nodes) return 0;
    if (g->node_type[src] == GHOST || g->node_type[dst] == GHOST) return 0;      /* :395-396 */
    if (src == dst) { if (cap >= 1) out_path[0] = src; return 1; }               /* :397-398 */
    int32_t *parent = (int32_t *)malloc((size_t)g->n_nodes * 4);
    uint8_t *vis = (uint8_t *)calloc((size_t)g->n_nodes, 1);
    vec32 qn = {0};
    v32_push(&qn, src); vis[src] = 1; parent[src] = -1;
    int64_t found = 0;
    for (int64_t head = 0; head < qn.n && !found; head++) {
        int32_t cur = qn.p[head];
        for (uint32_t p = g->fwd_off[cur]; p < g->fwd_off[cur + 1]; p++) {
            int32_t v = g->fwd_nbr[p];
            if (v == dst) {                                                      /* :404-405 target test precedes visited test */
                int64_t len = 2; for (int32_t u = cur; parent[u] >= 0; u = parent[u]) len++;
                if (len <= cap) {
                    out_path[len - 1] = dst;
                    int64_t i = len - 2; for (int32_t u = cur; u >= 0; u = parent[u]) out_path[i--] = u;
                }
                found = len; break;
            }
            if (!vis[v]) { vis[v] = 1; parent[v] = cur; v32_push(&qn, v); }       /* :406-408 */
        }
    }
    free(parent); free(vis); free(qn.p);
    return found;
}

/* ------------------------------------------------------------------ */
/* traverse_subgraph — container.py:438-538 (one query per call group) */
/* slice = queue in order (roots first, duplicates kept), aux = depth   */
/* edges = eid2 of every recorded candidate in scan order               */
/* ------------------------------------------------------------------ */
typedef struct {
    const int32_t *roots; const int64_t *root_off;   /* query q uses roots[root_off[q]..root_off[q+1]) */
    int32_t direction;      /* 1 forward, 2 reverse, 3 both */
    int32_t max_depth;
    int64_t max_nodes, max_edges;
    uint32_t rel_mask;      /* relationship_types ∧ static/dynamic folded by the caller */
    int32_t traversable_only, include_roots;
} trav_ctx;

static void trav_one(const orc_graph *g, scratch *s, int64_t q, void *vctx, orc_result *r) {
    const trav_ctx *c = (const trav_ctx *)vctx;
    int64_t base = s->nodes.n;
    int64_t visited_count = 0;
    vec32 newly = {0};   /* nodes marked this query, to undo the epoch trick for include_roots=0 re-marking */
    (void)newly;
    for (int64_t i = c->root_off[q]; i < c->root_off[q + 1]; i++) {              /* :465-472 */
        int32_t root = c->roots[i];
        if (root < 0 || root >= g->n_nodes || g->node_type[root] == GHOST) continue;
        v32_push(&s->nodes, root); v32_push(&s->aux, 0);
        if (c->include_roots && !seen(s, root)) { mark(s, root); visited_count++; }
    }
    int truncated = 0; int64_t edge_count = 0; int32_t maxd = 0;
    for (int64_t head = 0; base + head < s->nodes.n; head++) {
        int32_t cur = s->nodes.p[base + head], depth = s->aux.p[base + head];
        if (depth >= c->max_depth) continue;                                     /* :495-496 */
        s->n_exp++;
        int stop = 0;
        for (int pass = 0; pass < 2 && !stop; pass++) {                          /* :498-502 forward candidates then reverse */
            const uint32_t *off; const int32_t *nbr; const uint8_t *meta; const uint32_t *eid;
            if (pass == 0) { if (!(c->direction & 1)) continue; off = g->fwd_off; nbr = g->fwd_nbr; meta = g->fwd_meta; eid = g->fwd_eid; }
            else { if (!(c->direction & 2)) continue; off = g->rev_off; nbr = g->rev_nbr; meta = g->rev_meta; eid = g->rev_eid; }
            s->m_scan += (int64_t)(off[cur + 1] - off[cur]);
            for (uint32_t p = off[cur]; p < off[cur + 1]; p++) {
                uint8_t m = meta[p];
                if (!((c->rel_mask >> (m & META_REL)) & 1u)) continue;           /* :479-488 _edge_allowed */
                if (c->traversable_only && !(m & META_TRAV)) continue;
                edge_count++;                                                    /* :507 */
                if (c->max_edges >= 0 && edge_count > c->max_edges) { truncated = 1; stop = 1; break; }  /* :508-510 */
                vu32_push(&s->edges, eid[p]);                                    /* :512-513 */
                int32_t v = nbr[p];
                if (seen(s, v)) continue;                                        /* :515-516 */
                if (c->max_nodes >= 0 && visited_count >= c->max_nodes) { truncated = 1; continue; }  /* :517-519 */
                mark(s, v); visited_count++;
                v32_push(&s->nodes, v); v32_push(&s->aux, depth + 1);            /* :520-522 */
                s->n_disc++;
                if (depth + 1 > maxd) maxd = depth + 1;
            }
        }
        if (stop) break;                                                         /* :523-524 */
    }
    r->maxd[q] = maxd;
    r->flags[q] = truncated ? 1 : 0;
    r->edge_count[q] = edge_count;
}

orc_result *orc_traverse_many(const orc_graph *g, const int32_t *roots, const int64_t *root_off, int64_t n_queries,
                              int32_t direction, int32_t max_depth, int64_t max_nodes, int64_t max_edges,
                              uint32_t rel_mask, int32_t traversable_only, int32_t include_roots, int threads) {
    orc_result *r = result_new(n_queries, 0, 1);
    trav_ctx c = { roots, root_off, direction, max_depth, max_nodes, max_edges, rel_mask, traversable_only, include_roots };
    run_batch(g, n_queries, trav_one, &c, r, threads);
    return r;
}

/* ------------------------------------------------------------------ */
/* _derived_attack_paths — api/routes/graph.py:686-786                 */
/* Emits one row per (agent, server, vulnerable_source, finding) in the */
/* reference's emission order (before the final risk sort, which needs  */
/* the host's float risk and is applied by the caller).                 */
/* ------------------------------------------------------------------ */
typedef struct {
    int64_t n_paths;
    int32_t *hops;   /* [n_paths*4]: agent, server, vuln_source (-1 if == server), finding */
    int8_t *rels;    /* [n_paths*3]: relationship code per consecutive hop pair, -1 = pair has no edge (skipped) */
    int32_t *ncred;  /* un-deduplicated counts (graph.py:763-764) */
    int32_t *ntool;
} orc_paths;

void orc_paths_free(orc_paths *p) { if (!p) return; free(p->hops); free(p->rels); free(p->ncred); free(p->ntool); free(p); }
int64_t orc_paths_count(const orc_paths *p) { return p->n_paths; }
const int32_t *orc_paths_hops(const orc_paths *p) { return p->hops; }
const int8_t *orc_paths_rels(const orc_paths *p) { return p->rels; }
const int32_t *orc_paths_ncred(const orc_paths *p) { return p->ncred; }
const int32_t *orc_paths_ntool(const orc_paths *p) { return p->ntool; }

/* first relationship recorded for (a,b): by_pair.setdefault over graph.edges
   order incl. the (target,source) registration of bidirectional edges
   (graph.py:492-497) == first entry of fwd row a whose neighbour is b. */
static int first_rel(const orc_graph *g, int32_t a, int32_t b) {
    for (uint32_t p = g->fwd_off[a]; p < g->fwd_off[a + 1]; p++)
        if (g->fwd_nbr[p] == b) return g->fwd_meta[p] & META_REL;
    return -1;
}

static int cmp_rank(const void *x, const void *y, void *rk) {
    const int32_t *rank = (const int32_t *)rk;
    int32_t a = rank[*(const int32_t *)x], b = rank[*(const int32_t *)y];
    return (a > b) - (a < b);
}

/* rows of one finding, appended to the given vectors; returns the row count */
static int64_t paths_one(const orc_graph *g, int32_t f, const int32_t *node_rank, vec32 *hops, vec32 *rels, vec32 *nc, vec32 *nt,
                         vec32 *servers, vec32 *agents) {
    int64_t np = 0;
    if (f < 0 || f >= g->n_nodes) return 0;
    uint8_t ft = g->node_type[f];
    if (ft != ET_VULN && ft != ET_MISCONF) return 0;                             /* :708-710 */
    for (uint32_t p = g->rev_off[f]; p < g->rev_off[f + 1]; p++) {               /* :711 incoming[finding] = originals only */
        uint8_t m = g->rev_meta[p];
        if (m & META_REVCOPY) continue;
        if ((m & META_REL) != REL_VULNERABLE_TO) continue;                       /* :712-713 */
        int32_t vs = g->rev_nbr[p];
        if (g->node_type[vs] == GHOST) continue;                                 /* :714-716 */
        servers->n = 0;
        if (g->node_type[vs] == ET_SERVER) v32_push(servers, vs);                /* :719-720 */
        else {
            for (uint32_t p2 = g->rev_off[vs]; p2 < g->rev_off[vs + 1]; p2++) {  /* :722-726 */
                uint8_t m2 = g->rev_meta[p2];
                if (m2 & META_REVCOPY) continue;
                if ((m2 & META_REL) != REL_DEPENDS_ON) continue;
                int32_t sp = g->rev_nbr[p2];
                if (g->node_type[sp] == ET_SERVER) v32_push(servers, sp);
            }
        }
        for (int64_t si = 0; si < servers->n; si++) {
            int32_t srv = servers->p[si];
            agents->n = 0;
            for (uint32_t p3 = g->rev_off[srv]; p3 < g->rev_off[srv + 1]; p3++) {     /* :729-736 */
                uint8_t m3 = g->rev_meta[p3];
                if (m3 & META_REVCOPY) continue;
                if ((m3 & META_REL) != REL_USES) continue;
                int32_t a = g->rev_nbr[p3];
                uint8_t at = g->node_type[a];
                if (at == ET_AGENT || at == ET_USER || at == ET_SERVICE_ACCOUNT) v32_push(agents, a);
            }
            if (agents->n == 0) v32_push(agents, srv);                           /* :737-738 */
            int32_t ncred = 0, ntool = 0;
            for (uint32_t p4 = g->fwd_off[srv]; p4 < g->fwd_off[srv + 1]; p4++) {     /* :740-749 outgoing = originals only */
                uint8_t m4 = g->fwd_meta[p4];
                if (m4 & META_REVCOPY) continue;
                if (g->node_type[g->fwd_nbr[p4]] == GHOST) continue;
                if ((m4 & META_REL) == REL_EXPOSES_CRED) ncred++;
                else if ((m4 & META_REL) == REL_PROVIDES_TOOL) ntool++;
            }
            qsort_r(agents->p, (size_t)agents->n, 4, cmp_rank, (void *)node_rank);     /* :751 sorted(set(agent_ids)) */
            for (int64_t ai = 0; ai < agents->n; ai++) {
                if (ai && agents->p[ai] == agents->p[ai - 1]) continue;
                int32_t a = agents->p[ai];
                int32_t hp[4]; int nh = 0;
                hp[nh++] = a; hp[nh++] = srv; if (vs != srv) hp[nh++] = vs; hp[nh++] = f;     /* :752-755 */
                v32_push(hops, a); v32_push(hops, srv); v32_push(hops, vs != srv ? vs : -1); v32_push(hops, f);
                for (int k = 0; k < 3; k++) v32_push(rels, k + 1 < nh ? first_rel(g, hp[k], hp[k + 1]) : -2);
                v32_push(nc, ncred); v32_push(nt, ntool);
                np++;
            }
        }
    }
    return np;
}

/* node_rank[u] = position of u's id in the sorted list of all id strings
   (the reference sorts agent ids as strings, graph.py:751).  Findings are
   independent, so they are spread over the host threads and the per-thread
   row blocks are stitched back in finding order (the reference's emission order). */
orc_paths *orc_derived_paths(const orc_graph *g, const int32_t *findings, int64_t n_findings, const int32_t *node_rank, int threads) {
    int nt_ = 1;
#ifdef _OPENMP
    nt_ = threads > 0 ? threads : omp_get_max_threads();
#endif
    (void)threads;
    typedef struct { vec32 hops, rels, nc, nt, servers, agents; } tbuf;
    tbuf *T = (tbuf *)calloc((size_t)nt_, sizeof(tbuf));
    int32_t *f_thr = (int32_t *)malloc((size_t)(n_findings + 1) * 4);
    int64_t *f_start = (int64_t *)malloc((size_t)(n_findings + 1) * 8);
    int64_t *f_cnt = (int64_t *)calloc((size_t)n_findings + 1, 8);
#ifdef _OPENMP
#pragma omp parallel num_threads(nt_)
#endif
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        tbuf *b = &T[t];
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (int64_t fi = 0; fi < n_findings; fi++) {
            f_thr[fi] = t; f_start[fi] = b->nc.n;
            f_cnt[fi] = paths_one(g, findings[fi], node_rank, &b->hops, &b->rels, &b->nc, &b->nt, &b->servers, &b->agents);
        }
    }
    int64_t *off = (int64_t *)malloc((size_t)(n_findings + 1) * 8);
    off[0] = 0;
    for (int64_t fi = 0; fi < n_findings; fi++) off[fi + 1] = off[fi] + f_cnt[fi];
    int64_t np = off[n_findings];
    orc_paths *out = (orc_paths *)calloc(1, sizeof(orc_paths));
    out->n_paths = np;
    out->hops = (int32_t *)malloc((size_t)np * 16 + 16);
    out->rels = (int8_t *)malloc((size_t)np * 3 + 16);
    out->ncred = (int32_t *)malloc((size_t)np * 4 + 16);
    out->ntool = (int32_t *)malloc((size_t)np * 4 + 16);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt_)
#endif
    for (int64_t fi = 0; fi < n_findings; fi++) {
        if (!f_cnt[fi]) continue;
        tbuf *b = &T[f_thr[fi]];
        int64_t s = f_start[fi], c = f_cnt[fi], o = off[fi];
        memcpy(out->hops + o * 4, b->hops.p + s * 4, (size_t)c * 16);
        memcpy(out->ncred + o, b->nc.p + s, (size_t)c * 4);
        memcpy(out->ntool + o, b->nt.p + s, (size_t)c * 4);
        for (int64_t i = 0; i < c * 3; i++) out->rels[o * 3 + i] = (int8_t)b->rels.p[s * 3 + i];
    }
    for (int t = 0; t < nt_; t++) { free(T[t].hops.p); free(T[t].rels.p); free(T[t].nc.p); free(T[t].nt.p); free(T[t].servers.p); free(T[t].agents.p); }
    free(T); free(f_thr); free(f_start); free(f_cnt); free(off);
    return out;
}
