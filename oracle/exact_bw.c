/*
 * oracle/exact_bw.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Exact-predicate sequential Bowyer-Watson Delaunay triangulation of P u S
 * (S = the reference's finite super-simplex vertices), 3D and 2D.  This is the
 * authority for edge-set parity: in general position DT(P u S)
 * is unique, so any correct construction order yields the same simplices.
 *
 * What it restates from the reference (semantics, not data structures):
 *   - conflict region = connected set of simplices whose open circumsphere
 *     contains p            /root/reference/src/delaunay_tree.rs:33-75
 *   - new simplices = p joined to every facet between a killed simplex and a
 *     surviving neighbour   /root/reference/src/delaunay_tree.rs:77-123
 *   - neighbour repair (outer back-pointers + sibling links)
 *                           /root/reference/src/delaunay_tree.rs:147-169, :674-695
 *   - killed simplices are removed  /root/reference/src/delaunay_tree.rs:205-208
 * The reference's ghost simplices (radius 0, never in conflict, :467-502) are
 * the "outside" sentinel; here the outside is neighbour id -1.
 *
 * Insertion order is free (uniqueness), so the oracle uses a BRIO order
 * (random doubling rounds, Morton-sorted inside a round) and a visibility walk
 * from the last created simplex instead of the reference's kd-tree seed.
 *
 * Vertex ids: 0..M-1 = super vertices, M+i = input point i  (M = dim+1).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int vo_orient2d(const double *a, const double *b, const double *c);
int vo_orient3d(const double *a, const double *b, const double *c, const double *d);
int vo_incircle(const double *a, const double *b, const double *c, const double *d);
int vo_insphere(const double *a, const double *b, const double *c, const double *d, const double *e);

typedef struct {
    int dim, M;
    int n;          /* real points */
    double *x;      /* (n+M)*dim coordinates, supers first */
    int *tv, *tn;   /* M per simplex */
    uint8_t *alive;
    int cap, hi;    /* simplex slots allocated / high-water mark */
    int *freel;
    int nfree;
    int *mark;      /* stamp per simplex */
    int stamp;
    int last;       /* a live simplex to start the walk from */
    /* per-insertion scratch */
    int *cav, ncav, capcav;
    int *bnd, nbnd, capbnd; /* codes t*M+i */
    /* sibling hash */
    uint64_t *hkey;
    int *hval, *hstamp;
    int hsize, hstampv;
    /* stats */
    uint64_t created, killed, walk_steps, tests;
    int err; /* 0 ok, 1 duplicate/no-conflict, 2 outside super simplex */
    /* cached edges */
    uint32_t *edges;
    uint64_t nedges;
} bw_t;

static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static void grow_simplices(bw_t *b, int need) {
    if (need <= b->cap) return;
    int nc = b->cap * 2;
    if (nc < need) nc = need;
    b->tv = (int *)realloc(b->tv, sizeof(int) * (size_t)nc * b->M);
    b->tn = (int *)realloc(b->tn, sizeof(int) * (size_t)nc * b->M);
    b->alive = (uint8_t *)realloc(b->alive, (size_t)nc);
    b->mark = (int *)realloc(b->mark, sizeof(int) * (size_t)nc);
    memset(b->mark + b->cap, 0, sizeof(int) * (size_t)(nc - b->cap));
    b->freel = (int *)realloc(b->freel, sizeof(int) * (size_t)nc);
    b->cap = nc;
}

static int new_simplex(bw_t *b) {
    if (b->nfree > 0) return b->freel[--b->nfree];
    grow_simplices(b, b->hi + 1);
    return b->hi++;
}

static inline const double *P(const bw_t *b, int v) { return b->x + (size_t)v * b->dim; }

/* sign of orientation of simplex t with slot i replaced by point q */
static int orient_repl(const bw_t *b, int t, int i, const double *q) {
    const int *v = b->tv + (size_t)t * b->M;
    const double *p[4];
    for (int k = 0; k < b->M; k++) p[k] = (k == i) ? q : P(b, v[k]);
    if (b->dim == 3) return vo_orient3d(p[0], p[1], p[2], p[3]);
    return vo_orient2d(p[0], p[1], p[2]);
}

static int in_conflict(bw_t *b, int t, const double *q) {
    const int *v = b->tv + (size_t)t * b->M;
    b->tests++;
    if (b->dim == 3) return vo_insphere(P(b, v[0]), P(b, v[1]), P(b, v[2]), P(b, v[3]), q) > 0;
    return vo_incircle(P(b, v[0]), P(b, v[1]), P(b, v[2]), q) > 0;
}

static void push_cav(bw_t *b, int t) {
    if (b->ncav == b->capcav) {
        b->capcav *= 2;
        b->cav = (int *)realloc(b->cav, sizeof(int) * (size_t)b->capcav);
    }
    b->cav[b->ncav++] = t;
}
static void push_bnd(bw_t *b, int code) {
    if (b->nbnd == b->capbnd) {
        b->capbnd *= 2;
        b->bnd = (int *)realloc(b->bnd, sizeof(int) * (size_t)b->capbnd);
    }
    b->bnd[b->nbnd++] = code;
}

/* sibling matching: key = the (dim-1)-face of the new simplex that contains p,
 * identified by its other dim-1 vertices (an edge in 3D, a vertex in 2D). */
static void hash_reset(bw_t *b, int need) {
    int want = 64;
    while (want < need * 4) want *= 2;
    if (want > b->hsize) {
        b->hkey = (uint64_t *)realloc(b->hkey, sizeof(uint64_t) * (size_t)want);
        b->hval = (int *)realloc(b->hval, sizeof(int) * (size_t)want);
        b->hstamp = (int *)realloc(b->hstamp, sizeof(int) * (size_t)want);
        memset(b->hstamp, 0, sizeof(int) * (size_t)want);
        b->hsize = want;
        b->hstampv = 0;
    }
    b->hstampv++;
}
/* returns previous value for key (and removes nothing), or -1 after inserting */
static int hash_match(bw_t *b, uint64_t key, int val) {
    uint64_t h = mix64(key) & (uint64_t)(b->hsize - 1);
    for (;;) {
        if (b->hstamp[h] != b->hstampv) {
            b->hstamp[h] = b->hstampv;
            b->hkey[h] = key;
            b->hval[h] = val;
            return -1;
        }
        if (b->hkey[h] == key) return b->hval[h];
        h = (h + 1) & (uint64_t)(b->hsize - 1);
    }
}

static int bw_insert(bw_t *b, int vid) {
    const int M = b->M;
    const double *q = P(b, vid);
    /* 1. visibility walk to the simplex containing q */
    int t = b->last;
    int prev = -1;
    unsigned rot = (unsigned)vid;
    for (;;) {
        int moved = 0;
        for (int k = 0; k < M; k++) {
            int i = (int)((rot + (unsigned)k) % (unsigned)M);
            int nb = b->tn[(size_t)t * M + i];
            if (nb == prev && prev >= 0) continue;
            if (orient_repl(b, t, i, q) < 0) {
                if (nb < 0) { b->err = 2; return -1; }
                prev = t;
                t = nb;
                moved = 1;
                b->walk_steps++;
                rot = rot * 1664525u + 1013904223u;
                break;
            }
        }
        if (!moved) break;
    }
    /* 2. conflict region (flood) */
    if (!in_conflict(b, t, q)) { b->err = 1; return -1; } /* duplicate point */
    b->stamp++;
    b->ncav = 0;
    b->nbnd = 0;
    push_cav(b, t);
    b->mark[t] = b->stamp;
    for (int head = 0; head < b->ncav; head++) {
        int c = b->cav[head];
        for (int i = 0; i < M; i++) {
            int nb = b->tn[(size_t)c * M + i];
            if (nb >= 0 && b->mark[nb] == b->stamp) continue;
            if (nb >= 0 && in_conflict(b, nb, q)) {
                b->mark[nb] = b->stamp;
                push_cav(b, nb);
            } else {
                push_bnd(b, c * M + i);
            }
        }
    }
    /* 3. retriangulate: one new simplex per boundary facet */
    hash_reset(b, b->nbnd * (M - 1));
    int firstnew = -1;
    for (int j = 0; j < b->nbnd; j++) {
        int c = b->bnd[j] / M, i = b->bnd[j] % M;
        int nt = new_simplex(b);
        if (firstnew < 0) firstnew = nt;
        int *v = b->tv + (size_t)nt * M, *nn = b->tn + (size_t)nt * M;
        const int *cv = b->tv + (size_t)c * M;
        for (int k = 0; k < M; k++) { v[k] = cv[k]; nn[k] = -1; }
        v[i] = vid;
        b->alive[nt] = 1;
        b->mark[nt] = 0;
        int outer = b->tn[(size_t)c * M + i];
        nn[i] = outer;
        if (outer >= 0) {
            int *on = b->tn + (size_t)outer * M;
            for (int k = 0; k < M; k++)
                if (on[k] == c) on[k] = nt;
        }
        /* faces containing p: opposite slot k != i */
        for (int k = 0; k < M; k++) {
            if (k == i) continue;
            uint64_t key;
            if (M == 4) {
                int e[2], ne = 0;
                for (int m = 0; m < 4; m++)
                    if (m != i && m != k) e[ne++] = v[m];
                int lo = e[0] < e[1] ? e[0] : e[1], hi2 = e[0] < e[1] ? e[1] : e[0];
                key = ((uint64_t)(uint32_t)lo << 32) | (uint32_t)hi2;
            } else {
                int m = 3 - i - k;
                key = (uint64_t)(uint32_t)v[m];
            }
            int other = hash_match(b, key, nt * M + k);
            if (other >= 0) {
                int ot = other / M, ok = other % M;
                nn[k] = ot;
                b->tn[(size_t)ot * M + ok] = nt;
            }
        }
    }
    /* 4. remove killed */
    for (int j = 0; j < b->ncav; j++) {
        int c = b->cav[j];
        b->alive[c] = 0;
        b->freel[b->nfree++] = c;
    }
    b->created += (uint64_t)b->nbnd;
    b->killed += (uint64_t)b->ncav;
    b->last = firstnew;
    return 0;
}

typedef struct { uint64_t key; int idx; } okey_t;
static int cmp_okey(const void *a, const void *b) {
    const okey_t *x = (const okey_t *)a, *y = (const okey_t *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}
static uint64_t spread3(uint64_t v) { /* 21 bits -> every third bit */
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}
static uint64_t spread2(uint64_t v) { /* 28 bits -> every second bit */
    v &= 0xfffffff;
    v = (v | v << 16) & 0x0000ffff0000ffffULL;
    v = (v | v << 8) & 0x00ff00ff00ff00ffULL;
    v = (v | v << 4) & 0x0f0f0f0f0f0f0f0fULL;
    v = (v | v << 2) & 0x3333333333333333ULL;
    v = (v | v << 1) & 0x5555555555555555ULL;
    return v;
}

/*
 * super: M*dim doubles = the reference's super-simplex vertex coordinates
 * (computed by vo_bootstrap in bootstrap.c, bit-identical to
 * /root/reference/src/delaunay_tree.rs:392-406 / :547-558).
 */
bw_t *vo_bw_create(int dim, const double *pts, int n, const double *super) {
    bw_t *b = (bw_t *)calloc(1, sizeof(bw_t));
    b->dim = dim;
    b->M = dim + 1;
    const int M = b->M;
    b->n = n;
    b->x = (double *)malloc(sizeof(double) * (size_t)(n + M) * dim);
    memcpy(b->x, super, sizeof(double) * M * dim);
    memcpy(b->x + M * dim, pts, sizeof(double) * (size_t)n * dim);
    b->capcav = 64;
    b->cav = (int *)malloc(sizeof(int) * 64);
    b->capbnd = 64;
    b->bnd = (int *)malloc(sizeof(int) * 64);
    b->cap = 0;
    grow_simplices(b, (dim == 3 ? 8 : 3) * (n + 16));
    int t0 = new_simplex(b);
    for (int k = 0; k < M; k++) {
        b->tv[t0 * M + k] = k;
        b->tn[t0 * M + k] = -1;
    }
    /* make the root simplex positively oriented */
    int s = dim == 3 ? vo_orient3d(P(b, 0), P(b, 1), P(b, 2), P(b, 3)) : vo_orient2d(P(b, 0), P(b, 1), P(b, 2));
    if (s < 0) {
        b->tv[t0 * M + 0] = 1;
        b->tv[t0 * M + 1] = 0;
    }
    b->alive[t0] = 1;
    b->last = t0;

    /* BRIO order */
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < dim; k++) {
            double c = pts[(size_t)i * dim + k];
            if (c < lo[k]) lo[k] = c;
            if (c > hi[k]) hi[k] = c;
        }
    okey_t *ord = (okey_t *)malloc(sizeof(okey_t) * (size_t)(n > 0 ? n : 1));
    const int bits = dim == 3 ? 19 : 28;
    for (int i = 0; i < n; i++) {
        uint64_t r = mix64(0xD1B54A32D192ED03ULL ^ (uint64_t)i) % (uint64_t)n; /* pseudo rank */
        int stage = 0;
        while ((r >> stage) >= 256) stage++;
        uint64_t code = 0;
        for (int k = 0; k < dim; k++) {
            double ext = hi[k] - lo[k];
            double u = ext > 0 ? (pts[(size_t)i * dim + k] - lo[k]) / ext : 0.0;
            uint64_t q = (uint64_t)(u * (double)((1u << bits) - 1));
            code |= (dim == 3 ? spread3(q) : spread2(q)) << k;
        }
        ord[i].key = ((uint64_t)stage << 58) | code;
        ord[i].idx = i;
    }
    qsort(ord, (size_t)n, sizeof(okey_t), cmp_okey);
    for (int i = 0; i < n; i++) {
        if (bw_insert(b, M + ord[i].idx) != 0) break;
    }
    free(ord);
    return b;
}

void vo_bw_destroy(bw_t *b) {
    if (!b) return;
    free(b->x); free(b->tv); free(b->tn); free(b->alive); free(b->freel); free(b->mark);
    free(b->cav); free(b->bnd); free(b->hkey); free(b->hval); free(b->hstamp); free(b->edges);
    free(b);
}

int vo_bw_error(const bw_t *b) { return b->err; }

/* stats: [live simplices, created, killed, walk steps, in-sphere tests] */
void vo_bw_stats(const bw_t *b, uint64_t *out) {
    uint64_t live = 0;
    for (int t = 0; t < b->hi; t++) live += b->alive[t];
    out[0] = live;
    out[1] = b->created;
    out[2] = b->killed;
    out[3] = b->walk_steps;
    out[4] = b->tests;
}

/* live simplices as vertex-id tuples (ids: 0..M-1 super, M+i input point i) */
uint64_t vo_bw_simplices(const bw_t *b, int *out, uint64_t cap) {
    uint64_t m = 0;
    for (int t = 0; t < b->hi; t++) {
        if (!b->alive[t]) continue;
        if (out && m < cap) memcpy(out + m * b->M, b->tv + (size_t)t * b->M, sizeof(int) * b->M);
        m++;
    }
    return m;
}

/* structural + Delaunay validation: adjacency symmetric, positive orientation,
 * locally Delaunay across every interior facet (=> globally Delaunay). */
int vo_bw_validate(bw_t *b) {
    const int M = b->M;
    for (int t = 0; t < b->hi; t++) {
        if (!b->alive[t]) continue;
        const int *v = b->tv + (size_t)t * M;
        int s = b->dim == 3 ? vo_orient3d(P(b, v[0]), P(b, v[1]), P(b, v[2]), P(b, v[3]))
                            : vo_orient2d(P(b, v[0]), P(b, v[1]), P(b, v[2]));
        if (s <= 0) return 1;
        for (int i = 0; i < M; i++) {
            int nb = b->tn[(size_t)t * M + i];
            if (nb < 0) continue;
            if (!b->alive[nb]) return 2;
            const int *nv = b->tv + (size_t)nb * M;
            int back = -1;
            for (int k = 0; k < M; k++)
                if (b->tn[(size_t)nb * M + k] == t) back = k;
            if (back < 0) return 3;
            /* shared facet: all of nb's vertices except nv[back] are in t */
            for (int k = 0; k < M; k++) {
                if (k == back) continue;
                int found = 0;
                for (int m = 0; m < M; m++) found |= (v[m] == nv[k] && m != i);
                if (!found) return 4;
            }
            if (in_conflict(b, t, P(b, nv[back]))) return 5;
        }
    }
    return 0;
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/*
 * Canonical Delaunay-graph edge list: {lo,hi} input
 * indices of every pair of real vertices that share a live simplex, lo<hi,
 * sorted lexicographically, unique, little-endian u32 pairs.
 */
uint64_t vo_bw_edges(bw_t *b, uint32_t *out, uint64_t cap) {
    const int M = b->M;
    if (!b->edges) {
        uint64_t *deg = (uint64_t *)calloc((size_t)b->n + 1, sizeof(uint64_t));
        for (int t = 0; t < b->hi; t++) {
            if (!b->alive[t]) continue;
            const int *v = b->tv + (size_t)t * M;
            for (int i = 0; i < M; i++)
                for (int j = i + 1; j < M; j++) {
                    if (v[i] < M || v[j] < M) continue;
                    int lo = (v[i] < v[j] ? v[i] : v[j]) - M;
                    deg[lo + 1]++;
                }
        }
        for (int i = 0; i < b->n; i++) deg[i + 1] += deg[i];
        uint64_t tot = deg[b->n];
        uint32_t *hi = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(tot ? tot : 1));
        uint64_t *fill = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)b->n + 1));
        memcpy(fill, deg, sizeof(uint64_t) * ((size_t)b->n + 1));
        for (int t = 0; t < b->hi; t++) {
            if (!b->alive[t]) continue;
            const int *v = b->tv + (size_t)t * M;
            for (int i = 0; i < M; i++)
                for (int j = i + 1; j < M; j++) {
                    if (v[i] < M || v[j] < M) continue;
                    int lo = (v[i] < v[j] ? v[i] : v[j]) - M;
                    int h2 = (v[i] < v[j] ? v[j] : v[i]) - M;
                    hi[fill[lo]++] = (uint32_t)h2;
                }
        }
        free(fill);
        uint64_t m = 0;
        b->edges = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (size_t)(tot ? tot : 1));
        for (int i = 0; i < b->n; i++) {
            uint64_t s = deg[i], e = deg[i + 1];
            qsort(hi + s, (size_t)(e - s), sizeof(uint32_t), cmp_u32);
            for (uint64_t k = s; k < e; k++) {
                if (k > s && hi[k] == hi[k - 1]) continue;
                b->edges[2 * m] = (uint32_t)i;
                b->edges[2 * m + 1] = hi[k];
                m++;
            }
        }
        b->nedges = m;
        b->edges = (uint32_t *)realloc(b->edges, sizeof(uint32_t) * 2 * (size_t)(m ? m : 1));
        free(hi);
        free(deg);
    }
    if (out) {
        uint64_t k = b->nedges < cap ? b->nedges : cap;
        memcpy(out, b->edges, sizeof(uint32_t) * 2 * (size_t)k);
    }
    return b->nedges;
}
