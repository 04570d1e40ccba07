/* TEST INFRASTRUCTURE — NOT part of the product path.
 *
 * CPU restatement (plain C, FP64, no FMA contraction) of the reference's 2-D Barnes-Hut
 * simulation path, `implementation/project.cu` of DavidSevic/gpu-nbody-simulation.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or as the timed CPU baseline.  The product
 * (gpu_nbody_simulation_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit against raw FP64
 * dumps produced by the reference's own functions (oracle/ref_harness.cu compiled against the
 * unmodified project.cu) — node table, forces, accelerations, velocities, positions — on the
 * reference's shipped initial conditions and on seeded synthetic inputs; the dumps are
 * committed under tests/golden/ with the script that made them (tests/golden/make_golden.py).
 * The reference has no tests or golden vectors of its own.
 *
 * Every function cites the reference lines it restates.  Parameters that are source-level
 * constants in the reference (G, DELTA_T, THETA, QUADTREE_MAX_DEPTH, the 1e-15 offsets, the
 * 10 % padding) are runtime fields here, with the reference's values as defaults.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Node record: same 12-double layout as the reference's Quadrant (project.cu:46-58). */
enum { CH0 = 0, CH1, CH2, CH3, COMX, COMY, MASS, XMIN, XMAX, YMIN, YMAX, PIDX, QSZ };

typedef struct {
    double G;            /* project.cu:27  6.67e-11 */
    double dt;           /* project.cu:29  1.0 */
    double theta;        /* project.cu:60  0.5 */
    double dist_eps;     /* project.cu:634 1e-15, added to the distance (not to d^2) */
    double mass_eps;     /* project.cu:617 1e-15, nodes with mass <= this are skipped */
    double pad_frac;     /* project.cu:558 0.1 */
    double pad_fallback; /* project.cu:564 1e-6 */
    int32_t max_depth;   /* project.cu:61  10 (root is depth 1) */
    int32_t _pad;
} bho_params;

typedef struct {
    double* nodes; /* size * QSZ */
    int64_t size, cap;
    int64_t max_size; /* project.cu:62 */
    int32_t max_depth;
} bho_tree;

void bho_default_params(bho_params* p) {
    p->G = 6.67e-11; p->dt = 1.0; p->theta = 5e-1; p->dist_eps = 1e-15; p->mass_eps = 1e-15;
    p->pad_frac = 0.1; p->pad_fallback = 1e-6; p->max_depth = 10; p->_pad = 0;
}

/* project.cu:536-573 ComputeRootBounds.  std::min(a,b) is (b<a)?b:a, std::max(a,b) is (a<b)?b:a. */
void bho_root_bounds(const double* pos, int64_t n, const bho_params* p, double out[4]) {
    double xMin = INFINITY, xMax = -INFINITY, yMin = INFINITY, yMax = -INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        double x = pos[2 * i], y = pos[2 * i + 1];
        xMin = (x < xMin) ? x : xMin;
        xMax = (xMax < x) ? x : xMax;
        yMin = (y < yMin) ? y : yMin;
        yMax = (yMax < y) ? y : yMax;
    }
    double dx = xMax - xMin, dy = yMax - yMin;
    double maxDim = (dx < dy) ? dy : dx;
    double pad = p->pad_frac * maxDim;
    if (maxDim == 0.0) pad = p->pad_fallback;
    out[0] = xMin - pad; out[1] = xMax + pad; out[2] = yMin - pad; out[3] = yMax + pad;
}

static int64_t push_node(bho_tree* t, double xmin, double xmax, double ymin, double ymax) {
    if (t->size == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->nodes = (double*)realloc(t->nodes, (size_t)t->cap * QSZ * sizeof(double));
    }
    double* q = t->nodes + t->size * QSZ;
    q[CH0] = q[CH1] = q[CH2] = q[CH3] = -1; q[COMX] = q[COMY] = q[MASS] = 0.0;
    q[XMIN] = xmin; q[XMAX] = xmax; q[YMIN] = ymin; q[YMAX] = ymax; q[PIDX] = -1;
    return t->size++;
}

/* project.cu:348-356 DetermineChild (the four explicit comparisons, so NaNs go to child 3). */
static int determine_child(double x, double y, const double* node) {
    double mid_x = (node[XMIN] + node[XMAX]) / 2;
    double mid_y = (node[YMIN] + node[YMAX]) / 2;
    if (x < mid_x && y < mid_y) return 0;
    if (x >= mid_x && y < mid_y) return 1;
    if (x < mid_x && y >= mid_y) return 2;
    return 3;
}

/* project.cu:358-453 QuadInsert. */
static void quad_insert(bho_tree* t, int64_t particle, int64_t node_index, const double* pos,
                        const double* mass, int depth) {
    if (depth >= t->max_depth) { /* :360-382 cap level: running weighted average */
        double* node = t->nodes + node_index * QSZ;
        double m = mass[particle];
        double em = node[MASS], ex = node[COMX], ey = node[COMY];
        node[COMX] = (em * ex + m * pos[2 * particle]) / (em + m);
        node[COMY] = (em * ey + m * pos[2 * particle + 1]) / (em + m);
        node[MASS] += m;
        if (em == 0) node[PIDX] = (double)(-1 * particle - 2);
        else node[PIDX] = -1;
        return;
    }
    if (node_index >= t->size) return; /* :385-388 */
    double node[QSZ]; /* :390 works on a COPY and writes it back */
    memcpy(node, t->nodes + node_index * QSZ, sizeof node);
    double px = pos[2 * particle], py = pos[2 * particle + 1], m = mass[particle];
    int empty_leaf = node[CH0] == -1 && node[CH1] == -1 && node[CH2] == -1 && node[CH3] == -1 &&
                     node[MASS] == 0.0;
    if (empty_leaf) { /* :398-406 */
        node[COMX] = px; node[COMY] = py; node[MASS] = m; node[PIDX] = (double)particle;
        memcpy(t->nodes + node_index * QSZ, node, sizeof node);
        return;
    }
    if (node[MASS] > 0.0 && node[PIDX] > -1) { /* :408-448 split: all four children appended */
        for (int i = 0; i < 4; ++i) {
            if (t->size >= t->max_size) return; /* :411-414 */
            double mid_x = (node[XMIN] + node[XMAX]) / 2.0;
            double mid_y = (node[YMIN] + node[YMAX]) / 2.0;
            int64_t c;
            if (i == 0) c = push_node(t, node[XMIN], mid_x, node[YMIN], mid_y);
            else if (i == 1) c = push_node(t, mid_x, node[XMAX], node[YMIN], mid_y);
            else if (i == 2) c = push_node(t, node[XMIN], mid_x, mid_y, node[YMAX]);
            else c = push_node(t, mid_x, node[XMAX], mid_y, node[YMAX]);
            node[CH0 + i] = (double)c;
        }
        double ex = node[COMX], ey = node[COMY];
        int64_t existing = (int64_t)(int)node[PIDX];
        node[COMX] = 0.0; node[COMY] = 0.0; node[MASS] = 0.0; node[PIDX] = -1;
        memcpy(t->nodes + node_index * QSZ, node, sizeof node);
        int ec = determine_child(ex, ey, node);
        quad_insert(t, existing, (int64_t)node[CH0 + ec], pos, mass, depth + 1);
    }
    int c = determine_child(px, py, node); /* :451-452 */
    quad_insert(t, particle, (int64_t)node[CH0 + c], pos, mass, depth + 1);
}

/* project.cu:473-502 ComputeMass (post-order, children 0..3, sums start from 0.0). */
static void compute_mass(bho_tree* t, int64_t node_index, double* m_out, double* cx_out, double* cy_out) {
    double* node = t->nodes + node_index * QSZ;
    if (node[CH0] == -1) { *m_out = node[MASS]; *cx_out = node[COMX]; *cy_out = node[COMY]; return; }
    double total = 0.0, cx = 0.0, cy = 0.0;
    for (int i = 0; i < 4; ++i) {
        if (node[CH0 + i] != -1) {
            double cm, ccx, ccy;
            compute_mass(t, (int64_t)node[CH0 + i], &cm, &ccx, &ccy);
            node = t->nodes + node_index * QSZ;
            total += cm; cx += cm * ccx; cy += cm * ccy;
        }
    }
    if (total > 0.0) { cx /= total; cy /= total; }
    node[MASS] = total; node[COMX] = cx; node[COMY] = cy;
    *m_out = total; *cx_out = cx; *cy_out = cy;
}

/* project.cu:575-591 buildTree. */
bho_tree* bho_build_tree(const double* pos, const double* mass, int64_t n, const bho_params* p) {
    bho_tree* t = (bho_tree*)calloc(1, sizeof *t);
    t->max_depth = p->max_depth;
    t->max_size = (int64_t)(int)((pow(4, p->max_depth) - 1) / 3); /* :62 */
    double b[4];
    bho_root_bounds(pos, n, p, b);
    push_node(t, b[0], b[1], b[2], b[3]); /* :343-346 InitializeRoot */
    for (int64_t i = 0; i < n; ++i) quad_insert(t, i, 0, pos, mass, 1);
    double m, cx, cy;
    compute_mass(t, 0, &m, &cx, &cy);
    return t;
}

int64_t bho_tree_size(const bho_tree* t) { return t->size; }
const double* bho_tree_nodes(const bho_tree* t) { return t->nodes; }
void bho_tree_free(bho_tree* t) { if (t) { free(t->nodes); free(t); } }

/* counters: [0] node visits (pops), [1] accepted interactions (executions of :651-658),
 * [2] opens, [3] zero-mass skips, [4] self skips, [5] max stack depth. */
/* project.cu:593-675 computeForces, for bodies i0, i0+stride, ... < i1. */
void bho_compute_forces(const bho_tree* t, const double* pos, const double* mass, int64_t n,
                        const bho_params* p, int64_t i0, int64_t i1, int64_t stride, int nthreads,
                        double* forces, int64_t* counters) {
    (void)n;
    int64_t c_vis = 0, c_int = 0, c_open = 0, c_zero = 0, c_self = 0, c_stack = 0;
    const double* nodes = t->nodes;
    const double G = p->G, theta = p->theta, deps = p->dist_eps, meps = p->mass_eps;
    int64_t nb = (i1 - i0 + stride - 1) / stride;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) \
    reduction(+ : c_vis, c_int, c_open, c_zero, c_self) reduction(max : c_stack)
#endif
    for (int64_t k = 0; k < nb; ++k) {
        int64_t i = i0 + k * stride;
        int64_t stack[4 * 64 + 8];
        int top = 0;
        double sx = 0.0, sy = 0.0;
        double pix = pos[2 * i], piy = pos[2 * i + 1];
        stack[top++] = 0;
        while (top > 0) {
            if (top > c_stack) c_stack = top;
            int64_t ni = stack[--top];
            const double* node = nodes + ni * QSZ;
            ++c_vis;
            double nodeMass = node[MASS];
            if (nodeMass <= meps) { ++c_zero; continue; }
            int occ = (int)node[PIDX];
            int leaf = node[CH0] == -1 && node[CH1] == -1 && node[CH2] == -1 && node[CH3] == -1;
            double dx = node[COMX] - pix, dy = node[COMY] - piy;
            double d2 = dx * dx + dy * dy;
            double d = sqrt(d2) + deps;
            double w = node[XMAX] - node[XMIN], h = node[YMAX] - node[YMIN];
            double size = (w > h) ? w : h;
            if (leaf || (size / d < theta)) {
                if (leaf && (occ == i || (occ + 2) == -i)) { ++c_self; continue; }
                double fm = (G * mass[i] * nodeMass) / d2;
                double nx = dx / d, ny = dy / d;
                sx += fm * nx; sy += fm * ny;
                ++c_int;
            } else {
                ++c_open;
                for (int c = 0; c < 4; ++c) {
                    int64_t ch = (int64_t)(int)node[CH0 + c];
                    if (ch != -1) stack[top++] = ch;
                }
            }
        }
        forces[2 * i] = sx; forces[2 * i + 1] = sy;
    }
    if (counters) {
        counters[0] = c_vis; counters[1] = c_int; counters[2] = c_open; counters[3] = c_zero;
        counters[4] = c_self; counters[5] = c_stack;
    }
}

/* ---------------------------------------------------------------------------------------------
 * EXTENSION — NOT a restatement of the reference, hence not pinned by it.
 *
 * "Exact leaves": the traversal of bho_compute_forces above, with ONE change: a multi-body leaf at
 * the depth cap (PARTICLE_INDEX == -1, the nodes project.cu:360-382 fills with a running average)
 * is not applied as one monopole at its centre of mass — which in the reference includes the body
 * itself when it sits in that leaf — but as the sum over the leaf's bodies
 * j != i of the SAME pair expression, G m_i m_j / d2 * (dx, dy) / (d + eps).  Members are visited in
 * ascending body index.  Every other node (internal nodes, single-body leaves, the self test) is
 * handled exactly as above, so with every cap-level leaf holding one body the two functions agree
 * bit for bit.  The leaf's members are found from the cell key accumulated on the way down
 * (child index per level, as bho_body_keys builds it) in the key-sorted body list.
 * This is the specification the CUDA flag BH_FLAG_EXACT_LEAVES is tested against.
 * counters as above; [1] counts pair interactions inside such leaves individually. */
void bho_body_keys(const double* pos, int64_t n, const double bounds[4], int max_depth, uint32_t* keys);
typedef struct { uint32_t key; int64_t idx; } bho_keyidx;
static int bho_keyidx_cmp(const void* a, const void* b) {
    const bho_keyidx* x = (const bho_keyidx*)a; const bho_keyidx* y = (const bho_keyidx*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

void bho_compute_forces_exact_leaves(const bho_tree* t, const double* pos, const double* mass, int64_t n,
                                     const bho_params* p, int64_t i0, int64_t i1, int64_t stride,
                                     int nthreads, double* forces, int64_t* counters) {
    int64_t c_vis = 0, c_int = 0, c_open = 0, c_zero = 0, c_self = 0, c_stack = 0;
    const double* nodes = t->nodes;
    const double G = p->G, theta = p->theta, deps = p->dist_eps, meps = p->mass_eps;
    const int cap = t->max_depth;
    /* key-sorted body list (stable in body index) */
    uint32_t* keys = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    bho_keyidx* order = (bho_keyidx*)malloc(sizeof(bho_keyidx) * (size_t)(n > 0 ? n : 1));
    double bounds[4] = {nodes[XMIN], nodes[XMAX], nodes[YMIN], nodes[YMAX]};
    bho_body_keys(pos, n, bounds, cap, keys);
    for (int64_t i = 0; i < n; ++i) { order[i].key = keys[i]; order[i].idx = i; }
    qsort(order, (size_t)n, sizeof(bho_keyidx), bho_keyidx_cmp);
    int64_t nb = (i1 - i0 + stride - 1) / stride;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) \
    reduction(+ : c_vis, c_int, c_open, c_zero, c_self) reduction(max : c_stack)
#endif
    for (int64_t k = 0; k < nb; ++k) {
        int64_t i = i0 + k * stride;
        int64_t stack[4 * 64 + 8];
        uint32_t kstack[4 * 64 + 8];
        int dstack[4 * 64 + 8];
        int top = 0;
        double sx = 0.0, sy = 0.0;
        double pix = pos[2 * i], piy = pos[2 * i + 1];
        stack[top] = 0; kstack[top] = 0; dstack[top] = 1; ++top;
        while (top > 0) {
            if (top > c_stack) c_stack = top;
            --top;
            int64_t ni = stack[top];
            uint32_t code = kstack[top];
            int depth = dstack[top];
            const double* node = nodes + ni * QSZ;
            ++c_vis;
            double nodeMass = node[MASS];
            if (nodeMass <= meps) { ++c_zero; continue; }
            int occ = (int)node[PIDX];
            int leaf = node[CH0] == -1 && node[CH1] == -1 && node[CH2] == -1 && node[CH3] == -1;
            double dx = node[COMX] - pix, dy = node[COMY] - piy;
            double d2 = dx * dx + dy * dy;
            double d = sqrt(d2) + deps;
            double w = node[XMAX] - node[XMIN], h = node[YMAX] - node[YMIN];
            double size = (w > h) ? w : h;
            if (leaf || (size / d < theta)) {
                if (leaf && (occ == i || (occ + 2) == -i)) { ++c_self; continue; }
                if (leaf && depth >= cap && occ == -1) {
                    /* members: bodies whose cell key == code */
                    int64_t lo = 0, hi = n;
                    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (order[mid].key < code) lo = mid + 1; else hi = mid; }
                    for (int64_t q = lo; q < n && order[q].key == code; ++q) {
                        int64_t j = order[q].idx;
                        if (j == i) { ++c_self; continue; }
                        double ex = pos[2 * j] - pix, ey = pos[2 * j + 1] - piy;
                        double e2 = ex * ex + ey * ey;
                        double e = sqrt(e2) + deps;
                        double fm = (G * mass[i] * mass[j]) / e2;
                        sx += fm * (ex / e); sy += fm * (ey / e);
                        ++c_int;
                    }
                    continue;
                }
                double fm = (G * mass[i] * nodeMass) / d2;
                double nx = dx / d, ny = dy / d;
                sx += fm * nx; sy += fm * ny;
                ++c_int;
            } else {
                ++c_open;
                for (int c = 0; c < 4; ++c) {
                    int64_t ch = (int64_t)(int)node[CH0 + c];
                    if (ch != -1) { stack[top] = ch; kstack[top] = (code << 2) | (uint32_t)c; dstack[top] = depth + 1; ++top; }
                }
            }
        }
        forces[2 * i] = sx; forces[2 * i + 1] = sy;
    }
    free(keys); free(order);
    if (counters) {
        counters[0] = c_vis; counters[1] = c_int; counters[2] = c_open; counters[3] = c_zero;
        counters[4] = c_self; counters[5] = c_stack;
    }
}

/* project.cu:795-817 updateAccelerations / updateVelocities / updatePositions. */
void bho_update(const double* forces, const double* mass, double* acc, double* vel, double* pos,
                int64_t n, double dt) {
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < 2; ++k) acc[2 * i + k] = forces[2 * i + k] / mass[i];
    for (int64_t i = 0; i < 2 * n; ++i) vel[i] += acc[i] * dt;
    for (int64_t i = 0; i < 2 * n; ++i) pos[i] += vel[i] * dt;
}

/* One iteration of the loop body of runSimulationCpu (project.cu:883-910).  Returns node count. */
int64_t bho_step(double* pos, double* vel, const double* mass, double* acc, double* forces, int64_t n,
                 const bho_params* p, int nthreads, int64_t* counters) {
    bho_tree* t = bho_build_tree(pos, mass, n, p);
    int64_t nn = t->size;
    bho_compute_forces(t, pos, mass, n, p, 0, n, 1, nthreads, forces, counters);
    bho_update(forces, mass, acc, vel, pos, n, p->dt);
    bho_tree_free(t);
    return nn;
}

/* 2(D-1)-bit cell key of every body: the path DetermineChild (project.cu:348-356) takes from the
 * root down to the cap level, child index of the first split in the most significant pair. */
void bho_body_keys(const double* pos, int64_t n, const double bounds[4], int max_depth, uint32_t* keys) {
    for (int64_t i = 0; i < n; ++i) {
        double node[QSZ];
        node[XMIN] = bounds[0]; node[XMAX] = bounds[1]; node[YMIN] = bounds[2]; node[YMAX] = bounds[3];
        uint32_t key = 0;
        for (int l = 1; l < max_depth; ++l) {
            int c = determine_child(pos[2 * i], pos[2 * i + 1], node);
            double mx = (node[XMIN] + node[XMAX]) / 2.0, my = (node[YMIN] + node[YMAX]) / 2.0;
            if (c & 1) node[XMIN] = mx; else node[XMAX] = mx;
            if (c & 2) node[YMIN] = my; else node[YMAX] = my;
            key = (key << 2) | (uint32_t)c;
        }
        keys[i] = key;
    }
}

/* Canonical node table: DFS pre-order, children 0->3 (the order of TraverseTreeToFile,
 * project.cu:504-534).  Row = { depth (root 0), xmin, xmax, ymin, ymax, mass, comx, comy,
 * occupant (PARTICLE_INDEX as stored), is_internal }.  Independent of insertion order except for
 * last-bit rounding of mass / COM.  Returns rows written (<= cap). */
static int64_t canon_rec(const bho_tree* t, int64_t ni, int depth, double* out, int64_t cap, int64_t at) {
    const double* q = t->nodes + ni * QSZ;
    if (at < cap) {
        double* r = out + at * 10;
        r[0] = depth; r[1] = q[XMIN]; r[2] = q[XMAX]; r[3] = q[YMIN]; r[4] = q[YMAX];
        r[5] = q[MASS]; r[6] = q[COMX]; r[7] = q[COMY]; r[8] = q[PIDX]; r[9] = q[CH0] != -1;
    }
    ++at;
    for (int c = 0; c < 4; ++c)
        if (q[CH0 + c] != -1) at = canon_rec(t, (int64_t)q[CH0 + c], depth + 1, out, cap, at);
    return at;
}
int64_t bho_canonical_table(const bho_tree* t, double* out, int64_t cap_rows) {
    return canon_rec(t, 0, 0, out, cap_rows, 0);
}

/* project.cu:504-534 TraverseTreeToFile.  Default ostream formatting of double == "%g".
 * Deviation: for cap-level single leaves (occupant <= -2) the reference
 * indexes positions[] with the NEGATIVE encoded index (out-of-bounds read); this writer prints
 * the occupant's real position instead.  Comparisons with reference dumps mask those fields. */
static void dump_rec(const bho_tree* t, int64_t ni, const double* pos, int depth, FILE* f) {
    const double* q = t->nodes + ni * QSZ;
    fprintf(f, "%d %g %g %g %g %g", depth, q[XMIN], q[XMAX], q[YMIN], q[YMAX], q[MASS]);
    int occ = (int)q[PIDX];
    if (occ != -1) {
        int64_t b = occ >= 0 ? occ : (-(int64_t)occ - 2);
        fprintf(f, " occupantIndex=%d occupantPos=(%g,%g)", occ, pos[2 * b], pos[2 * b + 1]);
    } else if (q[MASS] > 0) {
        fprintf(f, " occupantIndex=%d occupantPos=(%g,%g)", occ, q[COMX], q[COMY]);
    }
    fputc('\n', f);
    for (int c = 0; c < 4; ++c)
        if ((int)q[CH0 + c] != -1) dump_rec(t, (int64_t)q[CH0 + c], pos, depth + 1, f);
}
int bho_dump_quadtree(const bho_tree* t, const double* pos, const char* path) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    dump_rec(t, 0, pos, 0, f);
    fclose(f);
    return 0;
}

/* Direct all-pairs force, the formula of main_approach_1.cpp:53-75:
 * F_i = sum_{j != i} G m_i m_j (r_j - r_i) / (d^2 * d), no softening.  Bodies i0..i1. */
void bho_direct_forces(const double* pos, const double* mass, int64_t n, double G, int64_t i0,
                       int64_t i1, int nthreads, double* forces) {
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int64_t i = i0; i < i1; ++i) {
        double sx = 0.0, sy = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            double dx = pos[2 * j] - pos[2 * i], dy = pos[2 * j + 1] - pos[2 * i + 1];
            double d2 = 0.0;
            d2 += dx * dx; d2 += dy * dy;
            double d = sqrt(d2);
            double factor = G * mass[i] * mass[j] / (d2 * d);
            sx += factor * dx; sy += factor * dy;
        }
        forces[2 * i] = sx; forces[2 * i + 1] = sy;
    }
}

int bho_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
