/* TEST INFRASTRUCTURE — NOT part of the product path. See vvoracle.h.
 *
 * Single-threaded C restatement of the reference's per-step particle path. Every function
 * cites the reference lines it follows (paths relative to /root/reference). Arithmetic is
 * written operation for operation like the reference (TVec helpers: `a/c` is `a*(1./c)`,
 * libvvhd/headers/TVec.hpp:24) and compiled with -ffp-contract=off.
 */
#include "vvoracle.h"

#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TREE_MAX_LIST 16 /* TSortedTree.cpp:10 */
static const double PI = 3.14159265358979323846; /* elementary.h:3-8 */
#define C_2PI (2. * PI)
#define C_1_2PI (1. / (2. * PI))
#define C_1_PI (1. / PI)

static double sqr(double v) { return v * v; }
static int sgn(double v) { return (v > 0) ? 1 : ((v < 0) ? -1 : 0); } /* TObj.hpp:8 */
static double dmax(double a, double b) { return (a < b) ? b : a; }     /* std::max */
static double dmin(double a, double b) { return (b < a) ? b : a; }     /* std::min */
/* MConvectiveFast.cpp:165 writes a bare `abs(src.g)`: with the reference's includes and -std=c++11 that is ::abs(int),
 * so the strength is truncated to an integer first (no regularisation at all for |g| < 1). Pinned against the
 * compiled reference by tests/test_oracle_port.py::test_sinks_match_reference. */
static double sink_abs(double g) { int t = (int)g; return (double)(t < 0 ? -t : t); }

/* bench.py's bounded CPU sample: epsilon / convective / diffusive visit every stride-th leaf only */
static int64_t g_stride = 1, g_phase = 0;
void vvo_set_leaf_sample(int64_t stride, int64_t phase) { g_stride = stride > 0 ? stride : 1; g_phase = phase; }
#define LEAF_LOOP(l, t) for (int64_t l = g_phase % g_stride; l < (t)->n_leaves; l += g_stride)

/* ------------------------------------------------------------------ tree */

typedef struct {
    vvo_tree* t;
    vvo_plist* p;
    const vvo_bodies* b;
    int far_criteria;
    double min_node, max_node;
    int64_t* seg_tmp;
} build_ctx;

static void tree_reserve(vvo_tree* t, int64_t need) {
    if (need <= t->cap) return;
    int64_t cap = t->cap ? t->cap : 64;
    while (cap < need) cap *= 2;
#define GROW(f, k) t->f = realloc(t->f, sizeof(*t->f) * (size_t)(cap * (k)))
    GROW(x, 1); GROW(y, 1); GROW(h, 1); GROW(w, 1); GROW(cmp, 3); GROW(cmm, 3);
    GROW(vfirst, 1); GROW(vlast, 1); GROW(sfirst, 1); GROW(slast, 1);
    GROW(ch1, 1); GROW(ch2, 1); GROW(leaf, 1); GROW(depth, 1);
#undef GROW
    t->cap = cap;
}

static int64_t new_node(vvo_tree* t, int64_t vf, int64_t vl, int64_t sf, int64_t sl, int64_t depth) {
    tree_reserve(t, t->n_nodes + 1);
    int64_t id = t->n_nodes++;
    t->vfirst[id] = vf; t->vlast[id] = vl; t->sfirst[id] = sf; t->slast[id] = sl;
    t->ch1[id] = t->ch2[id] = -1; t->leaf[id] = -1; t->depth[id] = depth;
    return id;
}

/* snode::Stretch, TSortedTree.cpp:101-137 (vortex range + body segments) */
static void stretch(build_ctx* c, int64_t id) {
    vvo_tree* t = c->t;
    double trx = -DBL_MAX, try_ = -DBL_MAX, blx = DBL_MAX, bly = DBL_MAX;
    for (int64_t k = t->sfirst[id]; k < t->slast[id]; k++) {
        int64_t s = t->seg_perm[k];
        trx = dmax(trx, c->b->rx[s]); try_ = dmax(try_, c->b->ry[s]);
        blx = dmin(blx, c->b->rx[s]); bly = dmin(bly, c->b->ry[s]);
    }
    for (int64_t i = t->vfirst[id]; i < t->vlast[id]; i++) {
        trx = dmax(trx, c->p->x[i]); try_ = dmax(try_, c->p->y[i]);
        blx = dmin(blx, c->p->x[i]); bly = dmin(bly, c->p->y[i]);
    }
    t->x[id] = (blx + trx) * 0.5;
    t->y[id] = (bly + try_) * 0.5;
    t->h[id] = try_ - bly;
    t->w[id] = trx - blx;
}

static void pswap(vvo_plist* p, int64_t a, int64_t b) {
#define SW(f) { double tmp = p->f[a]; p->f[a] = p->f[b]; p->f[b] = tmp; }
    SW(x) SW(y) SW(g) SW(vx) SW(vy) SW(ieps)
#undef SW
    int64_t o = p->orig[a]; p->orig[a] = p->orig[b]; p->orig[b] = o;
}

/* snode::DivideNode, TSortedTree.cpp:36-79 */
static void divide(build_ctx* c, int64_t id) {
    vvo_tree* t = c->t;
    double h = t->h[id], w = t->w[id], x = t->x[id], y = t->y[id];
    int is_leaf = 0;
    if (dmax(h, w) < c->max_node && dmin(h, w) <= c->min_node) is_leaf = 1;
    if (!is_leaf) {
        int64_t m = t->slast[id] - t->sfirst[id];
        if (t->vlast[id] - t->vfirst[id] > m) m = t->vlast[id] - t->vfirst[id];
        if (dmax(h, w) < c->max_node && m < TREE_MAX_LIST) is_leaf = 1;
    }
    if (is_leaf) {
        t->leaf[id] = t->n_leaves++;
        return;
    }
    int xsplit = (h < w);
    /* DistributeContent(range&), TSortedTree.cpp:81-99: Hoare partition, unstable */
    int64_t first = t->vfirst[id], last = t->vlast[id];
    int64_t p1 = first, p2 = last;
    vvo_plist* p = c->p;
    while (p1 < p2) {
        while (p1 < last && (xsplit ? (p->x[p1] < x) : (p->y[p1] < y))) p1++;
        do p2--;
        while (p2 >= first && (xsplit ? (p->x[p2] >= x) : (p->y[p2] >= y)));
        if (p1 < p2) pswap(p, p1, p2);
    }
    /* DistributeContent(LList&), TSortedTree.cpp:139-148: stable split of the segment list */
    int64_t sf = t->sfirst[id], sl = t->slast[id], n1 = 0, n2 = 0;
    for (int64_t k = sf; k < sl; k++) {
        int64_t s = t->seg_perm[k];
        if (xsplit ? (c->b->rx[s] < x) : (c->b->ry[s] < y)) t->seg_perm[sf + n1++] = s;
        else c->seg_tmp[n2++] = s;
    }
    for (int64_t k = 0; k < n2; k++) t->seg_perm[sf + n1 + k] = c->seg_tmp[k];

    int64_t d = t->depth[id];
    /* pre-order ids: child 1's whole subtree is numbered before child 2 */
    int64_t c1 = new_node(t, first, p1, sf, sf + n1, d + 1);
    t->ch1[id] = c1;
    stretch(c, c1);
    divide(c, c1);
    int64_t c2 = new_node(t, p1, last, sf + n1, sl, d + 1);
    t->ch2[id] = c2;
    stretch(c, c2);
    divide(c, c2);
    /* bllist.clear() on the parent, TSortedTree.cpp:71 */
    t->sfirst[id] = t->slast[id] = sf;
}

/* snode::CalculateCMass / FromScratch, TSortedTree.cpp:150-197 */
static void cmass(build_ctx* c, int64_t id) {
    vvo_tree* t = c->t;
    double* P = t->cmp + 3 * id;
    double* M = t->cmm + 3 * id;
    if (t->ch1[id] < 0) {
        P[0] = P[1] = P[2] = M[0] = M[1] = M[2] = 0;
        for (int64_t i = t->vfirst[id]; i < t->vlast[id]; i++) {
            double g = c->p->g[i];
            if (g > 0) { P[0] += c->p->x[i] * g; P[1] += c->p->y[i] * g; P[2] += g; }
            else { M[0] += c->p->x[i] * g; M[1] += c->p->y[i] * g; M[2] += g; }
        }
        if (P[2]) { double r = 1. / P[2]; P[0] *= r; P[1] *= r; } else { P[0] = t->x[id]; P[1] = t->y[id]; }
        if (M[2]) { double r = 1. / M[2]; M[0] *= r; M[1] *= r; } else { M[0] = t->x[id]; M[1] = t->y[id]; }
        return;
    }
    int64_t a = t->ch1[id], b2 = t->ch2[id];
    cmass(c, a);
    cmass(c, b2);
    for (int s = 0; s < 2; s++) {
        double* cm = s ? M : P;
        const double* A = (s ? t->cmm : t->cmp) + 3 * a;
        const double* B = (s ? t->cmm : t->cmp) + 3 * b2;
        double sumg = A[2] + B[2];
        if (sumg) {
            double r = 1. / sumg;
            cm[0] = (A[0] * A[2] + B[0] * B[2]) * r;
            cm[1] = (A[1] * A[2] + B[1] * B[2]) * r;
            cm[2] = sumg;
        } else { cm[0] = t->x[id]; cm[1] = t->y[id]; cm[2] = 0; }
    }
}

static void push_list(int64_t** arr, int64_t* cap, int64_t n, int64_t v) {
    if (n >= *cap) { *cap = *cap ? *cap * 2 : 1024; *arr = realloc(*arr, sizeof(int64_t) * (size_t)*cap); }
    (*arr)[n] = v;
}

/* snode::FindNearNodes, TSortedTree.cpp:199-217 */
static void find_near(build_ctx* c, int64_t leafnode, int64_t top, int64_t* nn, int64_t* nf) {
    vvo_tree* t = c->t;
    double drx = t->x[top] - t->x[leafnode], dry = t->y[top] - t->y[leafnode];
    double hp = t->h[top] + t->w[top] + t->h[leafnode] + t->w[leafnode];
    if (drx * drx + dry * dry > c->far_criteria * hp * hp) {
        push_list(&t->far_idx, &t->far_cap, (*nf)++, top);
        return;
    }
    if (t->ch1[top] >= 0) {
        find_near(c, leafnode, t->ch1[top], nn, nf);
        find_near(c, leafnode, t->ch2[top], nn, nf);
        return;
    }
    push_list(&t->near_idx, &t->near_cap, (*nn)++, t->leaf[top]);
}

/* stree::build, TSortedTree.cpp:232-265 */
vvo_tree* vvo_tree_build(vvo_plist* p, const vvo_bodies* b, int far_criteria, double min_node, double max_node) {
    static const vvo_bodies nobody;
    if (!b) b = &nobody;
    vvo_tree* t = calloc(1, sizeof(vvo_tree));
    build_ctx c = {t, p, b, far_criteria, min_node, max_node, NULL};
    t->nseg = b->nseg;
    t->seg_perm = malloc(sizeof(int64_t) * (size_t)(b->nseg + 1));
    c.seg_tmp = malloc(sizeof(int64_t) * (size_t)(b->nseg + 1));
    for (int64_t k = 0; k < b->nseg; k++) t->seg_perm[k] = k;
    int64_t root = new_node(t, 0, p->n, 0, b->nseg, 0);
    stretch(&c, root);
    divide(&c, root);
    cmass(&c, root);
    t->leaf_node = malloc(sizeof(int64_t) * (size_t)(t->n_leaves + 1));
    for (int64_t i = 0; i < t->n_nodes; i++)
        if (t->leaf[i] >= 0) t->leaf_node[t->leaf[i]] = i;
    t->near_ptr = malloc(sizeof(int64_t) * (size_t)(t->n_leaves + 1));
    t->far_ptr = malloc(sizeof(int64_t) * (size_t)(t->n_leaves + 1));
    int64_t nn = 0, nf = 0;
    for (int64_t l = 0; l < t->n_leaves; l++) {
        t->near_ptr[l] = nn; t->far_ptr[l] = nf;
        find_near(&c, t->leaf_node[l], root, &nn, &nf);
    }
    t->near_ptr[t->n_leaves] = nn; t->far_ptr[t->n_leaves] = nf;
    free(c.seg_tmp);
    return t;
}

void vvo_tree_free(vvo_tree* t) {
    if (!t) return;
    free(t->x); free(t->y); free(t->h); free(t->w); free(t->cmp); free(t->cmm);
    free(t->vfirst); free(t->vlast); free(t->sfirst); free(t->slast);
    free(t->ch1); free(t->ch2); free(t->leaf); free(t->depth); free(t->leaf_node); free(t->seg_perm);
    free(t->near_ptr); free(t->near_idx); free(t->far_ptr); free(t->far_idx);
    free(t);
}

/* stree::findNode, TSortedTree.cpp:284-303 */
int64_t vvo_find_node(const vvo_tree* t, double px, double py) {
    int64_t n = 0;
    while (t->ch1[n] >= 0) {
        if (t->h[n] < t->w[n]) n = (px < t->x[n]) ? t->ch1[n] : t->ch2[n];
        else n = (py < t->y[n]) ? t->ch1[n] : t->ch2[n];
    }
    return n;
}

/* ------------------------------------------------------------------ epsilon */

/* MEpsilonFast::nearestBodySegment, MEpsilonFast.cpp:214-253 */
static int64_t nearest_body_segment(const vvo_tree* t, const vvo_bodies* b, int64_t l, double px, double py) {
    int64_t att = -1;
    double res = DBL_MAX;
    for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
        int64_t nn = t->leaf_node[t->near_idx[k]];
        for (int64_t q = t->sfirst[nn]; q < t->slast[nn]; q++) {
            int64_t s = t->seg_perm[q];
            double d = sqr(px - b->rx[s]) + sqr(py - b->ry[s]);
            if (d < res) { res = d; att = s; }
        }
    }
    if (att >= 0) return att;
    for (int64_t ib = 0; ib < b->nbody; ib++) {
        /* `lobj += 9` skip-ahead on non-improving candidates (:248), then the loop's own ++ */
        for (int64_t s = b->bfirst[ib]; s < b->bfirst[ib + 1]; s++) {
            double d = sqr(px - b->rx[s]) + sqr(py - b->ry[s]);
            if (d < res) { res = d; att = s; }
            else s += 9;
        }
    }
    return att;
}

/* MEpsilonFast::MergeVortexes, MEpsilonFast.cpp:111-126 */
static void merge_vortexes(vvo_plist* p, int64_t a, int64_t b) {
    if (sgn(p->g[a]) == sgn(p->g[b])) {
        double r = 1. / (p->g[a] + p->g[b]);
        p->x[a] = (p->x[a] * p->g[a] + p->x[b] * p->g[b]) * r;
        p->y[a] = (p->y[a] * p->g[a] + p->y[b] * p->g[b]) * r;
    } else if (fabs(p->g[a]) < fabs(p->g[b])) {
        p->x[a] = p->x[b]; p->y[a] = p->y[b];
    }
    p->g[a] += p->g[b];
    p->g[b] = 0;
}

/* MEpsilonFast::epsv, MEpsilonFast.cpp:128-173 */
static double epsv(const vvo_tree* t, vvo_plist* p, int64_t l, int64_t lv, double merge_criteria_sq,
                   int64_t* merged) {
    double res1 = DBL_MAX, res2 = DBL_MAX;
    int64_t lv1 = -1, lv2 = -1;
    for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
        int64_t nn = t->leaf_node[t->near_idx[k]];
        for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
            if (!p->g[j] || j == lv) continue;
            double dx = p->x[lv] - p->x[j], dy = p->y[lv] - p->y[j];
            double d = dx * dx + dy * dy;
            if (res1 > d) { res2 = res1; lv2 = lv1; res1 = d; lv1 = j; }
            else if (res2 > d) { res2 = d; lv2 = j; }
        }
    }
    if (lv1 < 0) return DBL_MIN;
    if (lv2 < 0) return sqrt(res1);
    if (isnan(merge_criteria_sq)) return sqrt(res2);
    if ((res1 < merge_criteria_sq) ||
        ((sgn(p->g[lv1]) == sgn(p->g[lv2])) && (sgn(p->g[lv1]) != sgn(p->g[lv])))) {
        (*merged)++;
        merge_vortexes(p, lv, lv1);
        return epsv(t, p, l, lv, NAN, merged);
    }
    return sqrt(res2);
}

/* MEpsilonFast::CalcEpsilonFast, MEpsilonFast.cpp:11-63 (vortex list) */
int64_t vvo_epsilon(const vvo_tree* t, vvo_plist* p, const vvo_bodies* b, int merge) {
    static const vvo_bodies nobody;
    if (!b) b = &nobody;
    int64_t merged = 0;
    LEAF_LOOP(l, t) {
        int64_t node = t->leaf_node[l];
        double cx = t->x[node], cy = t->y[node];
        int64_t att = nearest_body_segment(t, b, l, cx, cy);
        double merge_criteria_sq;
        if (merge) {
            merge_criteria_sq = (att >= 0)
                ? 0.16 * (sqr(b->dlx[att]) + sqr(b->dly[att])) *
                      (1. + sqrt(sqr(cx - b->rx[att]) + sqr(cy - b->ry[att])))
                : 0;
        } else merge_criteria_sq = NAN;
        double eps_restriction = (att >= 0) ? sqrt(sqr(b->dlx[att]) + sqr(b->dly[att])) * (1.0 / 3.0) : 0;
        for (int64_t i = t->vfirst[node]; i < t->vlast[node]; i++) {
            if (!p->g[i]) continue;
            p->ieps[i] = 1.0 / dmax(epsv(t, p, l, i, merge_criteria_sq, &merged), eps_restriction);
        }
    }
    return merged;
}

/* ------------------------------------------------------------------ convective */

/* MConvectiveFast::SegmentInfluence_linear_source, MConvectiveFast.cpp:440-457 */
static void linear_source(double px, double py, const vvo_bodies* b, int64_t s, double q1, double q2, double* ox,
                          double* oy) {
    double complex z = px + I * py;
    double complex zc = b->rx[s] + I * b->ry[s];
    double complex dz = b->dlx[s] + I * b->dly[s];
    double complex zs1 = zc - dz * 0.5;
    double complex zs2 = zc + dz * 0.5;
    double complex z1 = z - zs1;
    double complex z2 = z - zs2;
    double complex zV = ((q2 - q1) - conj(q2 * z1 - q1 * z2) / conj(dz) * clog(conj(z2) / conj(z1))) / conj(dz);
    *ox = creal(zV); *oy = cimag(zV);
}

/* MConvectiveFast::body_list_influence, MConvectiveFast.cpp:172-215 */
static void body_list_influence(const vvo_bodies* b, double px, double py, double* ox, double* oy) {
    double rx = 0, ry = 0;
    for (int64_t ib = 0; ib < b->nbody; ib++) {
        const double* bp = b->bprop + 13 * ib;
        int any_slip = 0;
        for (int64_t s = b->bfirst[ib]; s < b->bfirst[ib + 1]; s++) any_slip |= (b->slip[s] != 0);
        if (any_slip) {
            for (int64_t s = b->bfirst[ib]; s < b->bfirst[ib + 1]; s++) {
                if (!b->slip[s]) continue;
                double dx = px - b->rx[s], dy = py - b->ry[s];
                double k = b->g[s] / (dx * dx + dy * dy + sqr(1. / b->ieps[s]));
                rx += -dy * k; ry += dx * k;
            }
        }
        double sx = bp[10], sy = bp[11], so = bp[12], ax = bp[0], ay = bp[1];
        if (!(fabs(sx) + fabs(sy) + fabs(so) < 1E-10)) { /* TVec3D::iszero */
            for (int64_t s = b->bfirst[ib]; s < b->bfirst[ib + 1]; s++) {
                double dx = px - b->rx[s], dy = py - b->ry[s];
                double drabs2 = dx * dx + dy * dy;
                double dlx = b->dlx[s], dly = b->dly[s];
                if (drabs2 < dlx * dlx + dly * dly) {
                    /* Vs1 = speed.r + speed.o*rotl(corner - axis) */
                    double ux = b->cx[s] - ax, uy = b->cy[s] - ay;
                    double v1x = sx + so * (-uy), v1y = sy + so * ux;
                    double g1 = -(v1x * dlx + v1y * dly);
                    double q1 = -((-v1y) * dlx + v1x * dly);
                    double wx = b->cx[s] + dlx - ax, wy = b->cy[s] + dly - ay;
                    double v2x = sx + so * (-wy), v2y = sy + so * wx;
                    double g2 = -(v2x * dlx + v2y * dly);
                    double q2 = -((-v2y) * dlx + v2x * dly);
                    double ix, iy;
                    linear_source(px, py, b, s, g1, g2, &ix, &iy);
                    rx += -iy; ry += ix;
                    linear_source(px, py, b, s, q1, q2, &ix, &iy);
                    rx += ix; ry += iy;
                } else {
                    double ux = b->rx[s] - ax, uy = b->ry[s] - ay;
                    double vx = sx + so * (-uy), vy = sy + so * ux;
                    double g = -(vx * dlx + vy * dly);
                    double q = -((-vy) * dlx + vx * dly);
                    double r = 1. / drabs2;
                    rx += (dx * q + (-dy) * g) * r;
                    ry += (dy * q + dx * g) * r;
                }
            }
        }
    }
    *ox = rx * C_1_2PI; *oy = ry * C_1_2PI;
}

/* MConvectiveFast::process_all_lists, MConvectiveFast.cpp:36-114 (vortex list) */
void vvo_convective(const vvo_tree* t, vvo_plist* p, const vvo_bodies* b, double inf_vx, double inf_vy, double dt,
                    const double* sinks, int64_t nsink) {
    static const vvo_bodies nobody;
    if (!b) b = &nobody;
    LEAF_LOOP(l, t) {
        int64_t node = t->leaf_node[l];
        double cx = t->x[node], cy = t->y[node];
        double T1 = 0, T2 = 0, T3 = 0, T4 = 0;
        for (int64_t k = t->far_ptr[l]; k < t->far_ptr[l + 1]; k++) {
            const double* P = t->cmp + 3 * t->far_idx[k];
            const double* M = t->cmm + 3 * t->far_idx[k];
            double dpx = cx - P[0], dpy = cy - P[1], dmx = cx - M[0], dmy = cy - M[1];
            double ap = dpx * dpx + dpy * dpy, am = dmx * dmx + dmy * dmy;
            double fp1 = P[2] / ap, fm1 = M[2] / am;
            double fp2 = P[2] / sqr(ap), fm2 = M[2] / sqr(am);
            T1 -= (fp1 * dpy + fm1 * dmy);
            T2 += (fp1 * dpx + fm1 * dmx);
            T3 += (fp2 * dpy * dpx + fm2 * dmy * dmx);
            T4 += (fp2 * (sqr(dpy) - sqr(dpx)) + fm2 * (sqr(dmy) - sqr(dmx)));
        }
        T1 *= C_1_2PI; T2 *= C_1_2PI; T3 *= C_1_PI; T4 *= C_1_2PI;
        for (int64_t i = t->vfirst[node]; i < t->vlast[node]; i++) {
            if (!p->g[i]) continue;
            double px = p->x[i], py = p->y[i];
            double dlx = px - cx, dly = py - cy;
            p->vx[i] += inf_vx; p->vy[i] += inf_vy;
            /* near_nodes_influence + biot_savart, :116-137 */
            double rx = 0, ry = 0;
            for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
                int64_t nn = t->leaf_node[t->near_idx[k]];
                for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
                    if (!p->g[j]) continue;
                    double dx = px - p->x[j], dy = py - p->y[j];
                    double q = p->g[j] / (dx * dx + dy * dy + sqr(1. / p->ieps[j]));
                    rx += -dy * q; ry += dx * q;
                }
            }
            p->vx[i] += rx * C_1_2PI; p->vy[i] += ry * C_1_2PI;
            /* sink_list_influence, :153-170 */
            double sx = 0, sy = 0;
            double eps2_div_srcg = dt * C_1_PI;
            for (int64_t k = 0; k < nsink; k++) {
                double dx = px - sinks[3 * k], dy = py - sinks[3 * k + 1], sg = sinks[3 * k + 2];
                double q = sg / (dx * dx + dy * dy + eps2_div_srcg * sink_abs(sg));
                sx += dx * q; sy += dy * q;
            }
            p->vx[i] += sx * C_1_2PI; p->vy[i] += sy * C_1_2PI;
            double bx, by;
            body_list_influence(b, px, py, &bx, &by);
            p->vx[i] += bx; p->vy[i] += by;
            p->vx[i] += T1; p->vy[i] += T2;
            p->vx[i] += T3 * dlx + T4 * dly;
            p->vy[i] += T4 * dlx + (-T3) * dly;
        }
    }
}

/* MConvectiveFast::velocity, MConvectiveFast.cpp:20-34: velocity at an arbitrary point = near particles of the
 * point's leaf (near_nodes_influence, :116-137) + the leaf's far nodes as eps = 0 monopoles AT THE POINT
 * (far_nodes_influence, :139-151; CMp/CMm carry _1_eps = inf, TSortedTree.cpp:25,179) + sinks + bodies + inf_speed.
 * Note the far part differs from the step loop's Taylor expansion about the leaf centre (:48-69). */
void vvo_velocity_at(const vvo_tree* t, const vvo_plist* p, const vvo_bodies* b, double inf_vx, double inf_vy, double dt,
                     const double* sinks, int64_t nsink, const double* xy, int64_t npts, double* out) {
    static const vvo_bodies nobody;
    if (!b) b = &nobody;
    for (int64_t q = 0; q < npts; q++) {
        const double px = xy[2 * q], py = xy[2 * q + 1];
        const int64_t node = vvo_find_node(t, px, py);
        const int64_t l = t->leaf[node];
        double resx = 0, resy = 0;
        double rx = 0, ry = 0;
        for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
            int64_t nn = t->leaf_node[t->near_idx[k]];
            for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
                if (!p->g[j]) continue;
                double dx = px - p->x[j], dy = py - p->y[j];
                double w = p->g[j] / (dx * dx + dy * dy + sqr(1. / p->ieps[j]));
                rx += -dy * w; ry += dx * w;
            }
        }
        resx += rx * C_1_2PI; resy += ry * C_1_2PI;
        rx = 0; ry = 0;
        for (int64_t k = t->far_ptr[l]; k < t->far_ptr[l + 1]; k++) {
            const double* P = t->cmp + 3 * t->far_idx[k];
            const double* M = t->cmm + 3 * t->far_idx[k];
            double dx = px - P[0], dy = py - P[1];
            double w = P[2] / (dx * dx + dy * dy + 0.);
            rx += -dy * w; ry += dx * w;
            dx = px - M[0]; dy = py - M[1];
            w = M[2] / (dx * dx + dy * dy + 0.);
            rx += -dy * w; ry += dx * w;
        }
        resx += rx * C_1_2PI; resy += ry * C_1_2PI;
        double sx = 0, sy = 0;
        double eps2_div_srcg = dt * C_1_PI;
        for (int64_t k = 0; k < nsink; k++) {
            double dx = px - sinks[3 * k], dy = py - sinks[3 * k + 1], sg = sinks[3 * k + 2];
            double w = sg / (dx * dx + dy * dy + eps2_div_srcg * sink_abs(sg));
            sx += dx * w; sy += dy * w;
        }
        resx += sx * C_1_2PI; resy += sy * C_1_2PI;
        double bx, by;
        body_list_influence(b, px, py, &bx, &by);
        resx += bx; resy += by;
        resx += inf_vx; resy += inf_vy;
        out[2 * q] = resx; out[2 * q + 1] = resy;
    }
}

/* MEpsilonFast::eps2h and ::h2 (static, MEpsilonFast.cpp:66-107) for node = findNode(p): squared distance to the
 * second-nearest particle of the near leaves (zero distances skipped, g ignored; nearest if only one; lowest() if
 * none) and squared distance to the nearest body segment of the near leaves (+inf if none). out = (eps2h, h2). */
void vvo_eps2h_h2_at(const vvo_tree* t, const vvo_plist* p, const vvo_bodies* b, const double* xy, int64_t npts,
                     double* out) {
    for (int64_t q = 0; q < npts; q++) {
        const double px = xy[2 * q], py = xy[2 * q + 1];
        const int64_t l = t->leaf[vvo_find_node(t, px, py)];
        double res1 = INFINITY, res2 = INFINITY, hh = INFINITY;
        for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
            int64_t nn = t->leaf_node[t->near_idx[k]];
            for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
                double dx = px - p->x[j], dy = py - p->y[j];
                double d = dx * dx + dy * dy;
                if (!d) continue;
                else if (d < res1) { res2 = res1; res1 = d; }
                else if (d < res2) res2 = d;
            }
            for (int64_t k2 = t->sfirst[nn]; k2 < t->slast[nn]; k2++) {
                int64_t s = t->seg_perm[k2];
                double dx = px - b->rx[s], dy = py - b->ry[s];
                hh = dmin(hh, dx * dx + dy * dy);
            }
        }
        out[2 * q] = isfinite(res2) ? res2 : (isfinite(res1) ? res1 : -DBL_MAX);
        out[2 * q + 1] = hh;
    }
}

/* MConvectiveFast::_2PI_Xi_g (MConvectiveFast.cpp:286-310) with _2PI_Xi_g_near (:275-279) and _2PI_Xi_g_dist
 * (:281-284): 2 pi Xi_gamma of a vortex at (px, py) with core radius rd on segment s */
static double xi_g_dist(double px, double py, double p1x, double p1y, double p2x, double p2y) {
    return 0.5 * log((sqr(px - p2x) + sqr(py - p2y)) / (sqr(px - p1x) + sqr(py - p1y)));
}
static double xi_g_near(double px, double py, double pcx, double pcy, double dlx, double dly, double rd) {
    return ((pcx - px) * dlx + (pcy - py) * dly) / sqr(rd);
}
static double xi_g(double px, double py, const vvo_bodies* b, int64_t s, double rd) {
    const double cx = b->cx[s], cy = b->cy[s], dlx = b->dlx[s], dly = b->dly[s], rx = b->rx[s], ry = b->ry[s];
    double rd_sqr = sqr(rd);
    double dr1_sqr = sqr(px - cx) + sqr(py - cy);
    double dr2_sqr = sqr(px - cx - dlx) + sqr(py - cy - dly);
    if (dr1_sqr >= rd_sqr && dr2_sqr >= rd_sqr) return 0.5 * log(dr2_sqr / dr1_sqr);
    else if (dr1_sqr <= rd_sqr && dr2_sqr <= rd_sqr) return xi_g_near(px, py, rx, ry, dlx, dly, rd);
    else {
        double a0 = dlx * dlx + dly * dly;
        double b0 = (px - rx) * dlx + (py - ry) * dly;
        double d = sqrt(b0 * b0 - a0 * ((sqr(px - rx) + sqr(py - ry)) - rd * rd));
        double k = (b0 + d) / a0; if ((k <= -0.5) || (k >= 0.5)) k = (b0 - d) / a0;
        double p3x = rx + k * dlx, p3y = ry + k * dly;
        if (dr1_sqr < rd_sqr)
            return xi_g_near(px, py, 0.5 * (p3x + cx), 0.5 * (p3y + cy), p3x - cx, p3y - cy, rd) +
                   xi_g_dist(px, py, p3x, p3y, cx + dlx, cy + dly);
        else
            return xi_g_dist(px, py, cx, cy, p3x, p3y) +
                   xi_g_near(px, py, 0.5 * (cx + dlx + p3x), 0.5 * (cy + dly + p3y), cx + dlx - p3x, cy + dly - p3y, rd);
    }
}

/* MConvectiveFast::NodeInfluence(*findNode(seg.r), seg), MConvectiveFast.cpp:398-418, for every segment: the free
 * vortices' term of the slip equation's right-hand side (fillSlipEquationForSegment, :459-467) */
void vvo_node_influence(const vvo_tree* t, const vvo_plist* p, const vvo_bodies* b, double* out) {
    for (int64_t s = 0; s < b->nseg; s++) {
        const int64_t l = t->leaf[vvo_find_node(t, b->rx[s], b->ry[s])];
        double res = 0;
        for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
            int64_t nn = t->leaf_node[t->near_idx[k]];
            for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
                if (!p->g[j]) continue;
                res += xi_g(p->x[j], p->y[j], b, s, 1. / p->ieps[j]) * p->g[j];
            }
        }
        const double c2x = b->cx[s] + b->dlx[s], c2y = b->cy[s] + b->dly[s];
        for (int64_t k = t->far_ptr[l]; k < t->far_ptr[l + 1]; k++) {
            const double* P = t->cmp + 3 * t->far_idx[k];
            const double* M = t->cmm + 3 * t->far_idx[k];
            res += xi_g_dist(P[0], P[1], b->cx[s], b->cy[s], c2x, c2y) * P[2];
            res += xi_g_dist(M[0], M[1], b->cx[s], b->cy[s], c2x, c2y) * M[2];
        }
        out[s] = res * C_1_2PI;
    }
}

/* XVorticity::evaluate + XVorticity::vorticity (libvvhd/src/XVorticity.cpp:26-97), the vorticity raster of vvplot.
 * `p` is the vortex list AFTER MFlowmove::vortex_shed (:36, the caller appends the attached vortices); it is permuted
 * by the tree this function builds for itself (far criteria 8, minNodeSize 20 dl, :38). Per particle (:46-52):
 * v.x = 1 / (eps_mult^2 max(eps2h(leaf, r), (0.6 dl)^2)), v.y = v.x g. Per raster point (:58-69, 76-97): 0 inside a
 * body, else sum over the near leaves' particles of v.y exp(-|p - r|^2 v.x) where that exponent > -6, times 1/pi,
 * plus 0.5 (1 - erf(h2 / (dl eps_mult)^2)) where that argument < 3. xmin, ymin, dxdy are floats as in XField.
 * out[yj * xres + xi] in double (the reference stores the same value rounded to float). */
void vvo_vorticity_raster(vvo_plist* p, const vvo_bodies* b, float xmin, float ymin, float dxdy, int xres, int yres,
                          double eps_mult, double dl, double* out) {
    static const vvo_bodies nobody;
    if (!b) b = &nobody;
    vvo_tree* t = vvo_tree_build(p, b, 8, dl * 20, DBL_MAX);
    double* vx = (double*)malloc(sizeof(double) * (size_t)(p->n > 0 ? p->n : 1));
    double* vy = (double*)malloc(sizeof(double) * (size_t)(p->n > 0 ? p->n : 1));
    for (int64_t l = 0; l < t->n_leaves; l++) {
        const int64_t node = t->leaf_node[l];
        for (int64_t i = t->vfirst[node]; i < t->vlast[node]; i++) {
            double res1 = INFINITY, res2 = INFINITY;   /* MEpsilonFast::eps2h(leaf, r_i), MEpsilonFast.cpp:66-93 */
            for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
                int64_t nn = t->leaf_node[t->near_idx[k]];
                for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
                    double dx = p->x[i] - p->x[j], dy = p->y[i] - p->y[j];
                    double d = dx * dx + dy * dy;
                    if (!d) continue;
                    else if (d < res1) { res2 = res1; res1 = d; }
                    else if (d < res2) res2 = d;
                }
            }
            double e2 = isfinite(res2) ? res2 : (isfinite(res1) ? res1 : -DBL_MAX);
            vx[i] = 1. / (sqr(eps_mult) * dmax(e2, sqr(0.6 * dl)));
            vy[i] = vx[i] * p->g[i];
        }
    }
    for (int yj = 0; yj < yres; yj++) {
        for (int xi = 0; xi < xres; xi++) {
            const double px = (double)xmin + (double)dxdy * (double)xi, py = (double)ymin + (double)dxdy * (double)yj;
            int inbody = 0;
            for (int64_t ib = 0; ib < b->nbody && !inbody; ib++) inbody = vvo_point_invalid(b, ib, px, py) >= 0;
            if (inbody) { out[(size_t)yj * xres + xi] = 0; continue; }
            const int64_t l = t->leaf[vvo_find_node(t, px, py)];
            double res = 0, hh = INFINITY;
            for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
                int64_t nn = t->leaf_node[t->near_idx[k]];
                for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
                    double dx = px - p->x[j], dy = py - p->y[j];
                    double exparg = -(dx * dx + dy * dy) * vx[j];
                    res += (exparg > -6) ? vy[j] * exp(exparg) : 0;
                }
                for (int64_t k2 = t->sfirst[nn]; k2 < t->slast[nn]; k2++) {
                    int64_t s = t->seg_perm[k2];
                    double dx = px - b->rx[s], dy = py - b->ry[s];
                    hh = dmin(hh, dx * dx + dy * dy);
                }
            }
            res *= C_1_PI;
            double erfarg = hh / sqr(dl * eps_mult);
            res += (erfarg < 3) ? 0.5 * (1 - erf(erfarg)) : 0;
            out[(size_t)yj * xres + xi] = res;
        }
    }
    free(vx); free(vy);
    vvo_tree_free(t);
}

/* XPressure::evaluate + XPressure::pressure (libvvhd/src/XPressure.cpp:32-146), the pressure raster of vvplot.
 * `p` is the vortex list AFTER MFlowmove::vortex_shed (:56) and b->gsum the segments' gsum after it (vortex_shed adds g,
 * MFlowmove.cpp:227); the list is permuted by the tree this function builds for itself (far criteria 8, minNodeSize
 * 20 dl, maxNodeSize 0.1, :27). One ordinary velocity pass (epsilon without merging, convective, diffusive, :61-64),
 * then per raster point (:84-92): 0 inside a body, else (:108-146)
 *   2pi Cp = sum over segments [(rotl(K) g_s + K q_s) . Vs] - sum over segments (dl/dt . rotl(K)) (running sum of gsum)
 *          + sum over ALL vortices (v . rotl(K(r, p))) g,      K(o, p) = (p - o) / |p - o|^2
 *   Cp = 2pi Cp / 2pi + (|inf_speed|^2 - |velocity(p)|^2) / 2  [+ |velocity(p) - ref_speed|^2 / 2 unless ref_frame 's'].
 * b->fric is left as it was (the diffusive pass adds to it). out[yj * xres + xi] in double. */
void vvo_diffusive(const vvo_tree* t, vvo_plist* p, vvo_bodies* b, double re);
void vvo_pressure_raster(vvo_plist* p, vvo_bodies* b, float xmin, float ymin, float dxdy, int xres, int yres, double dl,
                         double re, double dt, double inf_vx, double inf_vy, const double* sinks, int64_t nsink,
                         int use_ref_speed, double ref_vx, double ref_vy, double* out) {
    static vvo_bodies nobody;
    if (!b) b = &nobody;
    vvo_tree* t = vvo_tree_build(p, b, 8, dl * 20, 0.1);
    double* fric_keep = (double*)malloc(sizeof(double) * (size_t)(b->nseg > 0 ? b->nseg : 1));
    for (int64_t s = 0; s < b->nseg; s++) fric_keep[s] = b->fric[s];
    vvo_epsilon(t, p, b, 0);
    vvo_convective(t, p, b, inf_vx, inf_vy, dt, sinks, nsink);
    vvo_diffusive(t, p, b, re);
    for (int64_t s = 0; s < b->nseg; s++) b->fric[s] = fric_keep[s];
    free(fric_keep);
    for (int yj = 0; yj < yres; yj++) {
        for (int xi = 0; xi < xres; xi++) {
            /* TVec p = TVec(xmin, ymin) + dxdy*TVec(xi, yj), all floats promoted (:88) */
            const double px = (double)xmin + (double)dxdy * (double)xi, py = (double)ymin + (double)dxdy * (double)yj;
            int inbody = 0;
            for (int64_t ib = 0; ib < b->nbody && !inbody; ib++) inbody = vvo_point_invalid(b, ib, px, py) >= 0;
            if (inbody) { out[(size_t)yj * xres + xi] = 0; continue; }
            double cp = 0;
            for (int64_t ib = 0; ib < b->nbody; ib++) {
                const double* bp = b->bprop + 13 * ib;
                const double ax = bp[0], ay = bp[1], sx = bp[10], sy = bp[11], so = bp[12];
                for (int64_t s = b->bfirst[ib]; s < b->bfirst[ib + 1]; s++) {   /* first addend, :115-121 */
                    const double ux = b->rx[s] - ax, uy = b->ry[s] - ay;
                    const double vsx = sx + so * (-uy), vsy = sy + so * ux;        /* Vs = speed.r + speed.o * rotl(r - axis) */
                    const double g = -(vsx * b->dlx[s] + vsy * b->dly[s]);
                    const double q = -(-vsy * b->dlx[s] + vsx * b->dly[s]);
                    const double drx = px - b->rx[s], dry = py - b->ry[s], r2 = drx * drx + dry * dry;
                    const double kx = drx / r2, ky = dry / r2;
                    cp += (-ky * g + kx * q) * vsx + (kx * g + ky * q) * vsy;
                }
                double gtmp = 0;
                for (int64_t s = b->bfirst[ib]; s < b->bfirst[ib + 1]; s++) {   /* second addend, :124-130 */
                    gtmp += b->gsum[s];
                    const double drx = px - b->rx[s], dry = py - b->ry[s], r2 = drx * drx + dry * dry;
                    const double kx = drx / r2, ky = dry / r2;
                    cp -= (b->dlx[s] / dt * (-ky) + b->dly[s] / dt * kx) * gtmp;
                }
            }
            for (int64_t j = 0; j < p->n; j++) {   /* :133-136 */
                const double drx = px - p->x[j], dry = py - p->y[j], r2 = drx * drx + dry * dry;
                const double kx = drx / r2, ky = dry / r2;
                cp += (p->vx[j] * (-ky) + p->vy[j] * kx) * p->g[j];
            }
            double v[2], xy[2] = {px, py};
            vvo_velocity_at(t, p, b, inf_vx, inf_vy, dt, sinks, nsink, xy, 1, v);
            double res = C_1_2PI * cp + 0.5 * ((inf_vx * inf_vx + inf_vy * inf_vy) - (v[0] * v[0] + v[1] * v[1]));
            if (use_ref_speed) res += 0.5 * (sqr(v[0] - ref_vx) + sqr(v[1] - ref_vy));
            out[(size_t)yj * xres + xi] = res;
        }
    }
    vvo_tree_free(t);
}

/* ------------------------------------------------------------------ diffusive */

/* MDiffusiveFast::process_vort_list, MDiffusiveFast.cpp:8-48, with vortex_influence (:93-105)
 * and segment_influence (:107-123) */
void vvo_diffusive(const vvo_tree* t, vvo_plist* p, vvo_bodies* b, double re) {
    LEAF_LOOP(l, t) {
        int64_t node = t->leaf_node[l];
        for (int64_t i = t->vfirst[node]; i < t->vlast[node]; i++) {
            if (!p->g[i]) continue;
            double S2x = 0, S2y = 0, S3x = 0, S3y = 0, S1 = 0, S0 = 0;
            double px = p->x[i], py = p->y[i], ie = p->ieps[i], gi = p->g[i];
            for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) {
                int64_t nn = t->leaf_node[t->near_idx[k]];
                for (int64_t j = t->vfirst[nn]; j < t->vlast[nn]; j++) {
                    if (!p->g[j]) continue;
                    if (sgn(gi) != sgn(p->g[j])) continue;
                    double dx = px - p->x[j], dy = py - p->y[j];
                    if (fabs(dx) + fabs(dy) < 1E-10) continue;
                    double drabs = sqrt(dx * dx + dy * dy);
                    double exparg = -drabs * ie;
                    if (exparg < -8.) continue;
                    double i1tmp = p->g[j] * exp(exparg);
                    double q = i1tmp / drabs;
                    S2x += dx * q; S2y += dy * q;
                    S1 += i1tmp;
                }
                if (b) for (int64_t q = t->sfirst[nn]; q < t->slast[nn]; q++) {
                    int64_t s = t->seg_perm[q];
                    double dx = px - b->rx[s], dy = py - b->ry[s];
                    double drabs2 = dx * dx + dy * dy;
                    double drabs = sqrt(drabs2);
                    double exparg = -drabs * ie;
                    if (exparg < -8.) continue;
                    double expres = exp(exparg);
                    double dSx = -b->dly[s], dSy = b->dlx[s];
                    S3x += dSx * expres; S3y += dSy * expres;
                    S0 += (drabs * ie + 1) / drabs2 * (dx * dSx + dy * dSy) * expres;
                    b->fric[s] += sqr(ie) * gi * expres * sqrt(dSx * dSx + dSy * dSy);
                }
            }
            if ((sgn(S1) != sgn(gi)) || (fabs(S1) < fabs(0.1 * gi))) S1 = 0.1 * gi;
            double k2 = ie / (re * S1);
            p->vx[i] += k2 * S2x; p->vy[i] += k2 * S2y;
            if (S0 > PI) S0 = PI;
            double k3 = sqr(ie) / (re * (C_2PI - S0));
            p->vx[i] += k3 * S3x; p->vy[i] += k3 * S3y;
        }
    }
}

/* ------------------------------------------------------------------ flowmove */

/* TBody::isPointInvalid -> isPointInContour, TBody.cpp:217-220,238-281 */
int64_t vvo_point_invalid(const vvo_bodies* b, int64_t ib, double px, double py) {
    const double* bp = b->bprop + 13 * ib;
    int in = bp[9] != 0;
    if (!in && (px < bp[4] || py < bp[5] || px > bp[6] || py > bp[7] ||
                sqr(px - bp[2]) + sqr(py - bp[3]) > bp[8])) return -1;
    int64_t f = b->bfirst[ib], e = b->bfirst[ib + 1];
    for (int64_t i = f, j = e - 1; i < e; j = i++) {
        double vix = b->cx[i], viy = b->cy[i], vjx = b->cx[j], vjy = b->cy[j];
        if (((viy < vjy) && (viy < py) && (py <= vjy) && ((vjy - viy) * (px - vix) > (vjx - vix) * (py - viy))) ||
            ((viy > vjy) && (viy > py) && (py >= vjy) && ((vjy - viy) * (px - vix) < (vjx - vix) * (py - viy))))
            in = !in;
    }
    if (!in) return -1;
    int64_t nearest = -1;
    double nd = DBL_MAX;
    for (int64_t s = f; s < e; s++) {
        double d = sqr(b->rx[s] - px) + sqr(b->ry[s] - py);
        if (d < nd) { nearest = s; nd = d; }
    }
    return nearest;
}

/* MFlowmove::move_and_clean, MFlowmove.cpp:107-144,194-199 (vortex list; body kinematics,
 * collision detection and the attached-vortex bookkeeping of :201-214 stay with the caller) */
int64_t vvo_move_and_clean(vvo_plist* p, vvo_bodies* b, double dt, double remove_eps, int remove, int64_t* cleaned) {
    int64_t n = p->n, w = 0, nclean = 0;
    for (int64_t i = 0; i < n; i++) {
        p->x[i] += p->vx[i] * dt;
        p->y[i] += p->vy[i] * dt;
    }
    for (int64_t i = 0; i < n; i++) {
        if (fabs(p->g[i]) < remove_eps) continue;
        if (remove && b) {
            int64_t bad = -1, bb = -1;
            for (int64_t ib = 0; ib < b->nbody; ib++) {
                bad = vvo_point_invalid(b, ib, p->x[i], p->y[i]);
                if (bad >= 0) { bb = ib; break; }
            }
            if (bb >= 0) {
                const double* bp = b->bprop + 13 * bb;
                b->fdt_dead[3 * bb + 0] += -p->y[i] * p->g[i];
                b->fdt_dead[3 * bb + 1] += p->x[i] * p->g[i];
                b->fdt_dead[3 * bb + 2] += (sqr(p->x[i] - bp[0]) + sqr(p->y[i] - bp[1])) * p->g[i];
                b->gsum[bad] -= p->g[i];
                b->g_dead[bb] += p->g[i];
                nclean++;
                continue;
            }
        }
        p->x[w] = p->x[i]; p->y[w] = p->y[i]; p->g[w] = p->g[i];
        p->ieps[w] = p->ieps[i]; p->orig[w] = p->orig[i];
        p->vx[w] = 0; p->vy[w] = 0;
        w++;
    }
    p->n = w;
    if (cleaned) *cleaned = nclean;
    return w;
}

void vvo_count_interactions(const vvo_tree* t, const vvo_plist* p, int64_t l0, int64_t l1, double* near_pairs,
                            double* far_nodes) {
    double np = 0, nf = 0;
    int64_t* nz = malloc(sizeof(int64_t) * (size_t)(t->n_leaves + 1));
    for (int64_t l = 0; l < t->n_leaves; l++) {
        int64_t node = t->leaf_node[l], k = 0;
        for (int64_t i = t->vfirst[node]; i < t->vlast[node]; i++) k += (p->g[i] != 0);
        nz[l] = k;
    }
    for (int64_t l = l0; l < l1; l++) {
        double s = 0;
        for (int64_t k = t->near_ptr[l]; k < t->near_ptr[l + 1]; k++) s += (double)nz[t->near_idx[k]];
        np += (double)nz[l] * s;
        nf += (double)(t->far_ptr[l + 1] - t->far_ptr[l]);
    }
    free(nz);
    *near_pairs = np; *far_nodes = nf;
}
