/* avg_oracle.c — float64 CPU restatement of the reference hot path (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in the third-party `pybullet` C extension
 * (reference setup.py:18, unpinned; in practice a custom bullet3 fork — SURVEY.md §0 F2), which is absent from
 * /root/reference and not installable here, and the reference ships no tests, golden vectors or fixtures
 * (SURVEY.md §4, §8c).  What is restated below therefore has two kinds of source:
 *   (1) code that IS in the reference tree and is followed line by line (cited per function):
 *       AssistiveEnv.take_step            env.py:274-351
 *       enforce_realistic_human_joint_limits env.py:353-387 (right arm; the left-arm block never fires in sim ids)
 *       enforce_hard_human_joint_limits   env.py:389-410
 *       human_preferences                 env.py:412-448
 *       ScratchItchEnv.step/reward        scratch_itch.py:30-82
 *       ScratchItchEnv.get_total_force    scratch_itch.py:84-102
 *       ScratchItchEnv._get_obs           scratch_itch.py:104-128
 *       ScratchItchEnv.update_targets     scratch_itch.py:289-293
 *   (2) p.stepSimulation / setJointMotorControlArray / createConstraint / getContactPoints, restated from the
 *       published algorithms Bullet implements (Featherstone articulated-body algorithm; GJK closest points;
 *       projected Gauss-Seidel sequential impulses) with the Bullet semantics listed in SURVEY.md App. D.  Every
 *       such choice is tagged [UPSTREAM-BULLET] and remains unverified until a PyBullet capture exists.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * It deliberately shares no code with the CUDA product: forward dynamics here is the recursive O(n) ABA and the
 * solver iterates in velocity space like Bullet, whereas the kernels use a mass-matrix / Delassus formulation.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/avg_model.h"

#define MAXB AVG_MAX_BODY
#define MAXD AVG_MAX_DOF
#define MAXC AVG_MAX_CONTACT
#define MAXR AVG_MAX_ROWS

typedef struct { double x, y, z; } v3;
typedef struct { double m[3][3]; } m3;
typedef struct { double x, y, z, w; } quat;

static v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vscale(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
static double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static double vnorm(v3 a) { return sqrt(vdot(a, a)); }
static v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static v3 f3(const float* f) { return V(f[0], f[1], f[2]); }
static quat qf(const float* f) { quat q = {f[0], f[1], f[2], f[3]}; return q; }
static quat qmul(quat a, quat b) {
    quat r = {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
              a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
    return r;
}
static quat qnormalize(quat q) {
    double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    quat r = {q.x / n, q.y / n, q.z / n, q.w / n};
    return r;
}
static quat qconj(quat q) { quat r = {-q.x, -q.y, -q.z, q.w}; return r; }
static m3 qmat(quat q) {
    m3 r; double x = q.x, y = q.y, z = q.z, w = q.w;
    r.m[0][0] = 1 - 2 * (y * y + z * z); r.m[0][1] = 2 * (x * y - z * w); r.m[0][2] = 2 * (x * z + y * w);
    r.m[1][0] = 2 * (x * y + z * w); r.m[1][1] = 1 - 2 * (x * x + z * z); r.m[1][2] = 2 * (y * z - x * w);
    r.m[2][0] = 2 * (x * z - y * w); r.m[2][1] = 2 * (y * z + x * w); r.m[2][2] = 1 - 2 * (x * x + y * y);
    return r;
}
static v3 mmulv(const m3* a, v3 v) {
    return V(a->m[0][0] * v.x + a->m[0][1] * v.y + a->m[0][2] * v.z, a->m[1][0] * v.x + a->m[1][1] * v.y + a->m[1][2] * v.z,
             a->m[2][0] * v.x + a->m[2][1] * v.y + a->m[2][2] * v.z);
}
static v3 mtmulv(const m3* a, v3 v) {
    return V(a->m[0][0] * v.x + a->m[1][0] * v.y + a->m[2][0] * v.z, a->m[0][1] * v.x + a->m[1][1] * v.y + a->m[2][1] * v.z,
             a->m[0][2] * v.x + a->m[1][2] * v.y + a->m[2][2] * v.z);
}
static v3 qrot(quat q, v3 v) { m3 r = qmat(q); return mmulv(&r, v); }
static quat qaxis(v3 axis, double ang) {
    double s = sin(ang * 0.5); quat r = {axis.x * s, axis.y * s, axis.z * s, cos(ang * 0.5)}; return r;
}

/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const AvgModelHeader* h;
    const AvgBody* body;
    const AvgDof* dof;
    const AvgShape* shape;
    const float* vert;
    const float* plane;
    const uint32_t* pair;
    const AvgFrame* frame;
    const float* mlp;
    const float* target;     /* BedBathing wiping targets, float4 per target */
} Model;

static int model_open(const void* blob, Model* m) {
    const AvgModelHeader* h = (const AvgModelHeader*)blob;
    if (h->magic != AVG_MAGIC || h->version != AVG_VERSION) return -1;
    if (h->n_body + h->n_ebody > MAXB || h->n_dof > MAXD || h->n_ebody > AVG_MAX_EBODY || h->n_particle > AVG_MAX_PARTICLE) return -2;
    const char* b = (const char*)blob;
    m->h = h;
    m->body = (const AvgBody*)(b + h->off_body);
    m->dof = (const AvgDof*)(b + h->off_dof);
    m->shape = (const AvgShape*)(b + h->off_shape);
    m->vert = (const float*)(b + h->off_vert);
    m->plane = (const float*)(b + h->off_plane);
    m->pair = (const uint32_t*)(b + h->off_pair);
    m->frame = (const AvgFrame*)(b + h->off_frame);
    m->mlp = h->n_mlp > 0 ? (const float*)(b + h->off_mlp) : 0;
    m->target = h->n_target > 0 ? (const float*)(b + h->off_target) : 0;
    return 0;
}

/* per-sub-step kinematic state */
typedef struct {
    v3 p[MAXB];      /* body frame origin (composite COM), world */
    quat q[MAXB];
    m3 R[MAXB];
    v3 axis[MAXB];   /* joint axis, world */
    v3 org[MAXB];    /* joint origin, world */
} Kin;

static void body_pose(const Kin* k, int body, v3* p, quat* q) {
    if (body < 0) { *p = V(0, 0, 0); quat id = {0, 0, 0, 1}; *q = id; }
    else { *p = k->p[body]; *q = k->q[body]; }
}

/* forward kinematics: C_body = C_parent ∘ Ta ∘ joint(q) ∘ Tb  (include/avg_model.h AvgBody) */
static void fk(const Model* m, const double* env, Kin* k) {
    /* env-static bodies (the Feeding bowl, feeding.py:184-185): pose straight from the record, slots n_body + e */
    for (int e = 0; e < m->h->n_ebody; ++e) {
        int b = m->h->n_body + e;
        const double* q = env + AVG_E_EBODY + 7 * e;
        k->p[b] = V(q[0], q[1], q[2]);
        quat r = {q[3], q[4], q[5], q[6]};
        k->q[b] = qnormalize(r); k->R[b] = qmat(k->q[b]);
        k->axis[b] = V(0, 0, 0); k->org[b] = k->p[b];
    }
    for (int b = 0; b < m->h->n_body; ++b) {
        const AvgBody* B = &m->body[b];
        if (B->jtype == AVG_JOINT_FREE) {
            const double* q = env + AVG_E_Q + B->qidx;
            k->p[b] = V(q[0], q[1], q[2]);
            quat r = {q[3], q[4], q[5], q[6]};
            k->q[b] = qnormalize(r);
            k->R[b] = qmat(k->q[b]);
            k->axis[b] = V(0, 0, 0); k->org[b] = k->p[b];
            continue;
        }
        v3 pp; quat pq;
        body_pose(k, B->parent, &pp, &pq);
        v3 jp = vadd(pp, qrot(pq, f3(B->ta_pos)));
        quat jq = qmul(pq, qf(B->ta_quat));
        v3 ax = f3(B->axis);
        double qv = env[AVG_E_Q + B->qidx];
        k->axis[b] = qrot(jq, ax);
        k->org[b] = jp;
        if (B->jtype == AVG_JOINT_REVOLUTE) jq = qmul(jq, qaxis(ax, qv));
        else jp = vadd(jp, vscale(k->axis[b], qv));
        k->p[b] = vadd(jp, qrot(jq, f3(B->tb_pos)));
        k->q[b] = qnormalize(qmul(jq, qf(B->tb_quat)));
        k->R[b] = qmat(k->q[b]);
    }
}

static void frame_pose(const Model* m, const Kin* k, int f, v3* p, quat* q) {
    const AvgFrame* F = &m->frame[f];
    v3 bp; quat bq;
    body_pose(k, F->body, &bp, &bq);
    *p = vadd(bp, qrot(bq, f3(F->pos)));
    *q = qnormalize(qmul(bq, qf(F->quat)));
}

/* ---------------------------------------------------------------------------------------------------------------
 * Collision: support-function shapes, GJK closest points on the cores, rounding margins added afterwards.
 * [UPSTREAM-BULLET] Bullet runs GJK on margin-shrunk cores and reports distance minus both margins; one new point
 * per shape pair per step; a point is kept when distance < the manifold's contact-breaking threshold
 * (min over the two collision objects of 0.02 x angular-motion-disc).  Persistent 4-point manifolds and
 * warm starting are NOT restated (documented deviation).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
    const AvgShape* s;
    v3 p; m3 R;
    const float* verts;
    const float* planes;
} WShape;

static void shape_world(const Model* m, const Kin* k, int si, WShape* w) {
    const AvgShape* s = &m->shape[si];
    w->s = s; w->verts = m->vert + 4 * s->vert_off; w->planes = m->plane + 4 * s->plane_off;
    if (s->body < 0) { w->p = f3(s->pos); w->R = qmat(qf(s->quat)); }
    else {
        w->p = vadd(k->p[s->body], qrot(k->q[s->body], f3(s->pos)));
        w->R = qmat(qnormalize(qmul(k->q[s->body], qf(s->quat))));
    }
}

static void shape_aabb(const WShape* w, v3* c, v3* h) {
    const AvgShape* s = w->s;
    if (s->body < 0) { *c = f3(s->aabb_c); *h = f3(s->aabb_h); return; }
    v3 lc = f3(s->aabb_c), lh = f3(s->aabb_h);
    *c = vadd(w->p, mmulv(&w->R, lc));
    h->x = fabs(w->R.m[0][0]) * lh.x + fabs(w->R.m[0][1]) * lh.y + fabs(w->R.m[0][2]) * lh.z;
    h->y = fabs(w->R.m[1][0]) * lh.x + fabs(w->R.m[1][1]) * lh.y + fabs(w->R.m[1][2]) * lh.z;
    h->z = fabs(w->R.m[2][0]) * lh.x + fabs(w->R.m[2][1]) * lh.y + fabs(w->R.m[2][2]) * lh.z;
}

/* support point of the CORE of shape w in world direction d */
static v3 support(const WShape* w, v3 d) {
    const AvgShape* s = w->s;
    v3 l = mtmulv(&w->R, d), r;
    switch (s->type) {
    case AVG_SHAPE_SPHERE: r = V(0, 0, 0); break;
    case AVG_SHAPE_CAPSULE: r = V(0, 0, l.z > 1e-9 ? s->half[2] : (l.z < -1e-9 ? -s->half[2] : 0.0)); break;
    case AVG_SHAPE_BOX: {
        double hx = s->half[0] - s->margin, hy = s->half[1] - s->margin, hz = s->half[2] - s->margin;
        r = V(l.x >= 0 ? hx : -hx, l.y >= 0 ? hy : -hy, l.z >= 0 ? hz : -hz); break;
    }
    case AVG_SHAPE_CYLINDER: {
        double rc = s->radius - s->margin, hc = s->half[2] - s->margin;
        double n = sqrt(l.x * l.x + l.y * l.y);
        if (n > 1e-12) r = V(rc * l.x / n, rc * l.y / n, l.z >= 0 ? hc : -hc);
        else r = V(rc, 0, l.z >= 0 ? hc : -hc);
        break;
    }
    case AVG_SHAPE_HULL: {
        int best = 0; double bd = -1e300;
        for (int i = 0; i < s->vert_cnt; ++i) {
            double dd = l.x * w->verts[4 * i] + l.y * w->verts[4 * i + 1] + l.z * w->verts[4 * i + 2];
            if (dd > bd) { bd = dd; best = i; }
        }
        r = V(w->verts[4 * best], w->verts[4 * best + 1], w->verts[4 * best + 2]); break;
    }
    default: r = V(0, 0, 0);
    }
    return vadd(w->p, mmulv(&w->R, r));
}

/* Supporting FEATURE of the core of shape w in world direction d (|d| = 1): centroid of every core point whose support
 * value lies within FEAT_TOL of the maximum, and the feature's size class (1 vertex, 2 edge, 4 or more face).  A
 * polytope's support point along one of its own face normals is not unique -- which vertex of the face wins the arg-max is
 * decided by rounding noise -- so penetration witnesses are taken from the feature centroid, which float32 and float64
 * agree on. */
#define FEAT_TOL 1e-4
static v3 support_feature(const WShape* w, v3 d, int* count) {
    const AvgShape* s = w->s;
    v3 l = mtmulv(&w->R, d), r;
    int cnt = 1;
    switch (s->type) {
    case AVG_SHAPE_CAPSULE:
        if (fabs(l.z) <= FEAT_TOL) { r = V(0, 0, 0); cnt = 2; } else r = V(0, 0, l.z > 0 ? s->half[2] : -s->half[2]);
        break;
    case AVG_SHAPE_BOX: {
        double hh[3] = {s->half[0] - s->margin, s->half[1] - s->margin, s->half[2] - s->margin}, ll[3] = {l.x, l.y, l.z}, rr[3];
        for (int k = 0; k < 3; ++k) { if (fabs(ll[k]) <= FEAT_TOL) { rr[k] = 0; cnt *= 2; } else rr[k] = ll[k] > 0 ? hh[k] : -hh[k]; }
        r = V(rr[0], rr[1], rr[2]); break;
    }
    case AVG_SHAPE_CYLINDER: {
        double rc = s->radius - s->margin, hc = s->half[2] - s->margin;
        double n = sqrt(l.x * l.x + l.y * l.y);
        double z = 0; if (fabs(l.z) <= FEAT_TOL) cnt *= 2; else z = l.z > 0 ? hc : -hc;
        if (n > FEAT_TOL) r = V(rc * l.x / n, rc * l.y / n, z); else { r = V(0, 0, z); cnt *= 8; }
        break;
    }
    case AVG_SHAPE_HULL: {
        double bd = -1e300;
        for (int i = 0; i < s->vert_cnt; ++i) {
            double dd = l.x * w->verts[4 * i] + l.y * w->verts[4 * i + 1] + l.z * w->verts[4 * i + 2];
            if (dd > bd) bd = dd;
        }
        v3 acc = V(0, 0, 0); cnt = 0;
        for (int i = 0; i < s->vert_cnt; ++i) {
            double dd = l.x * w->verts[4 * i] + l.y * w->verts[4 * i + 1] + l.z * w->verts[4 * i + 2];
            if (dd >= bd - FEAT_TOL) { acc = vadd(acc, V(w->verts[4 * i], w->verts[4 * i + 1], w->verts[4 * i + 2])); cnt++; }
        }
        r = vscale(acc, 1.0 / (cnt > 0 ? cnt : 1)); break;
    }
    default: r = V(0, 0, 0);       /* sphere: its centre */
    }
    *count = cnt;
    return vadd(w->p, mmulv(&w->R, r));
}

typedef struct { v3 w[4], a[4], b[4]; double lam[4]; int n; } Simplex;

/* closest point to the origin on the simplex; reduces the simplex to the supporting feature, fills lam.
 * returns 1 if the origin is enclosed (tetrahedron case). */
static void closest_seg(Simplex* s, int i0, int i1, double* lam0, double* lam1, int* keep_mask) {
    v3 a = s->w[i0], b = s->w[i1];
    v3 ab = vsub(b, a);
    double t = -vdot(a, ab), den = vdot(ab, ab);
    if (t <= 0 || den <= 0) { *lam0 = 1; *lam1 = 0; *keep_mask = 1; }
    else if (t >= den) { *lam0 = 0; *lam1 = 1; *keep_mask = 2; }
    else { *lam1 = t / den; *lam0 = 1 - *lam1; *keep_mask = 3; }
}

/* closest point on triangle (i0,i1,i2) to origin (Ericson, Real-Time Collision Detection 5.1.5) */
static void closest_tri(const Simplex* s, int i0, int i1, int i2, double lam[3], int* keep_mask) {
    v3 a = s->w[i0], b = s->w[i1], c = s->w[i2];
    v3 ab = vsub(b, a), ac = vsub(c, a), ap = vneg(a);
    double d1 = vdot(ab, ap), d2 = vdot(ac, ap);
    if (d1 <= 0 && d2 <= 0) { lam[0] = 1; lam[1] = 0; lam[2] = 0; *keep_mask = 1; return; }
    v3 bp = vneg(b);
    double d3 = vdot(ab, bp), d4 = vdot(ac, bp);
    if (d3 >= 0 && d4 <= d3) { lam[0] = 0; lam[1] = 1; lam[2] = 0; *keep_mask = 2; return; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); lam[0] = 1 - v; lam[1] = v; lam[2] = 0; *keep_mask = 3; return; }
    v3 cp = vneg(c);
    double d5 = vdot(ab, cp), d6 = vdot(ac, cp);
    if (d6 >= 0 && d5 <= d6) { lam[0] = 0; lam[1] = 0; lam[2] = 1; *keep_mask = 4; return; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); lam[0] = 1 - w; lam[1] = 0; lam[2] = w; *keep_mask = 5; return; }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); lam[0] = 0; lam[1] = 1 - w; lam[2] = w; *keep_mask = 6; return;
    }
    double sum = va + vb + vc;
    if (!(sum > 0.0)) {
        /* zero-area triangle whose vertex / edge tests all fell through (collinear points): closest point of its longest
         * edge, so that no 0/0 reaches the search direction */
        v3 bc = vsub(c, b);
        double lab = vdot(ab, ab), lac = vdot(ac, ac), lbc = vdot(bc, bc);
        int first = (lab >= lac && lab >= lbc) ? 0 : (lac >= lbc ? 1 : 2);      /* ab, ac, bc */
        v3 p0 = first == 2 ? b : a, e = first == 0 ? ab : (first == 1 ? ac : bc);
        int i0 = first == 2 ? 1 : 0, i1 = first == 0 ? 1 : 2;
        double ee = vdot(e, e);
        double t = ee > 0.0 ? fmin(fmax(-vdot(p0, e) / ee, 0.0), 1.0) : 0.0;
        lam[0] = lam[1] = lam[2] = 0.0; lam[i0] = 1.0 - t; lam[i1] += t;
        *keep_mask = (1 << i0) | (t > 0.0 ? (1 << i1) : 0);
        if (t >= 1.0) *keep_mask = 1 << i1;
        return;
    }
    double den = 1.0 / sum;
    lam[1] = vb * den; lam[2] = vc * den; lam[0] = 1 - lam[1] - lam[2]; *keep_mask = 7;
}

static int simplex_closest(Simplex* s, v3* v) {
    double lam[4] = {0, 0, 0, 0};
    int mask = 0;
    if (s->n == 1) { lam[0] = 1; mask = 1; }
    else if (s->n == 2) { closest_seg(s, 0, 1, &lam[0], &lam[1], &mask); }
    else if (s->n == 3) { closest_tri(s, 0, 1, 2, lam, &mask); }
    else {
        /* tetrahedron.  The origin is enclosed only when it lies strictly on the inner side of all four faces of a
         * non-degenerate tetrahedron; otherwise the closest point is the best over ALL four faces (the closest point of a
         * tetrahedron to an outside point lies on its boundary, so looking at a face that turns out to face away costs
         * time, never correctness).  Culling faces by the sign of the apex side alone is not robust: the difference of a
         * hull and a capsule core contains parallelograms (h_i - c_j with two hull vertices and the two segment ends), so
         * four simplex points are routinely coplanar, the apex side is rounding noise, and a culled face could be the
         * closest one (found with the PR2 forearm hull against the human forearm capsule: 1.2 cm of distance error). */
        static const int F[4][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 3, 1}, {1, 2, 3, 0}};
        double best = 1e300; int inside = 1, degenerate = 0;
        for (int f = 0; f < 4; ++f) {
            v3 a = s->w[F[f][0]], b = s->w[F[f][1]], c = s->w[F[f][2]], d = s->w[F[f][3]];
            v3 n = vcross(vsub(b, a), vsub(c, a));
            double so = -vdot(a, n), sd = vdot(vsub(d, a), n);
            if (fabs(sd) <= 1e-9 * vnorm(n) * vnorm(vsub(d, a))) degenerate = 1;
            if (!(so * sd > 0)) inside = 0;
            double l3[3]; int km;
            closest_tri(s, F[f][0], F[f][1], F[f][2], l3, &km);
            v3 p = vadd(vadd(vscale(a, l3[0]), vscale(b, l3[1])), vscale(c, l3[2]));
            double dd = vdot(p, p);
            if (dd < best) {
                best = dd;
                lam[0] = lam[1] = lam[2] = lam[3] = 0;
                lam[F[f][0]] = l3[0]; lam[F[f][1]] = l3[1]; lam[F[f][2]] = l3[2];
                mask = ((km & 1) ? (1 << F[f][0]) : 0) | ((km & 2) ? (1 << F[f][1]) : 0) | ((km & 4) ? (1 << F[f][2]) : 0);
            }
        }
        if (inside && !degenerate) return 1;   /* origin inside */
    }
    /* compact */
    Simplex r; r.n = 0;
    v3 p = V(0, 0, 0);
    for (int i = 0; i < s->n; ++i) if (mask & (1 << i)) {
        r.w[r.n] = s->w[i]; r.a[r.n] = s->a[i]; r.b[r.n] = s->b[i]; r.lam[r.n] = lam[i];
        p = vadd(p, vscale(s->w[i], lam[i]));
        r.n++;
    }
    *s = r; *v = p;
    return 0;
}

/* returns 0 = separated cores (dist, pa, pb valid), 1 = cores overlap */
static int gjk(const WShape* A, const WShape* B, double* dist, v3* pa, v3* pb) {
    Simplex s; s.n = 0;
    v3 v = vsub(A->p, B->p);
    if (vdot(v, v) < 1e-12) v = V(1, 0, 0);
    for (int it = 0; it < 48; ++it) {
        v3 sa = support(A, vneg(v)), sb = support(B, v);
        v3 w = vsub(sa, sb);
        double vv = vdot(v, v), vw = vdot(v, w);
        if (s.n > 0 && (vv - vw) <= 1e-10 * vv + 1e-18) break;      /* no progress possible: v is the closest point */
        int dup = 0;
        for (int i = 0; i < s.n; ++i) { v3 d = vsub(s.w[i], w); if (vdot(d, d) < 1e-24) dup = 1; }
        if (dup) break;
        s.w[s.n] = w; s.a[s.n] = sa; s.b[s.n] = sb; s.n++;
        if (simplex_closest(&s, &v)) return 1;
        if (vdot(v, v) < 1e-20) return 1;
    }
    v3 a = V(0, 0, 0), b = V(0, 0, 0);
    for (int i = 0; i < s.n; ++i) { a = vadd(a, vscale(s.a[i], s.lam[i])); b = vadd(b, vscale(s.b[i], s.lam[i])); }
    *pa = a; *pb = b; *dist = vnorm(v);
    return 0;
}

/* deep-penetration fallback: smallest overlap among face-normal / centre axes evaluated with support functions */
static int is_round(const WShape* S) { return S->s->type == AVG_SHAPE_SPHERE || S->s->type == AVG_SHAPE_CAPSULE; }
/* bpa = core witness on A, see support_feature(). */
static void sat_axis(const WShape* A, const WShape* B, v3 n, double* best, v3* bn, v3* bpa) {
    double ln = vnorm(n);
    if (ln < 1e-12) return;
    n = vscale(n, 1.0 / ln);
    v3 sa = support(A, vneg(n)), sb = support(B, n);
    double depth = vdot(sb, n) - vdot(sa, n);
    if (depth < *best) {
        *best = depth; *bn = n;
        /* witness on A's core: from the round shape when there is one, otherwise from the smaller supporting feature
         * (a vertex pressing into a face marks the contact, the face's centroid does not) */
        int ca, cb;
        v3 fa = support_feature(A, vneg(n), &ca), fb = support_feature(B, n, &cb);
        if (is_round(B) && !is_round(A)) *bpa = vsub(fb, vscale(n, depth));
        else if (is_round(A)) *bpa = fa;
        else *bpa = (ca <= cb) ? fa : vsub(fb, vscale(n, depth));
    }
}
static void shape_axes(const WShape* S, const WShape* O, double sign, const WShape* A, const WShape* B, double* best, v3* bn, v3* bpa) {
    const AvgShape* s = S->s;
    /* `sign` = +1 when S is B (its outward normals point from B to A), -1 when S is A */
    if (s->type == AVG_SHAPE_BOX || s->type == AVG_SHAPE_CYLINDER) {
        for (int ax = 0; ax < 3; ++ax) {
            if (s->type == AVG_SHAPE_CYLINDER && ax < 2) continue;
            v3 n = V(S->R.m[0][ax], S->R.m[1][ax], S->R.m[2][ax]);
            sat_axis(A, B, vscale(n, sign), best, bn, bpa);
            sat_axis(A, B, vscale(n, -sign), best, bn, bpa);
        }
        if (s->type == AVG_SHAPE_CYLINDER) {
            v3 az = V(S->R.m[0][2], S->R.m[1][2], S->R.m[2][2]);
            v3 d = vsub(O->p, S->p);
            v3 rad = vsub(d, vscale(az, vdot(d, az)));
            sat_axis(A, B, vscale(rad, sign), best, bn, bpa);
        }
    } else if (s->type == AVG_SHAPE_HULL) {
        for (int i = 0; i < s->plane_cnt; ++i) {
            v3 n = mmulv(&S->R, V(S->planes[4 * i], S->planes[4 * i + 1], S->planes[4 * i + 2]));
            sat_axis(A, B, vscale(n, sign), best, bn, bpa);
        }
    }
}
static void deep_contact(const WShape* A, const WShape* B, double* depth, v3* n, v3* pa_core) {
    double best = 1e300; v3 bn = V(0, 0, 1), bpa = A->p;
    shape_axes(A, B, -1.0, A, B, &best, &bn, &bpa);
    shape_axes(B, A, +1.0, A, B, &best, &bn, &bpa);
    v3 cc = vsub(A->p, B->p);
    sat_axis(A, B, cc, &best, &bn, &bpa);
    if (best > 1e299) { best = 0; }
    *depth = best; *n = bn; *pa_core = bpa;
}

typedef struct {
    int sa, sb;
    v3 pa, pb, n;     /* world; n from B to A */
    double dist;
    double lambda_n;  /* normal impulse after the solve */
} Contact;

static int narrowphase(const WShape* A, const WShape* B, double thr, Contact* c) {
    double ma = A->s->margin, mb = B->s->margin;
    if (B->s->type == AVG_SHAPE_PLANE) {            /* static ground: half-space z <= 0 */
        v3 s = support(A, V(0, 0, -1));
        double d = s.z - ma;
        if (d >= thr) return 0;
        c->n = V(0, 0, 1); c->pa = V(s.x, s.y, s.z - ma); c->pb = V(s.x, s.y, 0); c->dist = d;
        return 1;
    }
    double dist; v3 pa, pb;
    if (!gjk(A, B, &dist, &pa, &pb)) {
        double d = dist - ma - mb;
        if (d >= thr) return 0;
        v3 n = vscale(vsub(pa, pb), 1.0 / dist);
        c->n = n; c->pa = vsub(pa, vscale(n, ma)); c->pb = vadd(pb, vscale(n, mb)); c->dist = d;
        return 1;
    }
    double depth; v3 n, pac;
    deep_contact(A, B, &depth, &n, &pac);
    c->n = n; c->dist = -depth - ma - mb;
    c->pa = vsub(pac, vscale(n, ma));
    c->pb = vadd(vadd(pac, vscale(n, depth)), vscale(n, mb));
    return 1;
}

static int n_shapes_all(const Model* m) { return m->h->n_shape + m->h->n_cshape + (m->h->n_particle > 0 ? 1 : 0); }

static int collide(const Model* m, const Kin* k, Contact* out, int* overflow) {
    /* the pair table lists the convex children of compound shapes explicitly (compiler/scene.py expand_compounds) */
    int ns = n_shapes_all(m), nc = 0;
    WShape* ws = (WShape*)malloc(sizeof(WShape) * ns);
    v3* ac = (v3*)malloc(sizeof(v3) * ns); v3* ah = (v3*)malloc(sizeof(v3) * ns);
    for (int i = 0; i < ns; ++i) { shape_world(m, k, i, &ws[i]); shape_aabb(&ws[i], &ac[i], &ah[i]); }
    for (int pi = 0; pi < m->h->n_pair; ++pi) {
        int a = m->pair[pi] & 0xffff, b = m->pair[pi] >> 16;
        double thr = fmin(ws[a].s->thr, ws[b].s->thr);
        if (ws[b].s->type != AVG_SHAPE_PLANE) {
            if (fabs(ac[a].x - ac[b].x) > ah[a].x + ah[b].x + thr) continue;
            if (fabs(ac[a].y - ac[b].y) > ah[a].y + ah[b].y + thr) continue;
            if (fabs(ac[a].z - ac[b].z) > ah[a].z + ah[b].z + thr) continue;
        } else if (ac[a].z - ah[a].z > thr) continue;
        Contact c;
        if (narrowphase(&ws[a], &ws[b], thr, &c)) {
            if (nc >= MAXC) { *overflow |= 1; break; }
            c.sa = a; c.sb = b; c.lambda_n = 0;
            out[nc++] = c;
        }
    }
    free(ws); free(ac); free(ah);
    return nc;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Dynamics: articulated-body algorithm with all spatial quantities expressed in world coordinates about the world
 * origin (motion [w; v_O], force [n_O; f]).  [UPSTREAM-BULLET] btMultiBody::computeAccelerationsArticulatedBody
 * AlgorithmMultiDof, including its velocity damping: every link gets the extra bias force
 * m v (k + k |v|) and I w (k + k |w|) with k = linear/angular damping 0.04, and the gyroscopic term.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct { double v[6]; } sv;            /* spatial vector */
typedef struct { double m[6][6]; } sm;         /* spatial matrix */

static sv sv_make(v3 a, v3 l) { sv r = {{a.x, a.y, a.z, l.x, l.y, l.z}}; return r; }
static v3 sv_ang(const sv* s) { return V(s->v[0], s->v[1], s->v[2]); }
static v3 sv_lin(const sv* s) { return V(s->v[3], s->v[4], s->v[5]); }
static double sv_dot(const sv* a, const sv* b) { double r = 0; for (int i = 0; i < 6; ++i) r += a->v[i] * b->v[i]; return r; }
static sv sm_mul(const sm* A, const sv* x) {
    sv r; for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += A->m[i][j] * x->v[j]; r.v[i] = s; } return r;
}
/* motion cross motion: [w;v] x [w2;v2] = [w x w2 ; w x v2 + v x w2] */
static sv crm(const sv* a, const sv* b) {
    v3 w = sv_ang(a), v = sv_lin(a), w2 = sv_ang(b), v2 = sv_lin(b);
    return sv_make(vcross(w, w2), vadd(vcross(w, v2), vcross(v, w2)));
}
/* motion cross force: [w;v] x* [n;f] = [w x n + v x f ; w x f] */
static sv crf(const sv* a, const sv* f) {
    v3 w = sv_ang(a), v = sv_lin(a), n = sv_ang(f), ff = sv_lin(f);
    return sv_make(vadd(vcross(w, n), vcross(v, ff)), vcross(w, ff));
}
static void skew(v3 c, double S[3][3]) {
    S[0][0] = 0; S[0][1] = -c.z; S[0][2] = c.y; S[1][0] = c.z; S[1][1] = 0; S[1][2] = -c.x; S[2][0] = -c.y; S[2][1] = c.x; S[2][2] = 0;
}
/* spatial inertia about the world origin of a body with mass m, COM c, rotational inertia Ic (world axes) */
static void spatial_inertia(double mass, v3 c, const double Ic[3][3], sm* I) {
    double C[3][3]; skew(c, C);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double cct = 0; for (int k = 0; k < 3; ++k) cct += C[i][k] * C[j][k];      /* C C^T */
        I->m[i][j] = Ic[i][j] + mass * cct;
        I->m[i][j + 3] = mass * C[i][j];
        I->m[i + 3][j] = mass * C[j][i];
        I->m[i + 3][j + 3] = (i == j) ? mass : 0.0;
    }
}

typedef struct {
    sv S[MAXB];        /* joint motion subspace (world) */
    sv vel[MAXB];      /* body spatial velocity */
    sv cb[MAXB];       /* velocity-product acceleration */
    sm IA[MAXB];       /* articulated inertia */
    sv U[MAXB];
    double D[MAXB], invD[MAXB];
    double Iw[MAXB][3][3];   /* rotational inertia about COM, world axes */
    double IwInv[MAXB][3][3];
} Dyn;

/* changeDynamics(mass=0) per episode (world_creation.py:157-161): bodies in the record's AVG_E_FROZEN mask count as massless */
static int body_frozen(const double* env, int b) { return (((uint32_t)env[AVG_E_FROZEN]) >> b) & 1u; }

static void world_inertia(const Model* m, const Kin* k, const double* env, Dyn* d) {
    for (int b = 0; b < m->h->n_body; ++b) {
        const AvgBody* B = &m->body[b];
        double ine[3] = {B->inertia[0], B->inertia[1], B->inertia[2]};
        if (body_frozen(env, b)) ine[0] = ine[1] = ine[2] = 0.0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double s = 0, si = 0;
            for (int a = 0; a < 3; ++a) {
                s += k->R[b].m[i][a] * ine[a] * k->R[b].m[j][a];
                si += (ine[a] > 0) ? k->R[b].m[i][a] / ine[a] * k->R[b].m[j][a] : 0.0;
            }
            d->Iw[b][i][j] = s; d->IwInv[b][i][j] = si;
        }
    }
}

/* unconstrained accelerations; also caches IA, U, D for the impulse-response solves */
static void aba(const Model* m, const Kin* k, const double* env, Dyn* d, double* qdd) {
    const AvgModelHeader* h = m->h;
    int nb = h->n_body;
    sv pA[MAXB];
    world_inertia(m, k, env, d);
    /* pass 1 */
    for (int b = 0; b < nb; ++b) {
        const AvgBody* B = &m->body[b];
        if (B->jtype == AVG_JOINT_FREE) continue;
        const double bmass = body_frozen(env, b) ? 0.0 : B->mass;
        double qd = env[AVG_E_QD + B->dof];
        if (B->jtype == AVG_JOINT_REVOLUTE) d->S[b] = sv_make(k->axis[b], vcross(k->org[b], k->axis[b]));
        else d->S[b] = sv_make(V(0, 0, 0), k->axis[b]);
        sv vj = d->S[b]; for (int i = 0; i < 6; ++i) vj.v[i] *= qd;
        if (B->parent >= 0) { for (int i = 0; i < 6; ++i) d->vel[b].v[i] = d->vel[B->parent].v[i] + vj.v[i]; }
        else d->vel[b] = vj;
        d->cb[b] = crm(&d->vel[b], &vj);
        spatial_inertia(bmass, k->p[b], d->Iw[b], &d->IA[b]);
        sv Iv = sm_mul(&d->IA[b], &d->vel[b]);
        pA[b] = crf(&d->vel[b], &Iv);
        /* external force at the COM: gravity and Bullet's velocity damping */
        v3 w = sv_ang(&d->vel[b]);
        v3 vc = vadd(sv_lin(&d->vel[b]), vcross(w, k->p[b]));
        v3 f = vscale(f3(B->gravity), bmass);
        f = vsub(f, vscale(vc, bmass * (h->lin_damp + h->lin_damp * vnorm(vc))));
        v3 Iw_w = V(d->Iw[b][0][0] * w.x + d->Iw[b][0][1] * w.y + d->Iw[b][0][2] * w.z,
                    d->Iw[b][1][0] * w.x + d->Iw[b][1][1] * w.y + d->Iw[b][1][2] * w.z,
                    d->Iw[b][2][0] * w.x + d->Iw[b][2][1] * w.y + d->Iw[b][2][2] * w.z);
        v3 n = vneg(vscale(Iw_w, h->ang_damp + h->ang_damp * vnorm(w)));
        v3 nO = vadd(n, vcross(k->p[b], f));
        pA[b].v[0] -= nO.x; pA[b].v[1] -= nO.y; pA[b].v[2] -= nO.z;
        pA[b].v[3] -= f.x; pA[b].v[4] -= f.y; pA[b].v[5] -= f.z;
    }
    /* pass 2 */
    double u[MAXB];
    for (int b = nb - 1; b >= 0; --b) {
        const AvgBody* B = &m->body[b];
        if (B->jtype == AVG_JOINT_FREE) continue;
        d->U[b] = sm_mul(&d->IA[b], &d->S[b]);
        d->D[b] = sv_dot(&d->S[b], &d->U[b]);
        d->invD[b] = (d->D[b] >= 2.2e-16) ? 1.0 / d->D[b] : 0.0;       /* [UPSTREAM-BULLET] D < eps => joint frozen */
        /* joint torque: none applied (motors are constraint rows) except Bullet's explicit joint damping
         * -m_jointDamping * qd (URDF <dynamics damping>; nonzero on the PR2 only) [UPSTREAM-BULLET] */
        u[b] = -(double)m->dof[B->dof].damping * env[AVG_E_QD + B->dof] - sv_dot(&d->S[b], &pA[b]);
        if (B->parent >= 0) {
            sm Ia = d->IA[b];
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Ia.m[i][j] -= d->U[b].v[i] * d->invD[b] * d->U[b].v[j];
            sv Iac = sm_mul(&Ia, &d->cb[b]);
            for (int i = 0; i < 6; ++i) {
                pA[B->parent].v[i] += pA[b].v[i] + Iac.v[i] + d->U[b].v[i] * d->invD[b] * u[b];
                for (int j = 0; j < 6; ++j) d->IA[B->parent].m[i][j] += Ia.m[i][j];
            }
        }
    }
    /* pass 3 */
    sv acc[MAXB];
    for (int b = 0; b < nb; ++b) {
        const AvgBody* B = &m->body[b];
        if (B->jtype == AVG_JOINT_FREE) {
            /* free rigid body (floating-base btMultiBody without links): Newton-Euler with the same damping */
            const double* qd = env + AVG_E_QD + B->dof;
            v3 v = V(qd[0], qd[1], qd[2]), w = V(qd[3], qd[4], qd[5]);
            v3 g = f3(B->gravity);
            v3 a = vsub(g, vscale(v, h->lin_damp + h->lin_damp * vnorm(v)));
            v3 Iw_w = V(d->Iw[b][0][0] * w.x + d->Iw[b][0][1] * w.y + d->Iw[b][0][2] * w.z,
                        d->Iw[b][1][0] * w.x + d->Iw[b][1][1] * w.y + d->Iw[b][1][2] * w.z,
                        d->Iw[b][2][0] * w.x + d->Iw[b][2][1] * w.y + d->Iw[b][2][2] * w.z);
            v3 tau = vsub(vneg(vscale(Iw_w, h->ang_damp + h->ang_damp * vnorm(w))), vcross(w, Iw_w));
            v3 al = V(d->IwInv[b][0][0] * tau.x + d->IwInv[b][0][1] * tau.y + d->IwInv[b][0][2] * tau.z,
                      d->IwInv[b][1][0] * tau.x + d->IwInv[b][1][1] * tau.y + d->IwInv[b][1][2] * tau.z,
                      d->IwInv[b][2][0] * tau.x + d->IwInv[b][2][1] * tau.y + d->IwInv[b][2][2] * tau.z);
            if (B->mass <= 0) { a = V(0, 0, 0); }
            qdd[B->dof + 0] = a.x; qdd[B->dof + 1] = a.y; qdd[B->dof + 2] = a.z;
            qdd[B->dof + 3] = al.x; qdd[B->dof + 4] = al.y; qdd[B->dof + 5] = al.z;
            continue;
        }
        sv ap;
        if (B->parent >= 0) { for (int i = 0; i < 6; ++i) ap.v[i] = acc[B->parent].v[i] + d->cb[b].v[i]; }
        else ap = d->cb[b];
        double a = d->invD[b] * (u[b] - sv_dot(&d->U[b], &ap));
        qdd[B->dof] = a;
        for (int i = 0; i < 6; ++i) acc[b].v[i] = ap.v[i] + d->S[b].v[i] * a;
    }
}

/* response M^-1 tau for a generalized impulse tau (Bullet: calcAccelerationDeltasMultiDof) */
static void aba_delta(const Model* m, const Dyn* d, const double* tau, double* out) {
    int nb = m->h->n_body;
    sv p[MAXB]; double u[MAXB];
    memset(p, 0, sizeof(p));
    for (int b = nb - 1; b >= 0; --b) {
        const AvgBody* B = &m->body[b];
        if (B->jtype == AVG_JOINT_FREE) continue;
        u[b] = tau[B->dof] - sv_dot(&d->S[b], &p[b]);
        if (B->parent >= 0) for (int i = 0; i < 6; ++i) p[B->parent].v[i] += p[b].v[i] + d->U[b].v[i] * d->invD[b] * u[b];
    }
    sv acc[MAXB];
    for (int b = 0; b < nb; ++b) {
        const AvgBody* B = &m->body[b];
        if (B->jtype == AVG_JOINT_FREE) {
            const double* t = tau + B->dof;
            double im = B->mass > 0 ? 1.0 / B->mass : 0.0;
            out[B->dof + 0] = t[0] * im; out[B->dof + 1] = t[1] * im; out[B->dof + 2] = t[2] * im;
            for (int i = 0; i < 3; ++i) out[B->dof + 3 + i] = d->IwInv[b][i][0] * t[3] + d->IwInv[b][i][1] * t[4] + d->IwInv[b][i][2] * t[5];
            continue;
        }
        sv ap; memset(&ap, 0, sizeof(ap));
        if (B->parent >= 0) ap = acc[B->parent];
        double a = d->invD[b] * (u[b] - sv_dot(&d->U[b], &ap));
        out[B->dof] = a;
        for (int i = 0; i < 6; ++i) acc[b].v[i] = ap.v[i] + d->S[b].v[i] * a;
    }
}

/* Jacobian row of "velocity of the point r on `body` along n" (body < 0: static, contributes nothing) */
static void jac_point(const Model* m, const Kin* k, const Dyn* d, int body, v3 r, v3 n, double sign, double* J) {
    if (body < 0 || body >= m->h->n_body) return;             /* static world / env-static body */
    const AvgBody* B = &m->body[body];
    if (B->jtype == AVG_JOINT_FREE) {
        v3 rn = vcross(vsub(r, k->p[body]), n);
        J[B->dof + 0] += sign * n.x; J[B->dof + 1] += sign * n.y; J[B->dof + 2] += sign * n.z;
        J[B->dof + 3] += sign * rn.x; J[B->dof + 4] += sign * rn.y; J[B->dof + 5] += sign * rn.z;
        return;
    }
    sv f = sv_make(vcross(r, n), n);
    for (int b = body; b >= 0; b = m->body[b].parent) J[m->body[b].dof] += sign * sv_dot(&d->S[b], &f);
}
static void jac_ang(const Model* m, const Dyn* d, int body, v3 n, double sign, double* J) {
    if (body < 0 || body >= m->h->n_body) return;
    const AvgBody* B = &m->body[body];
    if (B->jtype == AVG_JOINT_FREE) { J[B->dof + 3] += sign * n.x; J[B->dof + 4] += sign * n.y; J[B->dof + 5] += sign * n.z; return; }
    sv f = sv_make(n, V(0, 0, 0));
    for (int b = body; b >= 0; b = m->body[b].parent) J[m->body[b].dof] += sign * sv_dot(&d->S[b], &f);
}

typedef struct {
    double J[MAXD], W[MAXD];
    double target;       /* desired change of J.v */
    double lo, hi;
    double diag, lambda;
    int friction_of;     /* row index of the normal row for friction rows, else -1 */
    double mu;
    /* particle part of the row (Feeding / Drinking): particle pa (and pb for particle-particle contacts), -1 = none.
     * Linear / angular Jacobian blocks and their M^-1 images (spheres: v += jl / m, w += ja / I). */
    int pa, pb;
    v3 jl_a, ja_a, jl_b, ja_b;
} Row;

static double limit_lo(const AvgDof* D, const double* env) { return (D->flags & AVG_DOF_HUMAN) ? D->lower * env[AVG_E_LIMIT_SCALE] : D->lower; }
static double limit_hi(const AvgDof* D, const double* env) { return (D->flags & AVG_DOF_HUMAN) ? D->upper * env[AVG_E_LIMIT_SCALE] : D->upper; }

static void finish_row(const Model* m, const Dyn* d, const double* qd, Row* r) {
    aba_delta(m, d, r->J, r->W);
    double diag = 0, u0 = 0;
    for (int i = 0; i < m->h->n_dof; ++i) { diag += r->J[i] * r->W[i]; u0 += r->J[i] * qd[i]; }
    r->diag = diag; r->lambda = 0;
    r->target -= u0;
    r->pa = r->pb = -1;
}

/* ---- food / water particles (feeding.py:291-307, drinking.py:291-312): free spheres, r = 5 mm, 1 g ------------------------
 * [UPSTREAM-BULLET] createMultiBody(baseMass, sphere, useMaximalCoordinates=False) makes each one a btMultiBody with a free
 * base and no links: gravity, the multibody velocity damping (0.04), sphere inertia 2/5 m r^2, contacts against everything.
 * Restated with these documented simplifications: a particle exchanges momentum with other particles and with the
 * tool (two-way, the tool is a free body); against links of the robot / human articulations the contact is one-way (the link
 * counts as kinematic with its start-of-step velocity: a 1 g sphere does not push a robot arm); particles removed from
 * self.foods / self.waters leave the simulation (the reference leaves spilled ones rolling on the floor). */
typedef struct {
    int p;               /* particle A */
    int q;               /* particle B or -1 */
    int shape_b;         /* shape index of B (-1 for particle-particle) */
    v3 n;                /* from B to A */
    double dist;
    v3 pb;               /* contact point on B */
} PContact;

#define P_X(part, p, c) ((part)[AVG_P_POS + 64 * (c) + (p)])
#define P_V(part, p, c) ((part)[AVG_P_VEL + 64 * (c) + (p)])
#define P_W(part, p, c) ((part)[AVG_P_ANG + 64 * (c) + (p)])
static int p_alive(const double* part, int p) { return (((uint32_t)part[AVG_P_ALIVE + (p >> 5)]) >> (p & 31)) & 1u; }
static void p_setbit(double* part, int slot, int p) { part[slot + (p >> 5)] = (double)(((uint32_t)part[slot + (p >> 5)]) | (1u << (p & 31))); }
static void p_clrbit(double* part, int slot, int p) { part[slot + (p >> 5)] = (double)(((uint32_t)part[slot + (p >> 5)]) & ~(1u << (p & 31))); }
static int p_getbit(const double* part, int slot, int p) { return (((uint32_t)part[slot + (p >> 5)]) >> (p & 31)) & 1u; }

/* Particle contacts of one internal step, in canonical order: for p ascending, the shapes in ascending shape-table index
 * (moving, static, compound children); then the particle-particle pairs (p, q > p) in lexicographic order. */
static int particle_collide(const Model* m, const Kin* k, const double* part, PContact* out, int* overflow) {
    const AvgModelHeader* h = m->h;
    int np = h->n_particle, nc = 0;
    if (np <= 0) return 0;
    const AvgShape* PS = &m->shape[h->pshape];
    const double r = PS->radius, pthr = PS->thr;
    int nall = h->n_shape + h->n_cshape;
    WShape* ws = (WShape*)malloc(sizeof(WShape) * nall);
    v3* ac = (v3*)malloc(sizeof(v3) * nall); v3* ah = (v3*)malloc(sizeof(v3) * nall);
    for (int i = 0; i < nall; ++i) { shape_world(m, k, i, &ws[i]); if (ws[i].s->type != AVG_SHAPE_COMPOUND) shape_aabb(&ws[i], &ac[i], &ah[i]); }
    for (int p = 0; p < np; ++p) {
        if (!p_alive(part, p)) continue;
        WShape A; A.s = PS; A.verts = 0; A.planes = 0;
        A.p = V(P_X(part, p, 0), P_X(part, p, 1), P_X(part, p, 2));
        quat id = {0, 0, 0, 1}; A.R = qmat(id);
        for (int b = 0; b < nall; ++b) {
            const AvgShape* S = ws[b].s;
            if (S->type == AVG_SHAPE_COMPOUND) continue;
            double thr = fmin(pthr, (double)S->thr);
            if (S->type != AVG_SHAPE_PLANE) {
                if (fabs(A.p.x - ac[b].x) > ah[b].x + r + thr) continue;
                if (fabs(A.p.y - ac[b].y) > ah[b].y + r + thr) continue;
                if (fabs(A.p.z - ac[b].z) > ah[b].z + r + thr) continue;
            } else if (A.p.z - r > thr) continue;
            Contact c;
            if (narrowphase(&A, &ws[b], thr, &c)) {
                if (nc >= AVG_MAX_PCONTACT) { *overflow |= 8; goto done; }
                out[nc].p = p; out[nc].q = -1; out[nc].shape_b = b; out[nc].n = c.n; out[nc].dist = c.dist; out[nc].pb = c.pb;
                nc++;
            }
        }
    }
    for (int p = 0; p < np; ++p) {
        if (!p_alive(part, p)) continue;
        v3 xp = V(P_X(part, p, 0), P_X(part, p, 1), P_X(part, p, 2));
        for (int q = p + 1; q < np; ++q) {
            if (!p_alive(part, q)) continue;
            v3 xq = V(P_X(part, q, 0), P_X(part, q, 1), P_X(part, q, 2));
            v3 dlt = vsub(xp, xq);
            double dd = vnorm(dlt), dist = dd - 2.0 * r;
            if (dist >= pthr) continue;
            if (nc >= AVG_MAX_PCONTACT) { *overflow |= 8; goto done; }
            out[nc].p = p; out[nc].q = q; out[nc].shape_b = -1; out[nc].dist = dist;
            out[nc].n = dd > 1e-12 ? vscale(dlt, 1.0 / dd) : V(0, 0, 1);
            out[nc].pb = vadd(xq, vscale(out[nc].n, r));
            nc++;
        }
    }
done:
    free(ws); free(ac); free(ah);
    return nc;
}

/* Order of the particle rows inside a block of the Gauss-Seidel sweep.  Bullet's order is the order of its overlapping-pair
 * cache (hash order: not reproducible from the reference, and of no physical meaning), so the order is build-defined, and it
 * is chosen so that the device can sweep many rows at once: the contacts, taken in canonical order (particle_collide), are
 * coloured greedily -- each gets the lowest colour not yet used by a contact of its particle(s), at most 32 contacts per
 * colour -- and the sweep runs colour by colour, canonical order inside a colour.  Contacts of one colour share no particle,
 * so they commute.  The kernels compute the same colouring (particles_prepare in avg_kernels.cu). */
static void particle_row_order(PContact* pcs, int npc, int* overflow) {
    unsigned long long used[AVG_MAX_PARTICLE]; memset(used, 0, sizeof(used));
    unsigned long long full = 0;
    int fill[64]; memset(fill, 0, sizeof(fill));
    int* colour = (int*)malloc(sizeof(int) * npc);
    for (int c = 0; c < npc; ++c) {
        unsigned long long mk = used[pcs[c].p] | (pcs[c].q >= 0 ? used[pcs[c].q] : 0ull) | full;
        int r = 63;
        if (mk != ~0ull) { r = 0; while ((mk >> r) & 1ull) ++r; } else *overflow |= 8;
        colour[c] = r;
        used[pcs[c].p] |= 1ull << r; if (pcs[c].q >= 0) used[pcs[c].q] |= 1ull << r;
        if (++fill[r] >= 32) full |= 1ull << r;
    }
    PContact* tmp = (PContact*)malloc(sizeof(PContact) * npc);
    int n = 0;
    for (int r = 0; r < 64; ++r) for (int c = 0; c < npc; ++c) if (colour[c] == r) tmp[n++] = pcs[c];
    memcpy(pcs, tmp, sizeof(PContact) * npc);
    free(tmp); free(colour);
}

static v3 plane_space1(v3 n) {                       /* btPlaneSpace1 */
    if (fabs(n.z) > 0.7071067811865476) { double a = n.y * n.y + n.z * n.z, kk = 1.0 / sqrt(a); return V(0, -n.z * kk, n.y * kk); }
    double a = n.x * n.x + n.y * n.y, kk = 1.0 / sqrt(a); return V(-n.y * kk, n.x * kk, 0);
}

/* one internal physics step: p.stepSimulation() with numSubSteps = 0 (scratch_itch.py:258), or one of the numSubSteps = 2
 * internal steps of Feeding / Drinking (feeding.py:289).  `part` = particle record (float64 image, ints as numbers) or NULL. */
static void substep(const Model* m, double* env, double* part, Contact* contacts, int* ncontact) {
    const AvgModelHeader* h = m->h;
    int nd = h->n_dof, nb = h->n_body;
    double dt = h->dt;
    Kin k; Dyn* d = (Dyn*)malloc(sizeof(Dyn));
    fk(m, env, &k);
    int overflow = 0;
    int nc = collide(m, &k, contacts, &overflow);
    const int np = part ? h->n_particle : 0;
    PContact* pcs = np > 0 ? (PContact*)malloc(sizeof(PContact) * AVG_MAX_PCONTACT) : 0;
    int npc = np > 0 ? particle_collide(m, &k, part, pcs, &overflow) : 0;
    if (npc > 1) particle_row_order(pcs, npc, &overflow);
    /* unconstrained velocity update */
    double qdd[MAXD], qd[MAXD];
    aba(m, &k, env, d, qdd);
    for (int i = 0; i < nd; ++i) {
        qd[i] = env[AVG_E_QD + i] + dt * qdd[i];
    }
    /* particles: gravity and the multibody base damping, v* = v + dt (g - v (k + k |v|)), w* = w - dt w (k + k |w|) */
    double pv[AVG_MAX_PARTICLE][6];
    const double pm = h->p_mass, prad = np > 0 ? m->shape[h->pshape].radius : 0.0;
    const double pinv_m = pm > 0 ? 1.0 / pm : 0.0, pinv_i = pm > 0 ? 1.0 / (0.4 * pm * prad * prad) : 0.0;
    for (int p = 0; p < np; ++p) {
        v3 v = V(P_V(part, p, 0), P_V(part, p, 1), P_V(part, p, 2)), w = V(P_W(part, p, 0), P_W(part, p, 1), P_W(part, p, 2));
        v3 a = vsub(f3(h->p_gravity), vscale(v, h->lin_damp + h->lin_damp * vnorm(v)));
        v3 al = vneg(vscale(w, h->ang_damp + h->ang_damp * vnorm(w)));
        pv[p][0] = v.x + dt * a.x; pv[p][1] = v.y + dt * a.y; pv[p][2] = v.z + dt * a.z;
        pv[p][3] = w.x + dt * al.x; pv[p][4] = w.y + dt * al.y; pv[p][5] = w.z + dt * al.z;
    }
    /* constraint rows, Bullet order: non-contact (motors, limits, fixed constraint), contact normals, friction */
    const int maxr = MAXR + 2 * AVG_MAX_PCONTACT;
    Row* rows = (Row*)calloc(maxr, sizeof(Row));
    int nr = 0;
    for (int i = 0; i < h->n_jdof; ++i) {          /* position motors, btMultiBodyJointMotor [UPSTREAM-BULLET] */
        const AvgDof* D = &m->dof[i];
        if (!(D->flags & AVG_DOF_MOTOR)) continue;
        Row* r = &rows[nr++]; memset(r, 0, sizeof(Row));
        double kp = (D->flags & AVG_DOF_HUMAN) ? env[AVG_E_HUMAN_KP] : D->kp;
        double maxf = (D->flags & AVG_DOF_HUMAN) ? h->task_f[AVG_TF_HUMAN_FORCE] * env[AVG_E_STRENGTH] : D->max_force;
        double q = env[AVG_E_Q + m->body[D->body].qidx];
        r->J[i] = 1.0;
        r->target = kp * (env[AVG_E_MTARGET + i] - q) / dt + (1.0 - D->kd) * qd[i];   /* desired velocity */
        r->lo = -maxf * dt; r->hi = maxf * dt; r->friction_of = -1;
        finish_row(m, d, qd, r);
    }
    for (int i = 0; i < h->n_jdof; ++i) {          /* joint limits, only while violated [UPSTREAM-BULLET] */
        const AvgDof* D = &m->dof[i];
        if (!(D->flags & AVG_DOF_LIMIT)) continue;
        double q = env[AVG_E_Q + m->body[D->body].qidx];
        double pen[2] = {q - limit_lo(D, env), limit_hi(D, env) - q};
        for (int s = 0; s < 2; ++s) {
            if (pen[s] > 0 || nr >= MAXR) continue;
            Row* r = &rows[nr++]; memset(r, 0, sizeof(Row));
            r->J[i] = s ? -1.0 : 1.0;
            r->target = -pen[s] * h->erp / dt;
            r->lo = 0; r->hi = 100.0; r->friction_of = -1;
            finish_row(m, d, qd, r);
        }
    }
    {                                               /* tool weld: btMultiBodyFixedConstraint, 6 rows */
        v3 pa, pb; quat qa, qb;
        frame_pose(m, &k, AVG_F_WELD_PARENT, &pa, &qa);
        frame_pose(m, &k, AVG_F_TOOL_BASE, &pb, &qb);
        quat dq = qmul(qa, qconj(qb));
        if (dq.w < 0) { dq.x = -dq.x; dq.y = -dq.y; dq.z = -dq.z; dq.w = -dq.w; }
        double s = sqrt(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z);
        v3 rotv = V(0, 0, 0);
        if (s > 1e-12) { double ang = 2.0 * atan2(s, dq.w); rotv = V(dq.x / s * ang, dq.y / s * ang, dq.z / s * ang); }
        v3 perr = vsub(pa, pb);
        double maxi = h->weld_max_force * dt;
        for (int ax = 0; ax < 6 && nr < MAXR; ++ax) {
            Row* r = &rows[nr++]; memset(r, 0, sizeof(Row));
            v3 e = V(ax % 3 == 0, ax % 3 == 1, ax % 3 == 2);
            if (ax < 3) {
                jac_point(m, &k, d, h->weld_body_a, pa, e, 1.0, r->J);
                jac_point(m, &k, d, h->weld_body_b, pb, e, -1.0, r->J);
                r->target = -vdot(perr, e) * h->erp / dt;
            } else {
                jac_ang(m, d, h->weld_body_a, e, 1.0, r->J);
                jac_ang(m, d, h->weld_body_b, e, -1.0, r->J);
                r->target = -vdot(rotv, e) * h->erp / dt;
            }
            r->lo = -maxi; r->hi = maxi; r->friction_of = -1;
            finish_row(m, d, qd, r);
        }
    }
    if (nc > (MAXR - nr) / 2) { overflow |= 2; nc = (MAXR - nr) / 2; }      /* flagged, never silent */
    int first_contact_row = nr;
    for (int c = 0; c < nc; ++c) {                  /* contact normals */
        Contact* C = &contacts[c];
        Row* r = &rows[nr++]; memset(r, 0, sizeof(Row));
        jac_point(m, &k, d, m->shape[C->sa].body, C->pa, C->n, 1.0, r->J);
        jac_point(m, &k, d, m->shape[C->sb].body, C->pb, C->n, -1.0, r->J);
        r->target = (C->dist > 0) ? -C->dist / dt : -C->dist * h->erp / dt;
        r->lo = 0; r->hi = 1e30; r->friction_of = -1;
        finish_row(m, d, qd, r);
    }
    /* particle contact normals.  Body B is another particle, the tool (free body: two-way), or kinematic (static shape, or a
     * link of an articulation with its start-of-step point velocity).  A sphere's normal impulse passes through its centre. */
    int first_pcontact_row = nr;
    v3* pc_vb = npc > 0 ? (v3*)malloc(sizeof(v3) * npc) : 0;      /* velocity of B's contact point when B is kinematic */
    for (int c = 0; c < npc; ++c) {
        PContact* C = &pcs[c];
        Row* r = &rows[nr++]; memset(r, 0, sizeof(Row));
        r->pa = C->p; r->pb = C->q; r->jl_a = C->n; r->ja_a = V(0, 0, 0);
        double diag = pinv_m, u0 = C->n.x * pv[C->p][0] + C->n.y * pv[C->p][1] + C->n.z * pv[C->p][2];
        pc_vb[c] = V(0, 0, 0);
        if (C->q >= 0) {
            r->jl_b = vneg(C->n); r->ja_b = V(0, 0, 0);
            diag += pinv_m; u0 -= C->n.x * pv[C->q][0] + C->n.y * pv[C->q][1] + C->n.z * pv[C->q][2];
        } else {
            int body = m->shape[C->shape_b].body;
            if (body >= 0 && body < nb && body == h->tool_body) {
                jac_point(m, &k, d, body, C->pb, C->n, -1.0, r->J);
                aba_delta(m, d, r->J, r->W);
                for (int i = 0; i < nd; ++i) { diag += r->J[i] * r->W[i]; u0 += r->J[i] * qd[i]; }
            } else if (body >= 0 && body < nb) {
                double Jx[3][MAXD]; memset(Jx, 0, sizeof(Jx));
                for (int ax = 0; ax < 3; ++ax) jac_point(m, &k, d, body, C->pb, V(ax == 0, ax == 1, ax == 2), 1.0, Jx[ax]);
                double vb[3] = {0, 0, 0};
                for (int ax = 0; ax < 3; ++ax) for (int i = 0; i < nd; ++i) vb[ax] += Jx[ax][i] * env[AVG_E_QD + i];
                pc_vb[c] = V(vb[0], vb[1], vb[2]);
                u0 -= vdot(C->n, pc_vb[c]);
            }
        }
        r->diag = diag; r->lambda = 0;
        r->target = ((C->dist > 0) ? -C->dist / dt : -C->dist * h->erp / dt) - u0;
        r->lo = 0; r->hi = 1e30; r->friction_of = -1;
    }
    for (int c = 0; c < nc; ++c) {                  /* one friction direction per contact [UPSTREAM-BULLET default] */
        Contact* C = &contacts[c];
        Row* r = &rows[nr++]; memset(r, 0, sizeof(Row));
        /* lateral relative velocity of the contact points before the solve */
        double Jx[3][MAXD]; memset(Jx, 0, sizeof(Jx));
        v3 vrel;
        for (int ax = 0; ax < 3; ++ax) {
            v3 e = V(ax == 0, ax == 1, ax == 2);
            jac_point(m, &k, d, m->shape[C->sa].body, C->pa, e, 1.0, Jx[ax]);
            jac_point(m, &k, d, m->shape[C->sb].body, C->pb, e, -1.0, Jx[ax]);
        }
        double vr[3] = {0, 0, 0};
        for (int ax = 0; ax < 3; ++ax) for (int i = 0; i < nd; ++i) vr[ax] += Jx[ax][i] * qd[i];
        vrel = V(vr[0], vr[1], vr[2]);
        v3 lat = vsub(vrel, vscale(C->n, vdot(vrel, C->n)));
        double ll = vnorm(lat);
        v3 t;
        if (ll > 1e-6) t = vscale(lat, 1.0 / ll);
        else t = plane_space1(C->n);
        for (int i = 0; i < nd; ++i) r->J[i] = t.x * Jx[0][i] + t.y * Jx[1][i] + t.z * Jx[2][i];
        r->target = 0; r->lo = 0; r->hi = 0; r->friction_of = first_contact_row + c;
        r->mu = (double)m->shape[C->sa].friction * (double)m->shape[C->sb].friction;
        finish_row(m, d, qd, r);
    }
    const int first_pfriction_row = nr;
    for (int c = 0; c < npc; ++c) {                 /* particle friction rows: the sphere's lever arm is -r n */
        PContact* C = &pcs[c];
        Row* r = &rows[nr++]; memset(r, 0, sizeof(Row));
        r->pa = C->p; r->pb = C->q;
        const double* va = pv[C->p];
        v3 ra = vscale(C->n, -prad);
        v3 vrel = vadd(V(va[0], va[1], va[2]), vcross(V(va[3], va[4], va[5]), ra));
        double Jx[3][MAXD]; memset(Jx, 0, sizeof(Jx));
        int body = -1, twoway = 0;
        double mu_b = 0.5;
        if (C->q >= 0) {
            const double* vb = pv[C->q];
            v3 rb = vscale(C->n, prad);
            vrel = vsub(vrel, vadd(V(vb[0], vb[1], vb[2]), vcross(V(vb[3], vb[4], vb[5]), rb)));
            mu_b = m->shape[h->pshape].friction;
        } else {
            body = m->shape[C->shape_b].body;
            mu_b = m->shape[C->shape_b].friction;
            if (body >= 0 && body < nb && body == h->tool_body) {
                twoway = 1;
                for (int ax = 0; ax < 3; ++ax) jac_point(m, &k, d, body, C->pb, V(ax == 0, ax == 1, ax == 2), -1.0, Jx[ax]);
                double vr[3] = {0, 0, 0};
                for (int ax = 0; ax < 3; ++ax) for (int i = 0; i < nd; ++i) vr[ax] += Jx[ax][i] * qd[i];
                vrel = vadd(vrel, V(vr[0], vr[1], vr[2]));
            } else vrel = vsub(vrel, pc_vb[c]);
        }
        v3 lat = vsub(vrel, vscale(C->n, vdot(vrel, C->n)));
        double ll = vnorm(lat);
        v3 t = ll > 1e-6 ? vscale(lat, 1.0 / ll) : plane_space1(C->n);
        r->jl_a = t; r->ja_a = vcross(ra, t);
        double diag = pinv_m + pinv_i * vdot(r->ja_a, r->ja_a);
        if (C->q >= 0) {
            v3 rb = vscale(C->n, prad);
            r->jl_b = vneg(t); r->ja_b = vneg(vcross(rb, t));
            diag += pinv_m + pinv_i * vdot(r->ja_b, r->ja_b);
        } else if (twoway) {
            for (int i = 0; i < nd; ++i) r->J[i] = t.x * Jx[0][i] + t.y * Jx[1][i] + t.z * Jx[2][i];
            aba_delta(m, d, r->J, r->W);
            for (int i = 0; i < nd; ++i) diag += r->J[i] * r->W[i];
        }
        r->diag = diag; r->lambda = 0;
        r->target = -vdot(vrel, t);                 /* target change of J.v: cancel the lateral velocity (clamped by mu * normal impulse) */
        r->lo = 0; r->hi = 0; r->friction_of = first_pcontact_row + c;
        r->mu = (double)m->shape[h->pshape].friction * mu_b;
    }
    /* Warm starting of the contact normal rows [UPSTREAM-BULLET: btMultiBodyConstraintSolver::setupMultiBodyContactConstraint,
     * m_appliedImpulse = cp.m_appliedImpulse * m_warmstartingFactor for normal rows, 0 for friction rows; motors, limits and the
     * fixed constraint are not warm-started ("disable warmstarting for btMultiBody ... gaining energy")].  Bullet finds the old
     * impulse in the persistent manifold point; with one point per shape pair the match is by pair: the record keeps the
     * (pair, impulse) of the first AVG_WCACHE_N contact points of the last internal step (AVG_E_WCACHE). */
    double dv[MAXD]; memset(dv, 0, sizeof(dv));
    if (h->warmstart > 0.0f) {
        for (int c = 0; c < nc; ++c) {
            const uint32_t key = (uint32_t)contacts[c].sa | ((uint32_t)contacts[c].sb << 16);
            for (int w = 0; w < AVG_WCACHE_N; ++w) {
                if ((uint32_t)env[AVG_E_WCACHE + 2 * w] != key) continue;
                Row* r = &rows[first_contact_row + c];
                r->lambda = (double)h->warmstart * env[AVG_E_WCACHE + 2 * w + 1];
                for (int i = 0; i < nd; ++i) dv[i] += r->W[i] * r->lambda;
                break;
            }
        }
    }
    /* projected Gauss-Seidel in velocity space, as btMultiBodyConstraintSolver */
    /* Particle rows and the tool (documented deviation from Bullet's strict row order): inside one block of particle
     * rows (the normal block, the friction block) every row sees the tool's velocity change as it was when the block began; the
     * reactions of the block reach the tool -- and the rows after the block -- at once.  A block then only orders the rows that
     * share a PARTICLE, which is what lets the device sweep a block in a few parallel rounds instead of one row at a time; the
     * coupling that is lagged by one block is of the order particle mass / tool mass (1 g against the spoon / cup). */
    double dv_block[MAXD]; memset(dv_block, 0, sizeof(dv_block));
    double dpv[AVG_MAX_PARTICLE][6]; memset(dpv, 0, sizeof(dpv));
    for (int it = 0; it < h->solver_iters; ++it) {
        double resid = 0;
        for (int ri = 0; ri < nr; ++ri) {
            Row* r = &rows[ri];
            if (npc > 0 && (ri == first_pcontact_row || ri == first_pfriction_row)) memcpy(dv_block, dv, sizeof(dv));
            if (r->diag < 1e-12) continue;
            double lo = r->lo, hi = r->hi;
            if (r->friction_of >= 0) { double lim = r->mu * rows[r->friction_of].lambda; lo = -lim; hi = lim; }
            const double* dvr = r->pa >= 0 ? dv_block : dv;
            double jdv = 0; for (int i = 0; i < nd; ++i) jdv += r->J[i] * dvr[i];
            if (r->pa >= 0) { const double* a = dpv[r->pa]; jdv += r->jl_a.x * a[0] + r->jl_a.y * a[1] + r->jl_a.z * a[2] + r->ja_a.x * a[3] + r->ja_a.y * a[4] + r->ja_a.z * a[5]; }
            if (r->pb >= 0) { const double* b = dpv[r->pb]; jdv += r->jl_b.x * b[0] + r->jl_b.y * b[1] + r->jl_b.z * b[2] + r->ja_b.x * b[3] + r->ja_b.y * b[4] + r->ja_b.z * b[5]; }
            double delta = (r->target - jdv) / r->diag;
            double sum = r->lambda + delta;
            if (sum < lo) sum = lo; if (sum > hi) sum = hi;
            delta = sum - r->lambda; r->lambda = sum;
            for (int i = 0; i < nd; ++i) dv[i] += r->W[i] * delta;
            if (r->pa >= 0) { double* a = dpv[r->pa]; a[0] += pinv_m * r->jl_a.x * delta; a[1] += pinv_m * r->jl_a.y * delta; a[2] += pinv_m * r->jl_a.z * delta;
                              a[3] += pinv_i * r->ja_a.x * delta; a[4] += pinv_i * r->ja_a.y * delta; a[5] += pinv_i * r->ja_a.z * delta; }
            if (r->pb >= 0) { double* b = dpv[r->pb]; b[0] += pinv_m * r->jl_b.x * delta; b[1] += pinv_m * r->jl_b.y * delta; b[2] += pinv_m * r->jl_b.z * delta;
                              b[3] += pinv_i * r->ja_b.x * delta; b[4] += pinv_i * r->ja_b.y * delta; b[5] += pinv_i * r->ja_b.z * delta; }
            double rv = delta * r->diag; if (rv * rv > resid) resid = rv * rv;
        }
        if (resid <= h->residual_thr) break;
    }
    for (int c = 0; c < nc; ++c) contacts[c].lambda_n = rows[first_contact_row + c].lambda;
    for (int w = 0; w < AVG_WCACHE_N; ++w) {          /* what the next internal step warm-starts from */
        env[AVG_E_WCACHE + 2 * w] = w < nc ? (double)((uint32_t)contacts[w].sa | ((uint32_t)contacts[w].sb << 16)) : 0.0;
        env[AVG_E_WCACHE + 2 * w + 1] = w < nc ? contacts[w].lambda_n : 0.0;
    }
    *ncontact = nc;
    /* integrate (btMultiBody::stepPositionsMultiDof): semi-implicit Euler */
    for (int i = 0; i < nd; ++i) {
        double v = qd[i] + dv[i];
        if (i < h->n_jdof) { if (v > h->max_vel) v = h->max_vel; if (v < -h->max_vel) v = -h->max_vel; }
        env[AVG_E_QD + i] = v;
    }
    for (int b = 0; b < nb; ++b) {
        const AvgBody* B = &m->body[b];
        if (B->jtype == AVG_JOINT_FREE) {
            double* q = env + AVG_E_Q + B->qidx; const double* v = env + AVG_E_QD + B->dof;
            q[0] += dt * v[0]; q[1] += dt * v[1]; q[2] += dt * v[2];
            v3 w = V(v[3], v[4], v[5]); double wn = vnorm(w);
            quat cur = {q[3], q[4], q[5], q[6]};
            if (wn * dt > 1e-12) cur = qmul(qaxis(vscale(w, 1.0 / wn), wn * dt), cur);
            cur = qnormalize(cur);
            q[3] = cur.x; q[4] = cur.y; q[5] = cur.z; q[6] = cur.w;
        } else env[AVG_E_Q + B->qidx] += dt * env[AVG_E_QD + B->dof];
    }
    if (np > 0) {
        /* which live particles have a contact point with the human / the table or the bowl (feeding.py:111,116: existence of
         * getContactPoints entries, i.e. of manifold points of this step's collision pass) */
        part[AVG_P_TOUCH_HUMAN] = part[AVG_P_TOUCH_HUMAN + 1] = 0; part[AVG_P_TOUCH_SPILL] = part[AVG_P_TOUCH_SPILL + 1] = 0;
        for (int c = 0; c < npc; ++c) {
            if (pcs[c].q >= 0) continue;
            int rb = m->shape[pcs[c].shape_b].ref_body;
            if (rb == AVG_REF_HUMAN) p_setbit(part, AVG_P_TOUCH_HUMAN, pcs[c].p);
            if (rb == AVG_REF_TABLE || rb == AVG_REF_BOWL) p_setbit(part, AVG_P_TOUCH_SPILL, pcs[c].p);
        }
        for (int p = 0; p < np; ++p) {
            if (!p_alive(part, p)) continue;
            for (int c = 0; c < 3; ++c) {
                P_V(part, p, c) = pv[p][c] + dpv[p][c]; P_W(part, p, c) = pv[p][3 + c] + dpv[p][3 + c];
                P_X(part, p, c) += dt * P_V(part, p, c);
            }
        }
        part[AVG_P_NCONTACT] = npc;
        if (overflow) part[AVG_P_NCONTACT + 1] = (double)((int)part[AVG_P_NCONTACT + 1] | overflow);
    }
    if (overflow) env[AVG_E_OVERFLOW] = (double)((int)env[AVG_E_OVERFLOW] | overflow);
    free(rows); free(d); free(pcs); free(pc_vb);
}

/* env.py:353-371: Keras Sequential Dense 4 -> 64 tanh -> 64 tanh -> 64 tanh -> 1 sigmoid on the remapped right-arm
 * angles; class 1 (sigmoid > 0.5 <=> logit > 0) remembers the pose, class 0 restores the last valid pose with zero
 * velocity.  Weights come from realistic_arm_limits_model.h5 via the model compiler. */
static double pymod(double a, double m) { double r = fmod(a, m); if (r < 0) r += m; return r; }
static double arm_limit_logit(const float* w, const double x[4]) {
    const float* W1 = w; const float* b1 = W1 + 256; const float* W2 = b1 + 64; const float* b2 = W2 + 4096;
    const float* W3 = b2 + 64; const float* b3 = W3 + 4096; const float* W4 = b3 + 64; const float* b4 = W4 + 64;
    double h1[64], h2[64], h3[64];
    for (int u = 0; u < 64; ++u) { double a = b1[u]; for (int k = 0; k < 4; ++k) a += x[k] * W1[k * 64 + u]; h1[u] = tanh(a); }
    for (int u = 0; u < 64; ++u) { double a = b2[u]; for (int k = 0; k < 64; ++k) a += h1[k] * W2[k * 64 + u]; h2[u] = tanh(a); }
    for (int u = 0; u < 64; ++u) { double a = b3[u]; for (int k = 0; k < 64; ++k) a += h2[k] * W3[k * 64 + u]; h3[u] = tanh(a); }
    double a = b4[0]; for (int k = 0; k < 64; ++k) a += h3[k] * W4[k];
    return a;
}
static void enforce_realistic_limits(const Model* m, double* env) {
    const AvgModelHeader* h = m->h;
    if (!m->mlp || !h->human_control) return;
    double q[4];
    for (int k = 0; k < 4; ++k) { if (h->mlp_dof[k] < 0) return; q[k] = env[AVG_E_Q + m->body[m->dof[h->mlp_dof[k]].body].qidx]; }
    const double twopi = 2.0 * 3.14159265358979323846;
    double x[4] = {pymod(-q[0] + twopi, twopi), pymod(q[1] + twopi, twopi), -q[2], pymod(-q[3] + twopi, twopi)};   /* env.py:360-363 */
    if (arm_limit_logit(m->mlp, x) > 0.0) {
        for (int k = 0; k < 4; ++k) env[AVG_E_VALID_POSE + k] = q[k];
        env[AVG_E_HAS_VALID] = 1;
    } else if (env[AVG_E_HAS_VALID] != 0) {
        for (int k = 0; k < 4; ++k) {
            env[AVG_E_Q + m->body[m->dof[h->mlp_dof[k]].body].qidx] = env[AVG_E_VALID_POSE + k];
            env[AVG_E_QD + h->mlp_dof[k]] = 0;
        }
    }
}

/* env.py:389-410 */
static void enforce_hard_limits(const Model* m, double* env) {
    for (int i = 0; i < m->h->n_jdof; ++i) {
        const AvgDof* D = &m->dof[i];
        if (!(D->flags & AVG_DOF_HARD_LIMIT)) continue;
        double* q = &env[AVG_E_Q + m->body[D->body].qidx];
        double lo = limit_lo(D, env), hi = limit_hi(D, env);
        if (*q < lo) { *q = lo; env[AVG_E_QD + i] = 0; }
        else if (*q > hi) { *q = hi; env[AVG_E_QD + i] = 0; }
    }
}

/* scratch_itch.py:289-293 */
static void update_target(const Model* m, double* env) {
    Kin k; fk(m, env, &k);
    v3 p; quat q;
    frame_pose(m, &k, (int)env[AVG_E_LIMB_FRAME], &p, &q);
    v3 t = vadd(p, qrot(q, V(env[AVG_E_TARGET_ON_ARM], env[AVG_E_TARGET_ON_ARM + 1], env[AVG_E_TARGET_ON_ARM + 2])));
    env[AVG_E_TARGET_POS] = t.x; env[AVG_E_TARGET_POS + 1] = t.y; env[AVG_E_TARGET_POS + 2] = t.z;
}

/* scratch_itch.py:104-128 */
static void get_obs(const Model* m, const double* env, double tool_force, double total_force_on_human,
                    double tool_force_at_target, double* obs) {
    const AvgModelHeader* h = m->h;
    Kin k; fk(m, env, &k);
    v3 torso, tool, sh, el, wr, chest; quat tq, dummy;
    frame_pose(m, &k, AVG_F_TORSO, &torso, &dummy);
    frame_pose(m, &k, AVG_F_TOOL_TIP, &tool, &tq);
    frame_pose(m, &k, AVG_F_SHOULDER, &sh, &dummy);
    frame_pose(m, &k, AVG_F_ELBOW, &el, &dummy);
    frame_pose(m, &k, AVG_F_WRIST, &wr, &dummy);
    frame_pose(m, &k, AVG_F_CHEST, &chest, &dummy);
    v3 tgt = V(env[AVG_E_TARGET_POS], env[AVG_E_TARGET_POS + 1], env[AVG_E_TARGET_POS + 2]);
    int o = 0;
#define PUT3(v) do { obs[o++] = (v).x; obs[o++] = (v).y; obs[o++] = (v).z; } while (0)
    v3 t;
    t = vsub(tool, torso); PUT3(t);
    obs[o++] = tq.x; obs[o++] = tq.y; obs[o++] = tq.z; obs[o++] = tq.w;
    t = vsub(tool, tgt); PUT3(t);
    t = vsub(tgt, torso); PUT3(t);
    for (int i = 0; i < h->n_jdof; ++i) if (m->dof[i].action >= 0 && m->dof[i].action < h->n_action_robot) obs[o++] = env[AVG_E_Q + m->body[m->dof[i].body].qidx];
    t = vsub(sh, torso); PUT3(t);
    t = vsub(el, torso); PUT3(t);
    t = vsub(wr, torso); PUT3(t);
    obs[o++] = tool_force;
    if (h->human_control) {
        t = vsub(tool, chest); PUT3(t);
        obs[o++] = tq.x; obs[o++] = tq.y; obs[o++] = tq.z; obs[o++] = tq.w;
        t = vsub(tool, tgt); PUT3(t);
        t = vsub(tgt, chest); PUT3(t);
        double hq[10]; memset(hq, 0, sizeof(hq));
        for (int i = 0; i < h->n_jdof; ++i) if (m->dof[i].human_slot >= 0) hq[m->dof[i].human_slot] = env[AVG_E_Q + m->body[m->dof[i].body].qidx];
        for (int i = 0; i < 10; ++i) obs[o++] = hq[i];
        t = vsub(sh, chest); PUT3(t);
        t = vsub(el, chest); PUT3(t);
        t = vsub(wr, chest); PUT3(t);
        obs[o++] = total_force_on_human; obs[o++] = tool_force_at_target;
    }
#undef PUT3
}

/* ---- BedBathing (bed_bathing.py) ------------------------------------------------------------------------------ */
/* world position of wiping target t: frame of human link 9 (upper arm) or 11 (forearm) applied to the point on the limb,
 * bed_bathing.py:382-394 */
static v3 bb_target_world(const Model* m, const Kin* k, int t) {
    v3 p; quat q;
    frame_pose(m, k, t < m->h->n_target_upper ? AVG_F_SHOULDER : AVG_F_ELBOW, &p, &q);
    return vadd(p, qrot(q, V(m->target[4 * t], m->target[4 * t + 1], m->target[4 * t + 2])));
}
/* min over every (tool link, human link) pair of the closest-point distance, bed_bathing.py:61
 * (p.getClosestPoints(tool, human, distance=4.0): one point per pair of collision shapes closer than 4 m) */
static double bb_closest_tool_human(const Model* m, const Kin* k) {
    int ns = m->h->n_shape;
    double best = 1e300;
    double range = m->h->task_f[AVG_TF_CLOSEST_RANGE];
    for (int a = 0; a < ns; ++a) {
        if (m->shape[a].ref_body != AVG_REF_TOOL) continue;
        WShape wa; shape_world(m, k, a, &wa);
        for (int b = 0; b < ns; ++b) {
            if (m->shape[b].ref_body != AVG_REF_HUMAN) continue;
            WShape wb; shape_world(m, k, b, &wb);
            Contact c;
            if (narrowphase(&wa, &wb, range, &c) && c.dist < best) best = c.dist;
        }
    }
    return best;
}
/* bed_bathing.py:129-153 */
static void bb_get_obs(const Model* m, const double* env, double tool_force, double total_force_on_human,
                       double tool_force_on_human, double* obs) {
    const AvgModelHeader* h = m->h;
    Kin k; fk(m, env, &k);
    v3 torso, tool, sh, el, wr; quat tq, dummy;
    frame_pose(m, &k, AVG_F_TORSO, &torso, &dummy);
    frame_pose(m, &k, AVG_F_TOOL_TIP, &tool, &tq);
    frame_pose(m, &k, AVG_F_SHOULDER, &sh, &dummy);
    frame_pose(m, &k, AVG_F_ELBOW, &el, &dummy);
    frame_pose(m, &k, AVG_F_WRIST, &wr, &dummy);
    int o = 0;
    v3 t;
    t = vsub(tool, torso); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
    obs[o++] = tq.x; obs[o++] = tq.y; obs[o++] = tq.z; obs[o++] = tq.w;
    for (int i = 0; i < h->n_jdof; ++i) if (m->dof[i].action >= 0 && m->dof[i].action < h->n_action_robot) obs[o++] = env[AVG_E_Q + m->body[m->dof[i].body].qidx];
    t = vsub(sh, torso); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
    t = vsub(el, torso); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
    t = vsub(wr, torso); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
    obs[o++] = tool_force;
    if (h->human_control) {                                   /* :136-139,149: positions relative to human link 3 */
        v3 chest; frame_pose(m, &k, AVG_F_CHEST, &chest, &dummy);
        t = vsub(tool, chest); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
        obs[o++] = tq.x; obs[o++] = tq.y; obs[o++] = tq.z; obs[o++] = tq.w;
        double hq[10]; memset(hq, 0, sizeof(hq));
        for (int i = 0; i < h->n_jdof; ++i) if (m->dof[i].human_slot >= 0) hq[m->dof[i].human_slot] = env[AVG_E_Q + m->body[m->dof[i].body].qidx];
        for (int i = 0; i < 10; ++i) obs[o++] = hq[i];
        t = vsub(sh, chest); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
        t = vsub(el, chest); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
        t = vsub(wr, chest); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
        obs[o++] = total_force_on_human; obs[o++] = tool_force_on_human;
    }
}
/* The part of BedBathingEnv.step after take_step (bed_bathing.py:53-75) with get_total_force (:77-127).
 * out_info: [0] total_force_on_human, [1] task_success flag, [2] tool_force, [3] tool_force_on_human,
 *           [4] reward_distance, [5] reward_action, [6] new_contact_points, [7] preferences_score */
static void bb_finish_step(const Model* m, double* env, const Contact* contacts, int nc, double raw_sq, double* obs,
                           double* reward, double* out_info) {
    const AvgModelHeader* h = m->h;
    const float* tf = h->task_f;
    double dt = h->dt;
    Kin k; fk(m, env, &k);
    double tool_force = 0, tool_force_on_human = 0, total_force_on_human = 0;
    int new_contact_points = 0;
    for (int c = 0; c < nc; ++c) {
        const AvgShape* sa = &m->shape[contacts[c].sa]; const AvgShape* sb = &m->shape[contacts[c].sb];
        double force = contacts[c].lambda_n / dt;
        int a_tool = sa->ref_body == AVG_REF_TOOL, b_tool = sb->ref_body == AVG_REF_TOOL;
        int a_hum = sa->ref_body == AVG_REF_HUMAN, b_hum = sb->ref_body == AVG_REF_HUMAN;
        int a_rob = sa->ref_body == AVG_REF_ROBOT, b_rob = sb->ref_body == AVG_REF_ROBOT;
        if (a_tool || b_tool) tool_force += force;                                         /* :83-85 */
        if ((a_rob && b_hum) || (b_rob && a_hum)) total_force_on_human += force;           /* :90-91 */
        if ((a_tool && b_hum) || (b_tool && a_hum)) {                                      /* :92-125 */
            total_force_on_human += force;
            int link_tool = a_tool ? sa->ref_link : sb->ref_link;
            int link_hum = a_tool ? sb->ref_link : sa->ref_link;
            v3 pos_h = a_tool ? contacts[c].pb : contacts[c].pa;                           /* positionOnB, B = human */
            if (link_tool == 1) {
                tool_force_on_human += force;
                if (link_hum < 0) continue;                                                /* :100-101 (human base) */
                for (int t = 0; t < h->n_target; ++t) {
                    int w = AVG_E_TARGET_MASK + (t >> 5);
                    uint32_t bits = (uint32_t)env[w];
                    if (!(bits & (1u << (t & 31)))) continue;
                    if (vnorm(vsub(pos_h, bb_target_world(m, &k, t))) < tf[AVG_TF_TARGET_RADIUS]) {
                        new_contact_points += 1;
                        env[AVG_E_TASK_SUCCESS] += 1;
                        env[w] = (double)(bits & ~(1u << (t & 31)));
                    }
                }
            }
        }
    }
    v3 tip; quat tq; frame_pose(m, &k, AVG_F_TOOL_TIP, &tip, &tq);
    int tb = m->frame[AVG_F_TOOL_TIP].body;
    const double* tv = env + AVG_E_QD + m->body[tb].dof;
    double ee_vel = vnorm(vadd(V(tv[0], tv[1], tv[2]), vcross(V(tv[3], tv[4], tv[5]), vsub(tip, k.p[tb]))));   /* :54 */
    bb_get_obs(m, env, tool_force, total_force_on_human, tool_force_on_human, obs);
    double pref = tf[AVG_TF_C_V] * (-ee_vel) + tf[AVG_TF_C_F] * (-(total_force_on_human - tool_force_on_human))
                + tf[AVG_TF_C_HF] * (tool_force_on_human < tf[AVG_TF_FORCE_CAP] ? 0.0 : -tool_force_on_human);   /* env.py:412-448 */
    double reward_distance = -bb_closest_tool_human(m, &k);                               /* :61 */
    double reward_action = -raw_sq;
    *reward = tf[AVG_TF_DISTANCE_W] * reward_distance + tf[AVG_TF_ACTION_W] * reward_action
            + tf[AVG_TF_SCRATCH_W] * new_contact_points + pref;                            /* :65 */
    env[AVG_E_EPISODE_RETURN] += *reward;
    if (out_info) {
        out_info[0] = total_force_on_human; out_info[1] = env[AVG_E_TASK_SUCCESS] >= tf[AVG_TF_SUCCESS_THR] ? 1.0 : 0.0;
        out_info[2] = tool_force; out_info[3] = tool_force_on_human; out_info[4] = reward_distance;
        out_info[5] = reward_action; out_info[6] = new_contact_points; out_info[7] = pref;
    }
}


/* ---- Feeding / Drinking (feeding.py, drinking.py) ------------------------------------------------------------- */
/* update_targets, feeding.py:345-349: the mouth = head link 27 frame o mouth_pos */
static v3 fd_mouth(const Model* m, const Kin* k) {
    v3 p; quat q; frame_pose(m, k, AVG_F_HEAD, &p, &q);
    const float* tf = m->h->task_f;
    return vadd(p, qrot(q, V(tf[AVG_TF_MOUTH], tf[AVG_TF_MOUTH + 1], tf[AVG_TF_MOUTH + 2])));
}
/* feeding.py:123-142, drinking.py:138-157: robot obs[25] (+ human obs[23]) */
static void fd_get_obs(const Model* m, const double* env, double tool_force_on_human, double robot_force_on_human, double* obs) {
    const AvgModelHeader* h = m->h;
    Kin k; fk(m, env, &k);
    v3 torso, tool, head, chest; quat tq, hq, dummy;
    frame_pose(m, &k, AVG_F_TORSO, &torso, &dummy);
    frame_pose(m, &k, AVG_F_TOOL_TIP, &tool, &tq);
    frame_pose(m, &k, AVG_F_HEAD, &head, &hq);
    frame_pose(m, &k, AVG_F_CHEST, &chest, &dummy);
    v3 tgt = fd_mouth(m, &k);
    int o = 0; v3 t;
    t = vsub(tool, torso); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
    obs[o++] = tq.x; obs[o++] = tq.y; obs[o++] = tq.z; obs[o++] = tq.w;
    t = vsub(tool, tgt); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
    for (int i = 0; i < h->n_jdof; ++i) if (m->dof[i].action >= 0 && m->dof[i].action < h->n_action_robot) obs[o++] = env[AVG_E_Q + m->body[m->dof[i].body].qidx];
    t = vsub(head, torso); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
    obs[o++] = hq.x; obs[o++] = hq.y; obs[o++] = hq.z; obs[o++] = hq.w;
    obs[o++] = tool_force_on_human;
    if (h->human_control) {                                   /* positions relative to human link 3, head joints 24..27 */
        t = vsub(tool, chest); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
        obs[o++] = tq.x; obs[o++] = tq.y; obs[o++] = tq.z; obs[o++] = tq.w;
        t = vsub(tool, tgt); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
        double hq4[4] = {0, 0, 0, 0};
        for (int i = 0; i < h->n_jdof; ++i) if (m->dof[i].human_slot >= 0 && m->dof[i].human_slot < 4) hq4[m->dof[i].human_slot] = env[AVG_E_Q + m->body[m->dof[i].body].qidx];
        for (int i = 0; i < 4; ++i) obs[o++] = hq4[i];
        t = vsub(head, chest); obs[o++] = t.x; obs[o++] = t.y; obs[o++] = t.z;
        obs[o++] = hq.x; obs[o++] = hq.y; obs[o++] = hq.z; obs[o++] = hq.w;
        obs[o++] = robot_force_on_human; obs[o++] = tool_force_on_human;
    }
}
/* btQuaternion::getEulerZYX roll (p.getEulerFromQuaternion(q)[0]) [UPSTREAM-BULLET] */
static double quat_roll(quat q) {
    double sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
    double sarg = -2.0 * (q.x * q.z - q.w * q.y) / (sqx + sqy + sqz + sqw);
    if (sarg <= -0.99999 || sarg >= 0.99999) return 0.0;
    return atan2(2.0 * (q.y * q.z + q.w * q.x), sqw - sqx - sqy + sqz);
}
/* The part of FeedingEnv.step / DrinkingEnv.step after take_step (feeding.py:56-80, drinking.py:57-83) with get_total_force
 * (feeding.py:83-90) and get_food_rewards / get_water_rewards (feeding.py:92-121, drinking.py:95-136).
 * out_info: [0] total_force_on_human, [1] task_success flag, [2] robot_force_on_human, [3] tool_force_on_human,
 *           [4] reward_distance, [5] reward_action, [6] reward_food (+ tilt term for Drinking in [8] of the CUDA taps), [7] preferences_score */
static void fd_finish_step(const Model* m, double* env, double* part, const Contact* contacts, int nc, double raw_sq, double* obs,
                           double* reward, double* out_info) {
    const AvgModelHeader* h = m->h;
    const float* tf = h->task_f;
    const double dt = h->dt;
    const int drinking = h->task == AVG_TASK_DRINKING;
    Kin k; fk(m, env, &k);
    double robot_force_on_human = 0, tool_force_on_human = 0;
    for (int c = 0; c < nc; ++c) {
        const AvgShape* sa = &m->shape[contacts[c].sa]; const AvgShape* sb = &m->shape[contacts[c].sb];
        double force = contacts[c].lambda_n / dt;
        int a_tool = sa->ref_body == AVG_REF_TOOL, b_tool = sb->ref_body == AVG_REF_TOOL;
        int a_hum = sa->ref_body == AVG_REF_HUMAN, b_hum = sb->ref_body == AVG_REF_HUMAN;
        int a_rob = sa->ref_body == AVG_REF_ROBOT, b_rob = sb->ref_body == AVG_REF_ROBOT;
        if ((a_rob && b_hum) || (b_rob && a_hum)) robot_force_on_human += force;           /* feeding.py:86-87 */
        if ((a_tool && b_hum) || (b_tool && a_hum)) tool_force_on_human += force;          /* :88-89 */
    }
    const v3 mouth = fd_mouth(m, &k);
    v3 tool; quat tq; frame_pose(m, &k, AVG_F_TOOL_TIP, &tool, &tq);
    /* particles */
    double food_reward = 0, hit_reward = 0, mouth_vel_sum = 0;
    part[AVG_P_EV_EAT] = part[AVG_P_EV_EAT + 1] = part[AVG_P_EV_SPILL] = part[AVG_P_EV_SPILL + 1] = part[AVG_P_EV_HIT] = part[AVG_P_EV_HIT + 1] = 0;
    v3 top = V(0, 0, 0), bottom = V(0, 0, 0), cup_p = tool; quat cup_q = tq;
    if (drinking) {                                           /* drinking.py:97-100: cup frame = base o ([0, 0.06, 0], rotX 90 deg) */
        quat rx = qaxis(V(1, 0, 0), 1.5707963267948966);
        cup_p = vadd(tool, qrot(tq, V(0, 0.06, 0))); cup_q = qnormalize(qmul(tq, rx));
        top = vadd(cup_p, qrot(cup_q, V(0, 0, tf[AVG_TF_CUP_TOP])));
        bottom = vadd(cup_p, qrot(cup_q, V(0, 0, tf[AVG_TF_CUP_BOTTOM])));
    }
    for (int p = 0; p < h->n_particle; ++p) {
        if (!p_alive(part, p)) continue;
        v3 x = V(P_X(part, p, 0), P_X(part, p, 1), P_X(part, p, 2));
        if (drinking) {                                       /* util.points_in_cylinder(top, bottom, 0.05, x), util.py:107-110 */
            v3 vec = vsub(bottom, top);
            double cst = tf[AVG_TF_CUP_RADIUS] * vnorm(vec);
            int inside = vdot(vsub(x, top), vec) >= 0 && vdot(vsub(x, bottom), vec) <= 0 && vnorm(vcross(vsub(x, top), vec)) <= cst;
            if (inside) continue;
        }
        double dist = vnorm(vsub(mouth, x));
        if (dist < tf[AVG_TF_EAT_RADIUS]) {                   /* feeding.py:102-110, drinking.py:114-122 */
            food_reward += tf[AVG_TF_EAT_REWARD];
            env[AVG_E_TASK_SUCCESS] += 1;
            /* Feeding reads the velocity before the teleport; Drinking after resetBasePositionAndOrientation, which zeroes it
             * (drinking.py:118-119) [UPSTREAM-BULLET] */
            if (!drinking) mouth_vel_sum += vnorm(V(P_V(part, p, 0), P_V(part, p, 1), P_V(part, p, 2)));
            p_clrbit(part, AVG_P_ALIVE, p); p_setbit(part, AVG_P_EV_EAT, p);
            continue;
        }
        int spill = x.z < tf[AVG_TF_Z_MIN] || (!drinking && p_getbit(part, AVG_P_TOUCH_SPILL, p));   /* feeding.py:111, drinking.py:124 */
        if (spill) {
            food_reward += tf[AVG_TF_SPILL_REWARD];
            p_clrbit(part, AVG_P_ALIVE, p); p_setbit(part, AVG_P_EV_SPILL, p);
            continue;
        }
        if (p_getbit(part, AVG_P_TOUCH_HUMAN, p)) {
            if (drinking) {                                   /* drinking.py:131-134: removed on the first touch */
                hit_reward -= 1; p_clrbit(part, AVG_P_ALIVE, p); p_setbit(part, AVG_P_EV_HIT, p);
            } else if (!p_getbit(part, AVG_P_HIT, p)) {       /* feeding.py:116-119: penalised once, stays in play */
                hit_reward -= 1; p_setbit(part, AVG_P_HIT, p); p_setbit(part, AVG_P_EV_HIT, p);
            }
        }
    }
    int tb = m->frame[AVG_F_TOOL_TIP].body;
    const double* tv = env + AVG_E_QD + m->body[tb].dof;
    double ee_vel = vnorm(V(tv[0], tv[1], tv[2]));                                         /* getBaseVelocity(spoon)[0], feeding.py:59 */
    fd_get_obs(m, env, tool_force_on_human, robot_force_on_human, obs);
    /* human_preferences, env.py:412-448 with total_force_on_human = robot force, tool_force_at_target = tool force (feeding.py:63) */
    double pref = tf[AVG_TF_C_V] * (-ee_vel) + tf[AVG_TF_C_F] * (-robot_force_on_human)
                + tf[AVG_TF_C_HF] * (tool_force_on_human < tf[AVG_TF_FORCE_CAP] ? 0.0 : -tool_force_on_human)
                + tf[AVG_TF_C_FD] * hit_reward + tf[AVG_TF_C_FDV] * (-mouth_vel_sum);
    double reward_distance, reward_tilt = 0;
    if (drinking) {
        reward_distance = -vnorm(vsub(mouth, top));                                        /* drinking.py:68 */
        double roll = quat_roll(cup_q);
        reward_tilt = tf[AVG_TF_TILT_SIGN] > 0 ? -fabs(roll + 1.5707963267948966) : -fabs(roll - 1.5707963267948966);   /* :72 */
    } else reward_distance = -vnorm(vsub(mouth, tool));                                    /* feeding.py:68 */
    double reward_action = -raw_sq;
    *reward = tf[AVG_TF_DISTANCE_W] * reward_distance + tf[AVG_TF_ACTION_W] * reward_action + tf[AVG_TF_TILT_W] * reward_tilt
            + tf[AVG_TF_FOOD_W] * food_reward + pref;                                      /* feeding.py:71, drinking.py:74 */
    env[AVG_E_EPISODE_RETURN] += *reward;
    env[AVG_E_TARGET_POS] = mouth.x; env[AVG_E_TARGET_POS + 1] = mouth.y; env[AVG_E_TARGET_POS + 2] = mouth.z;
    if (out_info) {
        out_info[0] = robot_force_on_human + tool_force_on_human;
        out_info[1] = env[AVG_E_TASK_SUCCESS] >= tf[AVG_TF_SUCCESS_THR] ? 1.0 : 0.0;
        out_info[2] = robot_force_on_human; out_info[3] = tool_force_on_human; out_info[4] = reward_distance;
        out_info[5] = reward_action; out_info[6] = food_reward + tf[AVG_TF_TILT_W] * reward_tilt; out_info[7] = pref;
    }
}

/* `n` calls of p.stepSimulation() without actions or per-step hooks: the settle loop of reset() (feeding.py:318-320,
 * drinking.py:320-322: "Drop food in the spoon"). */
int avg_oracle_settle(const void* blob, double* env, double* part, int n) {
    Model m; if (model_open(blob, &m)) return -1;
    Contact contacts[MAXC]; int nc = 0;
    const int n_internal = m.h->n_internal > 0 ? m.h->n_internal : 1;
    for (int s = 0; s < n * n_internal; ++s) substep(&m, env, part, contacts, &nc);
    return 0;
}
/* particle contacts at the current state (no stepping): rows of 8 doubles (p, q, shape_b, n(3), dist, 0) */
int avg_oracle_particle_collide(const void* blob, const double* env, const double* part, double* out, int* n_out) {
    Model m; if (model_open(blob, &m)) return -1;
    Kin k; fk(&m, env, &k);
    PContact* pcs = (PContact*)malloc(sizeof(PContact) * AVG_MAX_PCONTACT);
    int overflow = 0;
    int n = particle_collide(&m, &k, part, pcs, &overflow);
    for (int c = 0; c < n; ++c) {
        double* o = out + 8 * c;
        o[0] = pcs[c].p; o[1] = pcs[c].q; o[2] = pcs[c].shape_b; o[3] = pcs[c].n.x; o[4] = pcs[c].n.y; o[5] = pcs[c].n.z; o[6] = pcs[c].dist; o[7] = 0;
    }
    *n_out = n; free(pcs);
    return overflow;
}

/* exported ---------------------------------------------------------------------------------------------------- */
int avg_oracle_sizes(int* sizes) {
    sizes[0] = (int)sizeof(AvgModelHeader); sizes[1] = (int)sizeof(AvgBody); sizes[2] = (int)sizeof(AvgDof);
    sizes[3] = (int)sizeof(AvgShape); sizes[4] = (int)sizeof(AvgFrame); sizes[5] = (int)sizeof(AvgContact);
    sizes[6] = (int)sizeof(AvgResetTable);
    return 7;
}

/* initial observation, scratch_itch.py:268  (_get_obs([0],[0,0]) after generate_target) */
int avg_oracle_reset_obs(const void* blob, double* env, double* obs) {
    Model m; if (model_open(blob, &m)) return -1;
    if (m.h->task == AVG_TASK_BED_BATHING) { bb_get_obs(&m, env, 0, 0, 0, obs); return 0; }     /* bed_bathing.py:350 */
    if (m.h->task == AVG_TASK_FEEDING || m.h->task == AVG_TASK_DRINKING) { fd_get_obs(&m, env, 0, 0, obs); return 0; }   /* feeding.py:325 */
    update_target(&m, env);
    get_obs(&m, env, 0, 0, 0, obs);
    return 0;
}

/* One env.step(action): take_step (env.py:274-351) + ScratchItchEnv.step (scratch_itch.py:53-82).
 * out_info: [0] total_force_on_human, [1] task_success flag, [2] tool_force, [3] tool_force_at_target,
 *           [4] reward_distance, [5] reward_action, [6] reward_force_scratch, [7] preferences_score
 * contacts_out: AvgContactD records of the last sub-step (10 doubles each + 2 ints packed as doubles):
 *           [sa, sb, pa(3), pb(3), n(3), dist, force] = 13 doubles */
static void fd_finish_step(const Model* m, double* env, double* part, const Contact* contacts, int nc, double raw_sq, double* obs,
                           double* reward, double* out_info);
static void fd_get_obs(const Model* m, const double* env, double tool_force_on_human, double robot_force_on_human, double* obs);

int avg_oracle_step_fd(const void* blob, double* env, double* part, const float* action, double* obs, double* reward, double* out_info,
                       double* contacts_out, int* ncontacts_out);
int avg_oracle_step(const void* blob, double* env, const float* action, double* obs, double* reward, double* out_info,
                    double* contacts_out, int* ncontacts_out) {
    return avg_oracle_step_fd(blob, env, 0, action, obs, reward, out_info, contacts_out, ncontacts_out);
}
/* `part`: the particle record of Feeding / Drinking (AVG_P_STRIDE doubles, masks as numbers), NULL for the other tasks */
int avg_oracle_step_fd(const void* blob, double* env, double* part, const float* action, double* obs, double* reward, double* out_info,
                       double* contacts_out, int* ncontacts_out) {
    Model m; if (model_open(blob, &m)) return -1;
    if (m.h->n_particle > 0 && !part) return -3;
    const AvgModelHeader* h = m.h;
    int na = h->n_action_robot + h->n_action_human;
    float act[64];
    double raw_sq = 0;
    for (int i = 0; i < na; ++i) {
        float a = action[i];
        raw_sq += (double)a * (double)a;                      /* reward_action uses the raw action, scratch_itch.py:64 */
        if (a < -1.0f) a = -1.0f; if (a > 1.0f) a = 1.0f;      /* env.py:275 */
        act[i] = a * h->action_scale;                         /* env.py:280 (float32 arithmetic like numpy) */
    }
    int human_active = h->human_control || env[AVG_E_TREMOR_ON] != 0.0;      /* env.py:307 */
    /* robot targets, env.py:320-326 */
    double ar[32], rpos[32]; int rdof[32], nrob = 0;
    for (int i = 0; i < h->n_jdof; ++i) if (m.dof[i].action >= 0 && m.dof[i].action < h->n_action_robot) {
        rdof[nrob] = i; ar[nrob] = act[m.dof[i].action]; rpos[nrob] = env[AVG_E_Q + m.body[m.dof[i].body].qidx]; nrob++;
    }
    /* human targets, env.py:307-318 */
    double ah[10], hpos[10], hlo[10], hhi[10]; int hdof[10];
    for (int s = 0; s < 10; ++s) { ah[s] = 0; hpos[s] = 0; hlo[s] = 0; hhi[s] = 0; hdof[s] = -1; }
    if (human_active) {
        for (int i = 0; i < h->n_jdof; ++i) if (m.dof[i].human_slot >= 0) {
            int s = m.dof[i].human_slot; hdof[s] = i;
            hpos[s] = env[AVG_E_Q + m.body[m.dof[i].body].qidx];
            hlo[s] = limit_lo(&m.dof[i], env); hhi[s] = limit_hi(&m.dof[i], env);
        }
        if (h->human_control) for (int s = 0; s < 10 && s < h->n_action_human; ++s) ah[s] = act[h->n_action_robot + s];
    }
    for (int f = 0; f < h->substeps; ++f) {
        for (int j = 0; j < nrob; ++j) {
            const AvgDof* D = &m.dof[rdof[j]];
            if (rpos[j] + ar[j] < D->rep_lower) ar[j] = 0;
            if (rpos[j] + ar[j] > D->rep_upper) ar[j] = 0;
            rpos[j] += ar[j];
        }
        if (human_active) {
            for (int s = 0; s < 10; ++s) {
                if (hpos[s] + ah[s] < hlo[s]) ah[s] = 0;
                if (hpos[s] + ah[s] > hhi[s]) ah[s] = 0;
            }
            if (env[AVG_E_TREMOR_ON] != 0.0) {                 /* env.py:330-332 */
                double sgn = (((int)env[AVG_E_ITERATION]) % 2 == 0) ? 1.0 : -1.0;
                for (int s = 0; s < 10; ++s) {
                    hpos[s] = env[AVG_E_TARGET_H + s] + env[AVG_E_TREMOR + s] * sgn;
                    env[AVG_E_TARGET_H + s] += ah[s];
                }
            }
            for (int s = 0; s < 10; ++s) hpos[s] += ah[s];
        }
    }
    for (int j = 0; j < nrob; ++j) env[AVG_E_MTARGET + rdof[j]] = rpos[j];       /* env.py:335 */
    if (human_active) {
        for (int s = 0; s < 10; ++s) if (hdof[s] >= 0) env[AVG_E_MTARGET + hdof[s]] = hpos[s];   /* env.py:337 */
        env[AVG_E_HUMAN_KP] = h->task_f[AVG_TF_HUMAN_KP_ACTIVE];
    }
    Contact contacts[MAXC]; int nc = 0;
    const double dt = h->dt;
    const int n_internal = h->n_internal > 0 ? h->n_internal : 1;
    for (int f = 0; f < h->substeps; ++f) {                    /* env.py:341-349 */
        for (int i = 0; i < n_internal; ++i) substep(&m, env, part, contacts, &nc);     /* p.stepSimulation = numSubSteps internal steps */
        enforce_realistic_limits(&m, env);                     /* env.py:343-344, human_control only */
        enforce_hard_limits(&m, env);
        if (h->task == AVG_TASK_SCRATCH_ITCH) update_target(&m, env);
    }
    env[AVG_E_ITERATION] += 1;                                 /* env.py:351 */
    if (h->task == AVG_TASK_BED_BATHING) {
        bb_finish_step(&m, env, contacts, nc, raw_sq, obs, reward, out_info);
        goto report_contacts;
    }
    if (h->task == AVG_TASK_FEEDING || h->task == AVG_TASK_DRINKING) {
        fd_finish_step(&m, env, part, contacts, nc, raw_sq, obs, reward, out_info);
        goto report_contacts;
    }

    /* get_total_force, scratch_itch.py:84-102 */
    double total_force_on_human = 0, tool_force = 0, tool_force_at_target = 0;
    int have_tcp = 0; v3 tcp = V(0, 0, 0);
    v3 tgt = V(env[AVG_E_TARGET_POS], env[AVG_E_TARGET_POS + 1], env[AVG_E_TARGET_POS + 2]);
    for (int c = 0; c < nc; ++c) {
        const AvgShape* sa = &m.shape[contacts[c].sa]; const AvgShape* sb = &m.shape[contacts[c].sb];
        double force = contacts[c].lambda_n / dt;
        int a_tool = sa->ref_body == AVG_REF_TOOL, b_tool = sb->ref_body == AVG_REF_TOOL;
        int a_hum = sa->ref_body == AVG_REF_HUMAN, b_hum = sb->ref_body == AVG_REF_HUMAN;
        int a_rob = sa->ref_body == AVG_REF_ROBOT, b_rob = sb->ref_body == AVG_REF_ROBOT;
        if (a_tool || b_tool) tool_force += force;
        if ((a_tool && b_hum) || (b_tool && a_hum)) {
            total_force_on_human += force;
            int link_tool = a_tool ? sa->ref_link : sb->ref_link;
            v3 pos_h = a_tool ? contacts[c].pb : contacts[c].pa;             /* positionOnB with B = human */
            if (link_tool == 0 || link_tool == 1) {
                if (vnorm(vsub(pos_h, tgt)) < h->task_f[AVG_TF_TARGET_RADIUS]) { tool_force_at_target += force; tcp = pos_h; have_tcp = 1; }
            }
        }
        if ((a_rob && b_hum) || (b_rob && a_hum)) total_force_on_human += force;
    }
    /* end effector velocity: linear velocity of tool link 1 COM, scratch_itch.py:54 */
    Kin k; fk(&m, env, &k);
    v3 tip; quat tq; frame_pose(&m, &k, AVG_F_TOOL_TIP, &tip, &tq);
    int tb = m.frame[AVG_F_TOOL_TIP].body;
    const double* tv = env + AVG_E_QD + m.body[tb].dof;
    v3 vt = vadd(V(tv[0], tv[1], tv[2]), vcross(V(tv[3], tv[4], tv[5]), vsub(tip, k.p[tb])));
    double ee_vel = vnorm(vt);
    get_obs(&m, env, tool_force, total_force_on_human, tool_force_at_target, obs);
    /* human_preferences, env.py:412-448 (scratch itch terms) */
    const float* tf = h->task_f;
    double pref = tf[AVG_TF_C_V] * (-ee_vel) + tf[AVG_TF_C_F] * (-(total_force_on_human - tool_force_at_target))
                + tf[AVG_TF_C_HF] * (tool_force_at_target < tf[AVG_TF_FORCE_CAP] ? 0.0 : -tool_force_at_target);
    double reward_distance = -vnorm(vsub(tgt, tip));
    double reward_action = -raw_sq;
    double reward_force_scratch = 0;
    v3 prev = V(env[AVG_E_PREV_CONTACT], env[AVG_E_PREV_CONTACT + 1], env[AVG_E_PREV_CONTACT + 2]);
    if (have_tcp && vnorm(vsub(tcp, prev)) > tf[AVG_TF_SCRATCH_MOVE] && tool_force_at_target < tf[AVG_TF_FORCE_CAP]) {
        reward_force_scratch = tool_force_at_target;
        env[AVG_E_PREV_CONTACT] = tcp.x; env[AVG_E_PREV_CONTACT + 1] = tcp.y; env[AVG_E_PREV_CONTACT + 2] = tcp.z;
        env[AVG_E_TASK_SUCCESS] += 1;
    }
    *reward = tf[AVG_TF_DISTANCE_W] * reward_distance + tf[AVG_TF_ACTION_W] * reward_action
            + tf[AVG_TF_TOOL_FORCE_W] * tool_force_at_target + tf[AVG_TF_SCRATCH_W] * reward_force_scratch + pref;
    env[AVG_E_EPISODE_RETURN] += *reward;
    if (out_info) {
        out_info[0] = total_force_on_human; out_info[1] = env[AVG_E_TASK_SUCCESS] >= tf[AVG_TF_SUCCESS_THR] ? 1.0 : 0.0;
        out_info[2] = tool_force; out_info[3] = tool_force_at_target; out_info[4] = reward_distance;
        out_info[5] = reward_action; out_info[6] = reward_force_scratch; out_info[7] = pref;
    }
report_contacts:
    if (contacts_out && ncontacts_out) {
        *ncontacts_out = nc;
        for (int c = 0; c < nc; ++c) {
            double* o = contacts_out + 13 * c;
            o[0] = contacts[c].sa; o[1] = contacts[c].sb;
            o[2] = contacts[c].pa.x; o[3] = contacts[c].pa.y; o[4] = contacts[c].pa.z;
            o[5] = contacts[c].pb.x; o[6] = contacts[c].pb.y; o[7] = contacts[c].pb.z;
            o[8] = contacts[c].n.x; o[9] = contacts[c].n.y; o[10] = contacts[c].n.z;
            o[11] = contacts[c].dist; o[12] = contacts[c].lambda_n / dt;
        }
    }
    return 0;
}

/* --- unit-test hooks (tests/test_oracle_*.py) ------------------------------------------------------------------ */
/* forward kinematics of frame f -> pos(3), quat(4) */
int avg_oracle_frame(const void* blob, const double* env, int f, double* out) {
    Model m; if (model_open(blob, &m)) return -1;
    Kin k; fk(&m, env, &k);
    v3 p; quat q; frame_pose(&m, &k, f, &p, &q);
    out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
    return 0;
}
int avg_oracle_body_pose(const void* blob, const double* env, int b, double* out) {
    Model m; if (model_open(blob, &m)) return -1;
    Kin k; fk(&m, env, &k);
    out[0] = k.p[b].x; out[1] = k.p[b].y; out[2] = k.p[b].z; out[3] = k.q[b].x; out[4] = k.q[b].y; out[5] = k.q[b].z; out[6] = k.q[b].w;
    return 0;
}
/* unconstrained accelerations (ABA) and, column by column, M^-1 (n_dof x n_dof, row-major) */
int avg_oracle_dynamics(const void* blob, const double* env, double* qdd, double* minv) {
    Model m; if (model_open(blob, &m)) return -1;
    Kin k; fk(&m, env, &k);
    Dyn* d = (Dyn*)malloc(sizeof(Dyn));
    aba(&m, &k, env, d, qdd);
    int nd = m.h->n_dof;
    if (minv) for (int j = 0; j < nd; ++j) {
        double tau[MAXD], col[MAXD]; memset(tau, 0, sizeof(tau)); tau[j] = 1.0;
        aba_delta(&m, d, tau, col);
        for (int i = 0; i < nd; ++i) minv[i * nd + j] = col[i];
    }
    free(d);
    return 0;
}
/* contacts at the current configuration (no stepping) */
int avg_oracle_collide(const void* blob, const double* env, double* contacts_out, int* ncontacts_out) {
    Model m; if (model_open(blob, &m)) return -1;
    Kin k; fk(&m, env, &k);
    Contact contacts[MAXC]; int overflow = 0;
    int nc = collide(&m, &k, contacts, &overflow);
    *ncontacts_out = nc;
    for (int c = 0; c < nc; ++c) {
        double* o = contacts_out + 13 * c;
        o[0] = contacts[c].sa; o[1] = contacts[c].sb;
        o[2] = contacts[c].pa.x; o[3] = contacts[c].pa.y; o[4] = contacts[c].pa.z;
        o[5] = contacts[c].pb.x; o[6] = contacts[c].pb.y; o[7] = contacts[c].pb.z;
        o[8] = contacts[c].n.x; o[9] = contacts[c].n.y; o[10] = contacts[c].n.z;
        o[11] = contacts[c].dist; o[12] = 0;
    }
    return overflow;
}
/* arm-limit classifier logit for raw joint angles (tz, tx, ty, qe) */
int avg_oracle_arm_limit(const void* blob, const double* q, double* logit) {
    Model m; if (model_open(blob, &m)) return -1;
    if (!m.mlp) return -2;
    const double twopi = 2.0 * 3.14159265358979323846;
    double x[4] = {pymod(-q[0] + twopi, twopi), pymod(q[1] + twopi, twopi), -q[2], pymod(-q[3] + twopi, twopi)};
    *logit = arm_limit_logit(m.mlp, x);
    return 0;
}
/* closest points between two shapes given explicit world poses (GJK unit tests): pose = pos(3)+quat(4) */
int avg_oracle_shape_pair(const void* blob, int sa, const double* pose_a, int sb, const double* pose_b, double thr, double* out) {
    Model m; if (model_open(blob, &m)) return -1;
    WShape A, B;
    A.s = &m.shape[sa]; A.verts = m.vert + 4 * A.s->vert_off; A.planes = m.plane + 4 * A.s->plane_off;
    B.s = &m.shape[sb]; B.verts = m.vert + 4 * B.s->vert_off; B.planes = m.plane + 4 * B.s->plane_off;
    quat qa = {pose_a[3], pose_a[4], pose_a[5], pose_a[6]}, qb = {pose_b[3], pose_b[4], pose_b[5], pose_b[6]};
    A.p = V(pose_a[0], pose_a[1], pose_a[2]); A.R = qmat(qnormalize(qa));
    B.p = V(pose_b[0], pose_b[1], pose_b[2]); B.R = qmat(qnormalize(qb));
    Contact c;
    int hit = narrowphase(&A, &B, thr, &c);
    if (hit) {
        out[0] = c.pa.x; out[1] = c.pa.y; out[2] = c.pa.z; out[3] = c.pb.x; out[4] = c.pb.y; out[5] = c.pb.z;
        out[6] = c.n.x; out[7] = c.n.y; out[8] = c.n.z; out[9] = c.dist;
    }
    return hit;
}
