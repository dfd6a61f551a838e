// avg_kernels.cu — sm_100a kernels of the batched Assistive-Gym simulator (ScratchItch path).
//
// One warp owns one environment.  An env.step() (reference AssistiveEnv.take_step, env.py:274-351, followed by
// ScratchItchEnv.step, scratch_itch.py:53-128) is a short pipeline of kernels on one stream:
//
//   avg_prologue_kernel   action -> limit-masked motor targets (env.py:275-337)
//   frame_skip x {                                               (what the reference delegates to p.stepSimulation)
//     avg_collide_kernel  forward kinematics, broadphase, GJK narrowphase              -> contact list
//     avg_dynamics_kernel forward kinematics, mass matrix, bias forces, M^-1, qd*,
//                         constraint rows (motors, limits, weld, contacts) + M^-1 J^T   -> row arena
//     avg_solve_kernel    projected Gauss-Seidel, integration, human hard limits (env.py:389-410)
//   }
//   avg_epilogue_kernel   forces, reward, observation, info (scratch_itch.py:53-128)
//
// The phases are separate kernels on purpose: fused, the step was > 400 KB of SASS, far beyond the 32 KB L1.5
// instruction cache, and ncu attributed most stalls to instruction fetch (profiles/).  Split, every kernel has a
// small footprint and its own register / shared-memory budget (the solver runs at ~3x the occupancy of the fused
// kernel), and the hand-off through HBM costs ~10 KB per env and sub-step against a ~6.5 TB/s budget.
//
// Lane mapping inside a warp:
//   kinematics/dynamics : lane b = dynamic body b (1-DoF joint bodies first, so lane i is also velocity dof i)
//   solver              : lane d = velocity dof d (delta-velocity in a register); rows staged in shared memory;
//                         J.dv by warp shuffles; accumulated impulses in registers (row r in lane r%32); strict
//                         row order (Gauss-Seidel semantics)
//   collision           : one lane per static shape in the broadphase, one lane per candidate pair in GJK
//
// The math is NOT a transcription of the CPU oracle: forward dynamics is mass-matrix based (composite-rigid-body
// inertia in a common frame, then an explicit inverse that the solver reuses for M^-1 J^T), whereas the oracle
// runs the recursive articulated-body algorithm.  Both are Featherstone algorithms for the same equations.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>
#include "../../include/avg_model.h"
#include "avg_math.cuh"
#include "avg_kernels.h"

#ifndef AVG_WPB_COLLIDE
#define AVG_WPB_COLLIDE 4         /* warps (= environments) per block of the collide kernel */
#endif
#ifndef AVG_WPB_DYN
#define AVG_WPB_DYN 12            /* warps per block of the dynamics kernel */
#endif
#ifndef AVG_OCC_COLLIDE
#define AVG_OCC_COLLIDE (20 / AVG_WPB_COLLIDE)
#endif
#ifndef AVG_OCC_DYN
#define AVG_OCC_DYN (24 / AVG_WPB_DYN)
#endif
#ifndef AVG_OCC_SOLVE
#define AVG_OCC_SOLVE 32
#endif
#ifndef AVG_REDUX
#define AVG_REDUX 2               /* solver: J.dv of dense rows by an integer redux.sync (1: fixed point 2^-22 m/s, 2: scaled to the largest product) instead of a 5-stage float butterfly (0) */
#endif
#ifndef AVG_LIM_SKIP
#define AVG_LIM_SKIP 1            /* solver: sweep only the block slots in which some articulation has an active limit row */
#endif
#ifndef AVG_OCC_NARROW
#define AVG_OCC_NARROW 6          /* blocks of 128 threads per SM the narrowphase kernel is compiled for */
#endif
#ifndef AVG_NARROW_WARPS
#define AVG_NARROW_WARPS (148 * 6)        /* narrowphase: warps a short work queue is spread over (measured: 12 / 24 per SM cost 5-15 % on in-contact 8192-env rollouts) */
#endif
#ifndef AVG_WELD_BATCH
#define AVG_WELD_BATCH 1          /* dynamics kernel: the six weld rows built together with one packed reduction (0: row by row) */
#endif
#ifndef AVG_GJK_SHRINK
#define AVG_GJK_SHRINK 1e-6f      /* float32 GJK stops when |v|^2 fails to shrink by this relative amount */
#endif
#ifndef AVG_FUSE_DEFAULT
#define AVG_FUSE_DEFAULT 2        /* dynamics + solve as one kernel: 0 never, 1 always, 2 for launches of <= AVG_FUSE_MAX environments (AVG_FUSE overrides) */
#endif
#ifndef AVG_FUSE_MAX
#define AVG_FUSE_MAX 2048         /* measured on a B200 (staggered episodes): 64 / 1024 / 2048 / 4096 envs (= halves of <= 2048) +4..7 %, ScratchItchJacoHuman / BedBathingPR2
                                     at 4096 envs +11 / +8 %; halves of 4096: -1 %, of 8192: -9 %, of 98304: -22 % (80 registers and no phase alignment
                                     of the 116 KB dynamics code across the warps of an SM) */
#endif
#ifndef AVG_GJK_ITERS
#define AVG_GJK_ITERS 32
#endif

namespace {

constexpr int kWarpsPerBlock = AVG_K_WARPS_PER_BLOCK;
constexpr int kMaxJ = AVG_K_MAXJ;          // 1-DoF joints per environment handled by these kernels
constexpr int kMaxMS = AVG_K_MAXMS;        // moving shapes
constexpr int kCandPerLane = 4;
constexpr int kMaxCand = 32 * kCandPerLane; // broadphase candidates per sub-step (before the bounding-capsule cull)
constexpr int kMaxC = AVG_MAX_CONTACT;
constexpr int kMaxDense = AVG_S_MAXDENSE;  // weld rows + contact normal + friction rows
constexpr int kMaxRows = AVG_MAX_ROWS;
constexpr int kSmDense = 10;               // dense rows staged in shared memory by the solver (weld + 2 contacts); more spill to L1
constexpr int kMaxBlk = 16;                // largest articulation (diagonal block of M^-1) the solver keeps in registers

struct KM {                                // device view of a ModelBlob
    const AvgModelHeader* h;
    const AvgBody* body;
    const AvgDof* dof;
    const AvgShape* shape;
    const float* vert;
    const float* plane;
    const uint32_t* pair;
    const AvgFrame* frame;
    const AvgBpStatic* bps;
    const uint32_t* bpm;
    const float4* bcap;
    const float* mlp;
    const float4* target;                  // BedBathing wiping targets (bed_bathing.py:360-379)
    const float4* caabb;                   // compound children: AABB centre | half extents in the owning body's frame
};

// Section pointers per (handle slot, variant), resolved on the host when a model is uploaded: a warp reads its model
// view with one constant-memory access instead of chasing header offsets through global memory.
__constant__ KM c_models[AVG_K_MAX_HANDLES][AVG_K_MAX_VARIANTS];

__device__ __forceinline__ KM open_model(const unsigned char* blob) {
    KM m;
    m.h = reinterpret_cast<const AvgModelHeader*>(blob);
    m.body = reinterpret_cast<const AvgBody*>(blob + m.h->off_body);
    m.dof = reinterpret_cast<const AvgDof*>(blob + m.h->off_dof);
    m.shape = reinterpret_cast<const AvgShape*>(blob + m.h->off_shape);
    m.vert = reinterpret_cast<const float*>(blob + m.h->off_vert);
    m.plane = reinterpret_cast<const float*>(blob + m.h->off_plane);
    m.pair = reinterpret_cast<const uint32_t*>(blob + m.h->off_pair);
    m.frame = reinterpret_cast<const AvgFrame*>(blob + m.h->off_frame);
    m.bps = reinterpret_cast<const AvgBpStatic*>(blob + m.h->off_bps);
    m.bpm = reinterpret_cast<const uint32_t*>(blob + m.h->off_bpm);
    m.bcap = reinterpret_cast<const float4*>(blob + m.h->off_bcap);
    m.mlp = m.h->n_mlp > 0 ? reinterpret_cast<const float*>(blob + m.h->off_mlp) : nullptr;
    m.target = m.h->n_target > 0 ? reinterpret_cast<const float4*>(blob + m.h->off_target) : nullptr;
    m.caabb = m.h->n_cshape > 0 ? reinterpret_cast<const float4*>(blob + m.h->off_caabb) : nullptr;
    return m;
}

// ---- per-kernel shared memory (one instance per warp).  Member names are shared so the device functions below can
//      be templated on the struct.
struct __align__(16) SmCollide {
    float q[32];                           // position coordinates of the env record
    float bp[32][3]; float bq[32][4];      // body poses
    float sp[kMaxMS][3]; float sR[kMaxMS][9]; float4 saabb[kMaxMS][2];
    float4 scap[kMaxMS][2];                // bounding capsules of the moving shapes, world frame
    uint32_t cand[kMaxCand], cand2[kMaxCand];         // broadphase candidates (unordered) / survivors of the culls in pair order
    uint8_t candf[kMaxCand];                          // per ordered candidate: certificate-cache entry with a direction hint, 254 none, 255 not queued
    float4 sep[3][AVG_S_NSEPMAX];          // separation certificates of the previous sub-step (see collide_warp)
    float sq[kMaxMS][4];                   // world orientation of the moving shapes
    uint8_t near_idx[256];                 // static shapes (index) that overlap the union box of the moving shapes
};
struct __align__(16) SmDyn {
    float env[AVG_ENV_STRIDE];
    float bp[32][3]; float bq[32][4];
    float freeInv[2][12];
    float tmp[32];
    float4 sub[kMaxJ][4];                  // per joint body: inertia about the reference point (10 floats) + bias force (6)
    float c_pa[kMaxC][3], c_pb[kMaxC][3], c_n[kMaxC][3], c_dist[kMaxC];
    int c_sa[kMaxC], c_sb[kMaxC];
};
struct __align__(16) SmSolve {
    float J[kSmDense][32];
    float W[kSmDense][32];
    float4 um[4][kMaxBlk];                 // motor row of dof block_start[g] + t: {target, 1/diag, hi (lo = -hi), diag}; zero when absent
    float4 ul[4][kMaxBlk];                 // limit row, same indexing: {target, 1/diag, diag, sign}; 1/diag = 0 when not violated
    float4 rd[kMaxDense][2];               // dense rows: {target, 1/diag, lo, hi}, {diag, mu, index, normal row}
    float lam[kMaxDense];                  // accumulated impulse of the contact rows (every lane writes the same value and reads back its own write)
};
struct __align__(16) SmEpi {
    float env[AVG_ENV_STRIDE];
    float bp[32][3]; float bq[32][4];
    float obs[64];
    float fr[AVG_F_COUNT][8];              // world poses of the frames of interest (pos, quat), one lane each (frames_warp)
};
struct __align__(16) SmEpiBB {             // BedBathing epilogue: + the tool / human shape lists of the closest-point query
    float env[AVG_ENV_STRIDE];
    float bp[32][3]; float bq[32][4];
    float obs[64];
    float fr[AVG_F_COUNT][8];
    float4 tcap[8][2];                     // bounding capsules of the tool shapes, world frame
    uint8_t tool_idx[8];
    uint8_t hum_idx[120];
};

template <class SM>
__device__ __forceinline__ void body_pose(const SM& s, int b, V3& p, Q4& q) {
    if (b < 0) { p = mk3(0, 0, 0); q = mkq(0, 0, 0, 1); }
    else { p = ld3(s.bp[b]); q = ldq(s.bq[b]); }
}
template <class SM>
__device__ __noinline__ void frame_pose(const KM& m, const SM& s, int f, V3& p, Q4& q) {
    const AvgFrame* F = &m.frame[f];
    V3 bp; Q4 bq;
    body_pose(s, F->body, bp, bq);
    p = bp + qrot(bq, ld3(F->pos));
    q = qnormalize(qmul(bq, ldq(F->quat)));
}

// The frames of interest of the observation / reward code, all at once: lane f evaluates frame f into s.fr (the
// epilogues used to walk them one after the other on lane 0).
template <class SM>
__device__ __forceinline__ void frames_warp(const KM& m, SM& s, int lane) {
    if (lane < AVG_F_COUNT && lane < m.h->n_frame) {
        V3 p; Q4 q;
        frame_pose(m, s, lane, p, q);
        float* f = s.fr[lane];
        f[0] = p.x; f[1] = p.y; f[2] = p.z; f[3] = q.x; f[4] = q.y; f[5] = q.z; f[6] = q.w;
    }
    __syncwarp();
}
template <class SM>
__device__ __forceinline__ void frame_cached(const SM& s, int f, V3& p, Q4& q) {
    const float* v = s.fr[f];
    p = mk3(v[0], v[1], v[2]); q = mkq(v[3], v[4], v[5], v[6]);
}

// ---------------------------------------------------------------------------------------------------------------
// Forward kinematics: every lane builds its body's transform relative to the parent body, then the chain products
// are formed by pointer jumping over the parent array (log2(depth) shuffle rounds instead of a serial walk).
// qpos = position coordinates (AVG_E_Q block of the env record).
// ---------------------------------------------------------------------------------------------------------------
// Env-static bodies (the Feeding bowl: pose drawn per episode, feeding.py:184) take the slots nb .. nb + n_ebody - 1 with
// their pose straight from the record (`ebody` = its AVG_E_EBODY block).
template <class SM>
__device__ void fk_warp(const KM& m, SM& s, const float* qpos, int lane, int nb, const float* ebody = nullptr) {
    V3 p = mk3(0, 0, 0); Q4 q = mkq(0, 0, 0, 1);
    int anc = -1;
    const int ne = ebody ? m.h->n_ebody : 0;
    if (lane >= nb && lane < nb + ne) { const float* e7 = ebody + 7 * (lane - nb); p = ld3(e7); q = ldq(e7 + 3); }
    if (lane < nb) {
        const AvgBody* B = &m.body[lane];
        if (B->jtype == AVG_JOINT_FREE) {
            const float* qq = qpos + B->qidx;
            p = ld3(qq); q = qnormalize(ldq(qq + 3));
        } else {
            V3 ax = ld3(B->axis);
            float qv = qpos[B->qidx];
            Q4 jq = ldq(B->ta_quat);
            V3 jp = ld3(B->ta_pos);
            if (B->jtype == AVG_JOINT_REVOLUTE) jq = qmul(jq, qaxis(ax, qv));
            else jp = jp + qrot(jq, ax * qv);
            p = jp + qrot(jq, ld3(B->tb_pos));
            q = qmul(jq, ldq(B->tb_quat));
            anc = B->parent;
        }
    }
    while (__any_sync(AVG_FULL, anc >= 0)) {
        int src = anc >= 0 ? anc : lane;
        V3 ap = shfl3(p, src); Q4 aq = shflq(q, src);
        int aa = __shfl_sync(AVG_FULL, anc, src);
        if (anc >= 0) { p = ap + qrot(aq, p); q = qmul(aq, q); anc = aa; }
    }
    if (lane < nb + ne) {
        q = qnormalize(q);
        st3(s.bp[lane], p);
        s.bq[lane][0] = q.x; s.bq[lane][1] = q.y; s.bq[lane][2] = q.z; s.bq[lane][3] = q.w;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Collision: support-function shapes, GJK closest points on the cores, margins added afterwards; one point per
// shape pair (see DESIGN.md "narrowphase").  One lane per candidate pair.
// ---------------------------------------------------------------------------------------------------------------
struct WShape {
    const AvgShape* s;
    V3 p;
    float R[9];
    const float* verts;
    const float* planes;
};

template <class SM>
__device__ __forceinline__ void load_wshape(const KM& m, const SM& s, int si, WShape& w) {
    const AvgShape* S = &m.shape[si];
    w.s = S; w.verts = m.vert + 4 * S->vert_off; w.planes = m.plane + 4 * S->plane_off;
    if (si < m.h->n_mshape) {
        w.p = ld3(s.sp[si]);
#pragma unroll
        for (int i = 0; i < 9; ++i) w.R[i] = s.sR[si][i];
    } else {
        w.p = ld3(S->pos);
        M3 r = qmat(ldq(S->quat));
#pragma unroll
        for (int i = 0; i < 9; ++i) w.R[i] = r.m[i];
    }
}

__device__ __noinline__ V3 support(const WShape& w, V3 d) {
    const AvgShape* S = w.s;
    V3 l = mtmul(w.R, d), r;
    switch (S->type) {
    case AVG_SHAPE_CAPSULE: r = mk3(0, 0, l.z > 1e-9f ? S->half[2] : (l.z < -1e-9f ? -S->half[2] : 0.0f)); break;
    case AVG_SHAPE_BOX: {
        float hx = S->half[0] - S->margin, hy = S->half[1] - S->margin, hz = S->half[2] - S->margin;
        r = mk3(l.x >= 0 ? hx : -hx, l.y >= 0 ? hy : -hy, l.z >= 0 ? hz : -hz); break;
    }
    case AVG_SHAPE_CYLINDER: {
        float rc = S->radius - S->margin, hc = S->half[2] - S->margin;
        float n = sqrtf(l.x * l.x + l.y * l.y);
        if (n > 1e-12f) r = mk3(rc * l.x / n, rc * l.y / n, l.z >= 0 ? hc : -hc);
        else r = mk3(rc, 0, l.z >= 0 ? hc : -hc);
        break;
    }
    case AVG_SHAPE_HULL: {
        float bd = -3.0e38f; float4 bv = make_float4(0, 0, 0, 0);
        const float4* v = reinterpret_cast<const float4*>(w.verts);
#pragma unroll 4
        for (int i = 0; i < S->vert_cnt; ++i) {
            const float4 p = __ldg(v + i);
            const float dd = fmaf(l.x, p.x, fmaf(l.y, p.y, l.z * p.z));
            if (dd > bd) { bd = dd; bv = p; }
        }
        r = mk3(bv.x, bv.y, bv.z); break;
    }
    default: r = mk3(0, 0, 0);       // sphere core = its centre
    }
    return w.p + mmul(w.R, r);
}

// Supporting FEATURE of the core in world direction d (|d| = 1): centroid of the core points whose support value is within
// kFeatTol of the maximum, and the feature's size class (1 vertex, 2 edge, >= 4 face).  Penetration witnesses come from
// here: the arg-max vertex of a polytope along one of its own face normals is decided by rounding noise, the centroid of
// the tied vertices is not (same rule as support_feature() of the oracle).  Serial; only the rare SAT path calls it.
constexpr float kFeatTol = 1e-4f;
__device__ __noinline__ V3 support_feature(const WShape& w, V3 d, int& count) {
    const AvgShape* S = w.s;
    const V3 l = mtmul(w.R, d);
    V3 r = mk3(0, 0, 0);
    int cnt = 1;
    switch (S->type) {
    case AVG_SHAPE_CAPSULE:
        if (fabsf(l.z) <= kFeatTol) cnt = 2; else r = mk3(0, 0, l.z > 0 ? S->half[2] : -S->half[2]);
        break;
    case AVG_SHAPE_BOX: {
        const float hx = S->half[0] - S->margin, hy = S->half[1] - S->margin, hz = S->half[2] - S->margin;
        if (fabsf(l.x) <= kFeatTol) cnt *= 2; else r.x = l.x > 0 ? hx : -hx;
        if (fabsf(l.y) <= kFeatTol) cnt *= 2; else r.y = l.y > 0 ? hy : -hy;
        if (fabsf(l.z) <= kFeatTol) cnt *= 2; else r.z = l.z > 0 ? hz : -hz;
        break;
    }
    case AVG_SHAPE_CYLINDER: {
        const float rc = S->radius - S->margin, hc = S->half[2] - S->margin;
        const float n = sqrtf(l.x * l.x + l.y * l.y);
        if (fabsf(l.z) <= kFeatTol) cnt *= 2; else r.z = l.z > 0 ? hc : -hc;
        if (n > kFeatTol) { r.x = rc * l.x / n; r.y = rc * l.y / n; } else cnt *= 8;
        break;
    }
    case AVG_SHAPE_HULL: {
        const float4* v = reinterpret_cast<const float4*>(w.verts);
        float bd = -3.0e38f;
        for (int i = 0; i < S->vert_cnt; ++i) { const float4 p = __ldg(v + i); bd = fmaxf(bd, fmaf(l.x, p.x, fmaf(l.y, p.y, l.z * p.z))); }
        V3 acc = mk3(0, 0, 0); cnt = 0;
        for (int i = 0; i < S->vert_cnt; ++i) {
            const float4 p = __ldg(v + i);
            if (fmaf(l.x, p.x, fmaf(l.y, p.y, l.z * p.z)) >= bd - kFeatTol) { acc = acc + mk3(p.x, p.y, p.z); cnt++; }
        }
        r = acc * (1.0f / (float)max(cnt, 1));
        break;
    }
    default: break;                  // sphere: its centre
    }
    count = cnt;
    return w.p + mmul(w.R, r);
}

// The same for one shape held by EVERY lane of a converged warp (the broadcast copies of sat_served): hull vertices are spread
// over the lanes, maximum / count / centroid by shuffles; every lane returns the result.  The centroid is summed lane by lane and
// then over the lanes instead of in vertex order: the same feature, rounding differences of 1e-7 of a hull's size.
__device__ __noinline__ V3 support_feature_warp(const WShape& w, V3 d, int& count, int lane) {
    const AvgShape* S = w.s;
    if (S->type != AVG_SHAPE_HULL) return support_feature(w, d, count);
    const V3 l = mtmul(w.R, d);
    const float4* v = reinterpret_cast<const float4*>(w.verts);
    const int nv = S->vert_cnt;
    float bd = -3.0e38f;
    for (int i = lane; i < nv; i += 32) { const float4 p = __ldg(v + i); bd = fmaxf(bd, fmaf(l.x, p.x, fmaf(l.y, p.y, l.z * p.z))); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bd = fmaxf(bd, __shfl_xor_sync(AVG_FULL, bd, o));
    V3 acc = mk3(0, 0, 0); int cnt = 0;
    for (int i = lane; i < nv; i += 32) {
        const float4 p = __ldg(v + i);
        if (fmaf(l.x, p.x, fmaf(l.y, p.y, l.z * p.z)) >= bd - kFeatTol) { acc = acc + mk3(p.x, p.y, p.z); cnt++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(AVG_FULL, acc.x, o); acc.y += __shfl_xor_sync(AVG_FULL, acc.y, o); acc.z += __shfl_xor_sync(AVG_FULL, acc.z, o);
        cnt += __shfl_xor_sync(AVG_FULL, cnt, o);
    }
    count = cnt;
    return w.p + mmul(w.R, acc * (1.0f / (float)max(cnt, 1)));
}

struct Simplex { V3 w[4], a[4], b[4]; float lam[4]; int n; };

__device__ void closest_tri(const V3& a, const V3& b, const V3& c, float lam[3], int& mask) {
    V3 ab = b - a, ac = c - a, ap = -a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) { lam[0] = 1; lam[1] = 0; lam[2] = 0; mask = 1; return; }
    V3 bp = -b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) { lam[0] = 0; lam[1] = 1; lam[2] = 0; mask = 2; return; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { float v = d1 / (d1 - d3); lam[0] = 1 - v; lam[1] = v; lam[2] = 0; mask = 3; return; }
    V3 cp = -c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) { lam[0] = 0; lam[1] = 0; lam[2] = 1; mask = 4; return; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { float w = d2 / (d2 - d6); lam[0] = 1 - w; lam[1] = 0; lam[2] = w; mask = 5; return; }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); lam[0] = 0; lam[1] = 1 - w; lam[2] = w; mask = 6; return;
    }
    const float sum = va + vb + vc;
    if (!(sum > 0.0f)) {
        // zero-area triangle whose vertex / edge tests all fell through (collinear points in float32): closest point of
        // its longest edge, so that no 0/0 reaches the search direction
        const V3 bc = c - b;
        const float lab = dot(ab, ab), lac = dot(ac, ac), lbc = dot(bc, bc);
        const V3 p0 = (lab >= lac && lab >= lbc) ? a : (lac >= lbc ? a : b);
        const V3 e = (lab >= lac && lab >= lbc) ? ab : (lac >= lbc ? ac : bc);
        const int i0 = (lab >= lac && lab >= lbc) ? 0 : (lac >= lbc ? 0 : 1), i1 = (lab >= lac && lab >= lbc) ? 1 : 2;
        const float ee = dot(e, e);
        const float t = ee > 0.0f ? fminf(fmaxf(-dot(p0, e) / ee, 0.0f), 1.0f) : 0.0f;
        lam[0] = lam[1] = lam[2] = 0.0f; lam[i0] = 1.0f - t; lam[i1] += t;
        mask = (1 << i0) | (t > 0.0f ? (1 << i1) : 0);
        if (t >= 1.0f) mask = 1 << i1;
        return;
    }
    float den = 1.0f / sum;
    lam[1] = vb * den; lam[2] = vc * den; lam[0] = 1 - lam[1] - lam[2]; mask = 7;
}

// closest point of the simplex to the origin; returns true when the origin is enclosed
__device__ bool simplex_closest(Simplex& s, V3& v) {
    float lam[4] = {0, 0, 0, 0};
    int mask = 0;
    if (s.n == 1) { lam[0] = 1; mask = 1; }
    else if (s.n == 2) {
        V3 ab = s.w[1] - s.w[0];
        float t = -dot(s.w[0], ab), den = dot(ab, ab);
        if (t <= 0 || den <= 0) { lam[0] = 1; mask = 1; }
        else if (t >= den) { lam[1] = 1; mask = 2; }
        else { lam[1] = t / den; lam[0] = 1 - lam[1]; mask = 3; }
    } else if (s.n == 3) {
        float l3[3]; closest_tri(s.w[0], s.w[1], s.w[2], l3, mask);
        lam[0] = l3[0]; lam[1] = l3[1]; lam[2] = l3[2];
    } else {
        // Enclosed only when the origin is strictly inside all four faces of a non-degenerate tetrahedron; otherwise the
        // best closest point over ALL four faces.  Hull-minus-capsule-core differences contain parallelograms, so four
        // simplex points are routinely coplanar and culling faces by the apex side is decided by rounding noise.
        const int F[4][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 3, 1}, {1, 2, 3, 0}};
        float best = 3.0e38f; bool inside = true, degenerate = false;
#pragma unroll 1
        for (int f = 0; f < 4; ++f) {
            V3 a = s.w[F[f][0]], b = s.w[F[f][1]], c = s.w[F[f][2]], d = s.w[F[f][3]];
            V3 n = cross(b - a, c - a);
            float so = -dot(a, n), sd = dot(d - a, n);
            if (sd * sd <= 1e-10f * dot(n, n) * dot(d - a, d - a)) degenerate = true;
            if (!(so * sd > 0.0f)) inside = false;
            float l3[3]; int km;
            closest_tri(a, b, c, l3, km);
            V3 p = a * l3[0] + b * l3[1] + c * l3[2];
            float dd = dot(p, p);
            if (dd < best) {
                best = dd;
                lam[0] = lam[1] = lam[2] = lam[3] = 0;
                lam[F[f][0]] = l3[0]; lam[F[f][1]] = l3[1]; lam[F[f][2]] = l3[2];
                mask = ((km & 1) ? (1 << F[f][0]) : 0) | ((km & 2) ? (1 << F[f][1]) : 0) | ((km & 4) ? (1 << F[f][2]) : 0);
            }
        }
        if (inside && !degenerate) return true;
    }
    int n = 0; V3 p = mk3(0, 0, 0);
    for (int i = 0; i < s.n; ++i) if (mask & (1 << i)) {
        s.w[n] = s.w[i]; s.a[n] = s.a[i]; s.b[n] = s.b[i]; s.lam[n] = lam[i];
        p = p + s.w[i] * lam[i];
        n++;
    }
    s.n = n; v = p;
    return false;
}

// GJK return codes: 0 cores separated (dist, pa, pb valid); 1 cores overlap; 2 separated by more than maxdist (early
// out: every support plane gives the lower bound v.w/|v| on the distance, so well-separated candidates leave after 1-2
// iterations).  Works relative to A's position to keep float32 magnitudes small.
// Support mapping for a converged warp whose lanes hold DIFFERENT shapes (one candidate pair per thread in the
// narrowphase kernel).  Closed-form shapes are evaluated by their own lane.  Hull scans -- the only long loops of GJK --
// are served by the whole warp, one requesting lane at a time: the request (vertex array, direction) is broadcast,
// lane i looks at vertices i, i + 32, ..., the warp agrees on the largest dot product (lowest vertex index among
// equals, exactly what a serial scan returns) and the requester takes the vertex.  Reads are coalesced and a 48-vertex
// scan costs ~45 warp instructions instead of ~500 divergent ones.
__device__ __noinline__ V3 support_any(const WShape& w, V3 d, bool active, int lane) {
    const bool hull = active && w.s->type == AVG_SHAPE_HULL;
    V3 res = mk3(0, 0, 0);
    if (active && !hull) res = support(w, d);
    unsigned req = __ballot_sync(AVG_FULL, hull);
    if (req == 0) return res;
    // Serving pays when a few lanes of the warp hold hulls (mixed shape types: the others would idle through a divergent
    // scan).  When most lanes need a scan (particle-vs-hull items of Feeding / Drinking: every lane a different small hull),
    // 32 served scans cost more than one divergent scan of every lane over its own hull; the vertex picked is the same
    // (first maximum in index order).
    if (__popc(req) >= 8) { if (hull) res = support(w, d); return res; }
    const V3 l = hull ? mtmul(w.R, d) : mk3(0, 0, 0);
    const unsigned long long vp = hull ? reinterpret_cast<unsigned long long>(w.verts) : 0ull;
    const int nv = hull ? w.s->vert_cnt : 0;
    while (req) {
        const int src = __ffs(req) - 1; req &= req - 1;
        const float4* v = reinterpret_cast<const float4*>(__shfl_sync(AVG_FULL, vp, src));
        const int n = __shfl_sync(AVG_FULL, nv, src);
        const float lx = __shfl_sync(AVG_FULL, l.x, src), ly = __shfl_sync(AVG_FULL, l.y, src), lz = __shfl_sync(AVG_FULL, l.z, src);
        float bd = -3.0e38f; int bi = 0x7fffffff;                  // a lane without a vertex (or a NaN direction) keeps the sentinel
        for (int i = lane; i < n; i += 32) {
            const float4 p = __ldg(v + i);
            const float dd = fmaf(lx, p.x, fmaf(ly, p.y, lz * p.z));
            if (dd > bd) { bd = dd; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(AVG_FULL, bd, o); const int oi = __shfl_xor_sync(AVG_FULL, bi, o);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == src) { const float4 bv = __ldg(v + (bi < n ? bi : 0)); res = w.p + mmul(w.R, mk3(bv.x, bv.y, bv.z)); }   // bi >= n only for a NaN direction
    }
    return res;
}

// GJK for a converged warp, one pair per lane, iterations in lockstep (same arithmetic per lane as gjk()).  `active`
// lanes run; the return codes are those of gjk(); inactive lanes return -1.
__device__ __noinline__ int gjk_lockstep(const WShape& A, const WShape& B, bool active, int lane, float maxdist, float& dist, V3& pa, V3& pb,
                                         V3& vout, float& gap, int& iters) {
    Simplex s; s.n = 0;
    gap = 0.0f;
    bool fresh = false;
    int result = -1;
    const V3 org = A.p;
    V3 v = A.p - B.p;
    if (dot(v, v) < 1e-12f) v = mk3(1, 0, 0);
    bool run = active;
#pragma unroll 1
    for (int it = 0; it < AVG_GJK_ITERS; ++it) {
        if (!__any_sync(AVG_FULL, run)) break;
        const V3 sa = support_any(A, -v, run, lane) - org, sb = support_any(B, v, run, lane) - org;
        if (run) {
            iters = it + 1;
            const V3 w = sa - sb;
            const float vv = dot(v, v), vw = dot(v, w);
            gap = vw * rsqrtf(vv); fresh = true;
            if (vw > 0.0f && vw * vw > maxdist * maxdist * vv) { vout = v; result = 2; run = false; }
            else if (s.n > 0 && (vv - vw) <= 1e-5f * vv + 1e-10f) { result = 0; run = false; }
            else {
                bool dup = false;
                for (int i = 0; i < s.n; ++i) { V3 d = s.w[i] - w; if (dot(d, d) < 1e-14f) dup = true; }
                if (dup) { result = 0; run = false; }
                else {
                    fresh = false;
                    const bool first = s.n == 0;
                    s.w[s.n] = w; s.a[s.n] = sa; s.b[s.n] = sb; s.n++;
                    if (simplex_closest(s, v) || !(dot(v, v) >= 1e-12f)) { result = 1; run = false; }     // also catches a non-finite v
                    // float32 termination: |v| must shrink from one simplex to the next; once rounding stops it (near-touching
                    // cores, where the relative test above drowns in the noise of the support points) further
                    // iterations only cycle through the same vertices until the cap
                    else if (!first && dot(v, v) >= vv * (1.0f - AVG_GJK_SHRINK)) { result = 0; run = false; }
                }
            }
        }
    }
    if (active && result < 0) result = 0;                        // iteration cap: report the current closest points
    if (result == 0) {
        V3 a = mk3(0, 0, 0), b = mk3(0, 0, 0);
        for (int i = 0; i < s.n; ++i) { a = a + s.a[i] * s.lam[i]; b = b + s.b[i] * s.lam[i]; }
        pa = a + org; pb = b + org; dist = norm(v); vout = v;
        if (!fresh) gap = 0.0f;              // v moved after the last support evaluation: no bound for it
    }
    return result;
}

// Number of candidate axes a shape contributes to the face-normal SAT, and the k-th of them (world frame, not
// normalised), in the order of shape_axes(): boxes +-x, +-y, +-z; cylinders +-z then the radial direction towards the
// other shape; hulls their face normals.  sign = -1 when S is A, +1 when S is B.
__device__ __forceinline__ int sat_axis_count(const WShape& S) {
    const int t = S.s->type;
    return t == AVG_SHAPE_BOX ? 6 : (t == AVG_SHAPE_CYLINDER ? 3 : (t == AVG_SHAPE_HULL ? S.s->plane_cnt : 0));
}
__device__ __forceinline__ V3 sat_axis_dir(const WShape& S, const WShape& O, float sign, int k) {
    const int t = S.s->type;
    if (t == AVG_SHAPE_HULL) return mmul(S.R, mk3(S.planes[4 * k], S.planes[4 * k + 1], S.planes[4 * k + 2])) * sign;
    if (t == AVG_SHAPE_BOX) { const int ax = k >> 1; return mk3(S.R[ax], S.R[3 + ax], S.R[6 + ax]) * ((k & 1) ? -sign : sign); }
    const V3 az = mk3(S.R[2], S.R[5], S.R[8]);                   // cylinder
    if (k < 2) return az * (k ? -sign : sign);
    const V3 d = O.p - S.p;
    return (d - az * dot(d, az)) * sign;
}

// Face-normal SAT (cores overlap; DESIGN.md "deep penetration") for a converged warp with one pair per lane: the lanes
// that need it are served one at a time, the requester's two shapes are broadcast and the candidate axes are spread
// over the lanes (every lane scans the same vertices at the same time: broadcast loads).  The winner is the smallest
// depth, lowest axis index among equals -- the result of the serial scan in shape_axes()/sat_axis().
__device__ __noinline__ void sat_served(const WShape& A, const WShape& B, bool need, int lane, float& best_out, V3& bn_out, V3& bpa_out) {
    unsigned req = __ballot_sync(AVG_FULL, need);
    while (req) {
        const int src = __ffs(req) - 1; req &= req - 1;
        WShape a, b;
        a.s = reinterpret_cast<const AvgShape*>(__shfl_sync(AVG_FULL, reinterpret_cast<unsigned long long>(A.s), src));
        b.s = reinterpret_cast<const AvgShape*>(__shfl_sync(AVG_FULL, reinterpret_cast<unsigned long long>(B.s), src));
        a.verts = reinterpret_cast<const float*>(__shfl_sync(AVG_FULL, reinterpret_cast<unsigned long long>(A.verts), src));
        b.verts = reinterpret_cast<const float*>(__shfl_sync(AVG_FULL, reinterpret_cast<unsigned long long>(B.verts), src));
        a.planes = reinterpret_cast<const float*>(__shfl_sync(AVG_FULL, reinterpret_cast<unsigned long long>(A.planes), src));
        b.planes = reinterpret_cast<const float*>(__shfl_sync(AVG_FULL, reinterpret_cast<unsigned long long>(B.planes), src));
        a.p = shfl3(A.p, src); b.p = shfl3(B.p, src);
#pragma unroll
        for (int i = 0; i < 9; ++i) { a.R[i] = __shfl_sync(AVG_FULL, A.R[i], src); b.R[i] = __shfl_sync(AVG_FULL, B.R[i], src); }
        const int ka = sat_axis_count(a), kb = sat_axis_count(b), K = ka + kb + 1;
        float best = 3.0e38f; int bk = 0x7fffffff; V3 bn = mk3(0, 0, 1);
        for (int k = lane; k < K; k += 32) {
            V3 n = k < ka ? sat_axis_dir(a, b, -1.0f, k) : (k < ka + kb ? sat_axis_dir(b, a, 1.0f, k - ka) : a.p - b.p);
            const float ln = norm(n);
            if (ln < 1e-9f) continue;
            n = n * (1.0f / ln);
            const V3 sa = support(a, -n), sb = support(b, n);
            const float depth = dot(sb - sa, n);
            if (depth < best) { best = depth; bk = k; bn = n; }
        }
        float wbest = best; int wk = bk;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(AVG_FULL, wbest, o); const int ok = __shfl_xor_sync(AVG_FULL, wk, o);
            if (od < wbest || (od == wbest && ok < wk)) { wbest = od; wk = ok; }
        }
        const int win = __ffs(__ballot_sync(AVG_FULL, bk == wk && wk != 0x7fffffff)) - 1;
        const int w = win < 0 ? 0 : win;
        const V3 n = shfl3(bn, w);
        int ca = 1, cb = 1;
        V3 fa = mk3(0, 0, 0), fb = fa;
        if (win >= 0) { fa = support_feature_warp(a, -n, ca, lane); fb = support_feature_warp(b, n, cb, lane); }   // a, b: the requester's shapes in every lane
        if (lane == src) {
            // witness on A's core: from the round shape when there is one, otherwise from the smaller supporting feature
            // (sat_axis() of the oracle)
            const bool ra = A.s->type == AVG_SHAPE_SPHERE || A.s->type == AVG_SHAPE_CAPSULE;
            const bool rb = B.s->type == AVG_SHAPE_SPHERE || B.s->type == AVG_SHAPE_CAPSULE;
            if (win >= 0) {
                best_out = wbest; bn_out = n;
                bpa_out = (rb && !ra) ? fb - n * wbest : (ra ? fa : (ca <= cb ? fa : fb - n * wbest));
            }
            else { best_out = 3.0e38f; bn_out = mk3(0, 0, 1); bpa_out = A.p; }
        }
    }
}

// World pose of any shape of the table from the body poses in shared memory: top-level moving shapes are cached (s.sp / s.sq),
// children of compound shapes hang on their body (dynamic or env-static), everything else is static.
template <class SM>
__device__ __forceinline__ void any_shape_pose(const KM& m, const SM& s, int si, V3& p, Q4& q) {
    const AvgShape* S = &m.shape[si];
    if (si < m.h->n_mshape) { p = ld3(s.sp[si]); q = ldq(s.sq[si]); }
    else if (S->body >= 0) { const Q4 bq = ldq(s.bq[S->body]); p = ld3(s.bp[S->body]) + qrot(bq, ld3(S->pos)); q = qnormalize(qmul(bq, ldq(S->quat))); }
    else { p = ld3(S->pos); q = ldq(S->quat); }
}
// World AABB (centre, half extents) of shape si for the expansion of compound candidates: children from their body-frame box
// (m.caabb), top-level moving shapes from the cache, static shapes from their broadphase record.
template <class SM>
__device__ __forceinline__ void any_shape_aabb(const KM& m, const SM& s, int si, V3& c, V3& hh) {
    const int nms = m.h->n_mshape, ns = m.h->n_shape;
    if (si >= ns) {
        const float4 c4 = __ldg(&m.caabb[2 * (si - ns)]), h4 = __ldg(&m.caabb[2 * (si - ns) + 1]);
        const int b = m.shape[si].body;
        const M3 R = qmat(ldq(s.bq[b]));
        c = ld3(s.bp[b]) + mmul(R.m, mk3(c4.x, c4.y, c4.z));
        hh = mk3(fabsf(R.m[0]) * h4.x + fabsf(R.m[1]) * h4.y + fabsf(R.m[2]) * h4.z, fabsf(R.m[3]) * h4.x + fabsf(R.m[4]) * h4.y + fabsf(R.m[5]) * h4.z,
                 fabsf(R.m[6]) * h4.x + fabsf(R.m[7]) * h4.y + fabsf(R.m[8]) * h4.z);
    } else if (si < nms) {
        const float4 a0 = s.saabb[si][0], a1 = s.saabb[si][1];
        c = mk3(a0.x, a0.y, a0.z); hh = mk3(a0.w, a1.x, a1.y);
    } else {
        const float4* rp = reinterpret_cast<const float4*>(&m.bps[si - nms]);
        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
        c = mk3(r0.x, r0.y, r0.z); hh = mk3(r0.w, r1.x, r1.y);
    }
}

template <class SM>
__device__ void collide_warp(const KM& m, SM& s, int lane, int env_index, int& ncontact, int& overflow, int& ncand_out, int nsep, float* gsep_f, int& nsep_out,
                             AvgNpItem* np_queue, int* np_count, int np_capacity, int dbg) {
    const AvgModelHeader* h = m.h;
    const int nms = h->n_mshape;
    // world pose + AABB of the moving shapes
    if (lane < nms) {
        const AvgShape* S = &m.shape[lane];
        V3 bp; Q4 bq;
        body_pose(s, S->body, bp, bq);
        V3 p = bp + qrot(bq, ld3(S->pos));
        const Q4 sq = qnormalize(qmul(bq, ldq(S->quat)));
        M3 R = qmat(sq);
        st3(s.sp[lane], p);
        s.sq[lane][0] = sq.x; s.sq[lane][1] = sq.y; s.sq[lane][2] = sq.z; s.sq[lane][3] = sq.w;
#pragma unroll
        for (int i = 0; i < 9; ++i) s.sR[lane][i] = R.m[i];
        V3 lc = ld3(S->aabb_c), lh = ld3(S->aabb_h);
        V3 c = p + mmul(R.m, lc);
        s.saabb[lane][0] = make_float4(c.x, c.y, c.z, fabsf(R.m[0]) * lh.x + fabsf(R.m[1]) * lh.y + fabsf(R.m[2]) * lh.z);
        s.saabb[lane][1] = make_float4(fabsf(R.m[3]) * lh.x + fabsf(R.m[4]) * lh.y + fabsf(R.m[5]) * lh.z,
                                       fabsf(R.m[6]) * lh.x + fabsf(R.m[7]) * lh.y + fabsf(R.m[8]) * lh.z, S->thr, 0.0f);
        const float4 c0 = __ldg(&m.bcap[2 * lane]), c1 = __ldg(&m.bcap[2 * lane + 1]);
        const V3 w0 = p + mmul(R.m, mk3(c0.x, c0.y, c0.z)), w1 = p + mmul(R.m, mk3(c1.x, c1.y, c1.z));
        s.scap[lane][0] = make_float4(w0.x, w0.y, w0.z, c0.w); s.scap[lane][1] = make_float4(w1.x, w1.y, w1.z, 0.0f);
    }
    __syncwarp();
    // broadphase.  Static shapes: one lane per static shape (its AABB, threshold and the bit mask of moving shapes it
    // may touch arrive in two coalesced 16-byte loads), moving-shape AABBs are broadcast from shared memory.
    int ncand = 0;
    const int nstat = h->n_shape - nms;
    // (a) union box of the moving shapes (their thresholds included); static shapes outside it are dropped at once and
    //     the survivors are compacted, so the pair loop below runs over ~1/3 of the static set
    float ulo[3], uhi[3];
    {
        const bool v = lane < nms;
        const float4 b0 = v ? s.saabb[lane][0] : make_float4(0, 0, 0, 0), b1 = v ? s.saabb[lane][1] : make_float4(0, 0, 0, 0);
        const float c[3] = {b0.x, b0.y, b0.z}, hh[3] = {b0.w + b1.z, b1.x + b1.z, b1.y + b1.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float lo = v ? c[k] - hh[k] : 3.0e38f, hi = v ? c[k] + hh[k] : -3.0e38f;
            // min / max over the warp with one redux.sync each, on the order-preserving integer image of the floats
            auto key = [](float f) { const unsigned b = __float_as_uint(f); return b ^ ((unsigned)((int)b >> 31) | 0x80000000u); };
            auto unkey = [](unsigned q) { return __uint_as_float(q ^ ((q >> 31) ? 0x80000000u : 0xffffffffu)); };
            ulo[k] = unkey(__reduce_min_sync(AVG_FULL, key(lo))); uhi[k] = unkey(__reduce_max_sync(AVG_FULL, key(hi)));
        }
    }
    int nnear = 0;
    for (int base = 0; base < nstat; base += 32) {
        const int si = base + lane;
        bool near = false;
        if (si < nstat) {
            const float4* rp = reinterpret_cast<const float4*>(&m.bps[si]);
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            near = r0.x + r0.w >= ulo[0] && r0.x - r0.w <= uhi[0] && r0.y + r1.x >= ulo[1] && r0.y - r1.x <= uhi[1] &&
                   r0.z + r1.y >= ulo[2] && r0.z - r1.y <= uhi[2];
        }
        const unsigned bal = __ballot_sync(AVG_FULL, near);
        if (near) s.near_idx[nnear + __popc(bal & ((1u << lane) - 1))] = (uint8_t)si;
        nnear += __popc(bal);
    }
    __syncwarp();
    // (b) one lane per surviving static shape (AABB, threshold and the bit mask of moving shapes it may touch arrive in
    //     two 16-byte loads); moving-shape AABBs are broadcast from shared memory.  Each lane first collects its hits
    //     in a bit mask (independent iterations, no votes inside the loop), then the warp compacts them with one
    //     prefix sum over the per-lane counts.  Order does not matter here: candidates are ranked below.
    auto emit = [&](uint32_t hm, uint32_t other) {
        const int cnt = __popc(hm);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(AVG_FULL, inc, o); if (lane >= o) inc += v; }
        int slot = ncand + inc - cnt;
        while (hm) {
            const int a = __ffs(hm) - 1; hm &= hm - 1;
            if (slot < kMaxCand) s.cand[slot] = (uint32_t)a | (other << 16);
            ++slot;
        }
        ncand += __shfl_sync(AVG_FULL, inc, 31);
    };
    if (dbg & 2) nnear = 0;
    for (int base = 0; base < nnear; base += 32) {
        const int k = base + lane;
        int si = 0;
        float4 r0 = make_float4(0, 0, 0, 0), r1 = make_float4(0, 0, 0, 0);
        uint32_t mask = 0;
        if (k < nnear) {
            si = s.near_idx[k];
            const float4* rp = reinterpret_cast<const float4*>(&m.bps[si]);
            r0 = __ldg(rp); r1 = __ldg(rp + 1);
            mask = __float_as_uint(r1.w);
        }
        uint32_t hm = 0;
#pragma unroll 4
        for (int a = 0; a < nms; ++a) {
            const float4 a0 = s.saabb[a][0], a1 = s.saabb[a][1];
            const float thr = fminf(a1.z, r1.z);
            const bool hit = fabsf(a0.x - r0.x) <= a0.w + r0.w + thr && fabsf(a0.y - r0.y) <= a1.x + r1.x + thr &&
                             fabsf(a0.z - r0.z) <= a1.y + r1.y + thr;
            hm |= (hit ? 1u : 0u) << a;
        }
        emit(hm & mask, (uint32_t)(nms + si));
    }
    // moving-moving pairs: lane b against every a < b allowed by the filter masks (bit b of bpm[a])
    if (!(dbg & 8)) {
        float4 b0 = make_float4(0, 0, 0, 0), b1 = make_float4(0, 0, 0, 0);
        if (lane < nms) { b0 = s.saabb[lane][0]; b1 = s.saabb[lane][1]; }
        uint32_t hm = 0, allowed = 0;
#pragma unroll 4
        for (int a = 0; a < nms; ++a) {
            const uint32_t mask = __ldg(&m.bpm[a]);
            const float4 a0 = s.saabb[a][0], a1 = s.saabb[a][1];
            const float thr = fminf(a1.z, b1.z);
            const bool hit = fabsf(a0.x - b0.x) <= a0.w + b0.w + thr && fabsf(a0.y - b0.y) <= a1.x + b1.x + thr &&
                             fabsf(a0.z - b0.z) <= a1.y + b1.y + thr;
            hm |= (hit ? 1u : 0u) << a;
            allowed |= ((mask >> lane) & 1u) << a;
        }
        emit(lane < nms ? (hm & allowed) : 0u, (uint32_t)lane);
    }
    if (ncand > kMaxCand) { overflow |= 4; ncand = kMaxCand; }
    ncand_out = ncand;
    __syncwarp();
    // bounding-capsule cull (conservative): segment-segment distance minus radii against the pair threshold.  Elongated
    // links have fat AABBs; this removes most candidates before the much more expensive GJK.  One slab of 32
    // candidates at a time, survivors compacted in place.
    {
        int nk = 0;
#pragma unroll 1
        for (int base = 0; base < ncand; base += 32) {
            const int ci = base + lane;
            bool ok = false; uint32_t pr = 0;
            if (ci < ncand) {
                pr = s.cand[ci];
                const int a = pr & 0xffff, b = pr >> 16;
                const float4 a0 = s.scap[a][0], a1 = s.scap[a][1];
                float4 b0, b1;
                if (b < nms) { b0 = s.scap[b][0]; b1 = s.scap[b][1]; }
                else { b0 = __ldg(&m.bcap[2 * b]); b1 = __ldg(&m.bcap[2 * b + 1]); }
                const float thr = fminf(__ldg(&m.shape[a].thr), __ldg(&m.shape[b].thr));
                // closest distance between segments [a0,a1] and [b0,b1] (Ericson 5.1.9)
                const V3 p1 = mk3(a0.x, a0.y, a0.z), p2 = mk3(b0.x, b0.y, b0.z);
                const V3 d1 = mk3(a1.x, a1.y, a1.z) - p1, d2 = mk3(b1.x, b1.y, b1.z) - p2, r = p1 - p2;
                const float aa = dot(d1, d1), ee = dot(d2, d2), ff = dot(d2, r);
                float sc = 0.0f, tc = 0.0f;
                if (aa <= 1e-12f && ee <= 1e-12f) { }
                else if (aa <= 1e-12f) tc = fminf(fmaxf(ff / ee, 0.0f), 1.0f);
                else {
                    const float cc = dot(d1, r);
                    if (ee <= 1e-12f) sc = fminf(fmaxf(-cc / aa, 0.0f), 1.0f);
                    else {
                        const float bb = dot(d1, d2), den = aa * ee - bb * bb;
                        sc = den > 1e-12f ? fminf(fmaxf((bb * ff - cc * ee) / den, 0.0f), 1.0f) : 0.0f;
                        tc = (bb * sc + ff) / ee;
                        if (tc < 0.0f) { tc = 0.0f; sc = fminf(fmaxf(-cc / aa, 0.0f), 1.0f); }
                        else if (tc > 1.0f) { tc = 1.0f; sc = fminf(fmaxf((bb - cc) / aa, 0.0f), 1.0f); }
                    }
                }
                const V3 dd = (p1 + d1 * sc) - (p2 + d2 * tc);
                const float lim = a0.w + b0.w + thr + 1e-5f;
                ok = dot(dd, dd) <= lim * lim;
            }
            __syncwarp();
            const unsigned bal = __ballot_sync(AVG_FULL, ok);
            if (ok) s.cand[nk + __popc(bal & ((1u << lane) - 1))] = pr;      // nk <= base: never ahead of the slab being read
            nk += __popc(bal);
            __syncwarp();
        }
        ncand = nk;
    }
    // Compound shapes (VHACD tool / bowl / head): a surviving candidate that involves one is replaced by the pairs of its
    // convex children whose boxes overlap (children x other shape, or children x children), the lanes sharing the products.
    if (h->n_cshape > 0) {
        __syncwarp();
        int nx = 0;
#pragma unroll 1
        for (int base = 0; base < ncand; base += 32) {
            const int ci = base + lane;
            const uint32_t pr = ci < ncand ? s.cand[ci] : 0u;
            const bool comp = ci < ncand && (m.shape[pr & 0xffff].type == AVG_SHAPE_COMPOUND || m.shape[pr >> 16].type == AVG_SHAPE_COMPOUND);
            const unsigned bplain = __ballot_sync(AVG_FULL, ci < ncand && !comp);
            unsigned bcomp = __ballot_sync(AVG_FULL, comp);
            if (ci < ncand && !comp) { const int k = nx + __popc(bplain & ((1u << lane) - 1)); if (k < kMaxCand) s.cand2[k] = pr; }
            nx += __popc(bplain);
            while (bcomp) {
                const int src = __ffs(bcomp) - 1; bcomp &= bcomp - 1;
                const uint32_t q = __shfl_sync(AVG_FULL, pr, src);
                const int a = q & 0xffff, b = q >> 16;
                const AvgShape* SA = &m.shape[a]; const AvgShape* SB = &m.shape[b];
                const int fa = SA->type == AVG_SHAPE_COMPOUND ? SA->vert_off : a, na = SA->type == AVG_SHAPE_COMPOUND ? SA->vert_cnt : 1;
                const int fb = SB->type == AVG_SHAPE_COMPOUND ? SB->vert_off : b, nbb = SB->type == AVG_SHAPE_COMPOUND ? SB->vert_cnt : 1;
                const float thr = fminf(SA->thr, SB->thr);
                const int total = na * nbb;
#pragma unroll 1
                for (int t0 = 0; t0 < total; t0 += 32) {
                    const int t = t0 + lane;
                    bool hit = false; int ca = 0, cb = 0;
                    if (t < total) {
                        ca = fa + t / nbb; cb = fb + t % nbb;
                        V3 c1, h1, c2, h2;
                        any_shape_aabb(m, s, ca, c1, h1); any_shape_aabb(m, s, cb, c2, h2);
                        hit = fabsf(c1.x - c2.x) <= h1.x + h2.x + thr && fabsf(c1.y - c2.y) <= h1.y + h2.y + thr && fabsf(c1.z - c2.z) <= h1.z + h2.z + thr;
                    }
                    const unsigned bh = __ballot_sync(AVG_FULL, hit);
                    if (hit) { const int k = nx + __popc(bh & ((1u << lane) - 1)); if (k < kMaxCand) s.cand2[k] = (uint32_t)ca | ((uint32_t)cb << 16); }
                    nx += __popc(bh);
                }
            }
        }
        if (nx > kMaxCand) { overflow |= 4; nx = kMaxCand; }
        __syncwarp();
        for (int i = lane; i < nx; i += 32) s.cand[i] = s.cand2[i];
        ncand = nx;
        __syncwarp();
    }
    // canonical order = pair-table order: ascending (moving shape a, other shape b); rank by counting (the list is short)
#pragma unroll 1
    for (int base = 0; base < ncand; base += 32) {
        const int ci = base + lane;
        const uint32_t mine = ci < ncand ? s.cand[ci] : 0u;
        const uint32_t km = ((mine & 0xffffu) << 16) | (mine >> 16);
        int rank = 0;
        for (int j = 0; j < ncand; ++j) {
            const uint32_t o = s.cand[j];
            rank += (((o & 0xffffu) << 16) | (o >> 16)) < km;
        }
        if (ci < ncand) s.cand2[rank] = mine;
    }
    __syncwarp();
    // Hand-off to the narrowphase kernel, with separation certificates (temporal coherence).
    // Most candidates that survive the culls are close but not touching, sub-step after sub-step (hand against its own
    // finger tips, arm over the armrest).  When GJK proves a pair separated along a direction v by at least g
    // (support-plane bound, the predicate of its own early-out), the arena remembers, in A's frame, v, g and the pose of
    // B relative to A.  Next sub-step the pair is rejected
    //   (1) here, without touching the shapes, when g minus a bound on how far B's points can have moved in A's frame
    //       (translation + chord of the rotation x bounding radius) still exceeds the contact distance, else
    //   (2) in the narrowphase kernel by one support-plane test along the remembered direction (renews the certificate).
    // Both are sufficient conditions for "no contact"; a pair that fails them runs the full GJK, so certificates can
    // only skip work, never change a contact.  Everything not rejected by (1) is queued: one work item per pair for the
    // narrowphase kernel (one THREAD per item there, instead of one lane of an otherwise idle warp here), with a result
    // slot in pair order and a slot of the new certificate list to fill in.
    int ncarry = 0, nq = 0;
    float4* gsep = reinterpret_cast<float4*>(gsep_f);
    if (dbg & 1) ncand = 0;
    // pass A: decide per candidate (carried by its certificate / queued), carried certificates move to the front of the
    // new list; the decision and the cache entry that holds a direction hint are kept in shared memory for pass B
#pragma unroll 1
    for (int base = 0; base < ncand; base += 32) {
        const int ci = base + lane;
        bool carried = false, queued = false;
        float4 k0 = make_float4(0, 0, 0, 0), k1 = k0, k2 = k0;
        int f = -1;
        if (ci < ncand) {
            const uint32_t pr = s.cand2[ci];
            const int a = pr & 0xffff, b = pr >> 16;
            const AvgShape* SA = &m.shape[a]; const AvgShape* SB = &m.shape[b];
            queued = true;
            if (SB->type != AVG_SHAPE_PLANE) {
                for (int e = 0; e < nsep; ++e) if (__float_as_uint(s.sep[0][e].w) == pr) f = e;
                if (f >= 0) {
                    const float4 e0 = s.sep[0][f], e1 = s.sep[1][f], e2 = s.sep[2][f];
                    if (e1.w > 0.0f) {
                        V3 paw, pbw; Q4 qa, qb;
                        any_shape_pose(m, s, a, paw, qa); any_shape_pose(m, s, b, pbw, qb);
                        const V3 prel = qrot_inv(qa, pbw - paw);
                        const Q4 qrel = qmul(qconj(qa), qb);
                        const float md = fminf(SA->thr, SB->thr) + SA->margin + SB->margin;
                        const V3 dp = prel - mk3(e1.x, e1.y, e1.z);
                        const float sg = (qrel.x * e2.x + qrel.y * e2.y + qrel.z * e2.z + qrel.w * e2.w) < 0.0f ? -1.0f : 1.0f;
                        const float dx = qrel.x - sg * e2.x, dy = qrel.y - sg * e2.y, dz = qrel.z - sg * e2.z, dw = qrel.w - sg * e2.w;
                        const float chord = 2.0f * sqrtf(dx * dx + dy * dy + dz * dz + dw * dw);     // >= 2 sin(angle / 2)
                        const float4 c0 = __ldg(&m.bcap[2 * b]), c1 = __ldg(&m.bcap[2 * b + 1]);
                        const V3 o = SB->body >= 0 ? mk3(0, 0, 0) : ld3(SB->pos);
                        const float rb = fmaxf(norm(mk3(c0.x, c0.y, c0.z) - o), norm(mk3(c1.x, c1.y, c1.z) - o)) + c0.w;
                        if (e1.w - (norm(dp) + chord * rb) > md + 1e-5f) { carried = true; queued = false; k0 = e0; k1 = e1; k2 = e2; }
                    }
                }
            }
            s.candf[ci] = (uint8_t)(queued ? (f >= 0 ? f : 254) : 255);           // 255: carried, 254: queued without a hint
        }
        const unsigned bc = __ballot_sync(AVG_FULL, carried), bq = __ballot_sync(AVG_FULL, queued);
        if (carried) {
            const int slot = ncarry + __popc(bc & ((1u << lane) - 1));
            if (slot < AVG_S_NSEPMAX) { gsep[slot] = k0; gsep[AVG_S_NSEPMAX + slot] = k1; gsep[2 * AVG_S_NSEPMAX + slot] = k2; }
        }
        ncarry += __popc(bc); nq += __popc(bq);
    }
    ncarry = min(ncarry, AVG_S_NSEPMAX);
    if (nq > AVG_S_NQMAX) { overflow |= 1; nq = AVG_S_NQMAX; }
    // pass B: reserve queue space for the warp's items, write them, and seed the certificate slots they will fill in
    int qbase = 0;
    if (lane == 0 && nq > 0) qbase = atomicAdd(np_count, nq);
    qbase = __shfl_sync(AVG_FULL, qbase, 0);
    if (qbase + nq > np_capacity) { overflow |= 1; nq = max(0, min(nq, np_capacity - qbase)); }
    __syncwarp();
    {
        int qn = 0;
#pragma unroll 1
        for (int base = 0; base < ncand; base += 32) {
            const int ci = base + lane;
            const int fcode = ci < ncand ? s.candf[ci] : 255;
            const bool queued = fcode != 255;
            const unsigned bq = __ballot_sync(AVG_FULL, queued);
            const int qslot = qn + __popc(bq & ((1u << lane) - 1));
            if (queued && qslot < nq) {
                const int cs = ncarry + qslot < AVG_S_NSEPMAX ? ncarry + qslot : -1;
                if (cs >= 0) {                    // direction hint (or none), gap <= 0: not a certificate until the narrowphase says so
                    const float4 hint = fcode < 254 ? s.sep[0][fcode] : make_float4(0, 0, 0, 0);
                    gsep[cs] = make_float4(hint.x, hint.y, hint.z, __uint_as_float(0xffffffffu));
                    gsep[AVG_S_NSEPMAX + cs] = make_float4(0, 0, 0, 0);
                }
                AvgNpItem it; it.env = env_index; it.pair = s.cand2[ci]; it.slot = qslot; it.cert = cs;
                np_queue[qbase + qslot] = it;
            }
            qn += __popc(bq);
        }
    }
    nsep_out = min(ncarry + nq, AVG_S_NSEPMAX);
    ncontact = nq;                                // number of queued candidates (results arrive from the narrowphase kernel)
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Dynamics helpers
// ---------------------------------------------------------------------------------------------------------------
struct LaneDyn {
    Sv S;             // motion subspace of this lane's joint (zero for free bodies / idle lanes)
    uint32_t anc;     // ancestor-or-self mask
};

}  // namespace

#define AVG_KERNEL_PREAMBLE(SMTYPE)                                                             \
    extern __shared__ __align__(16) unsigned char smem_raw[];                                    \
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;                                  \
    const int e = a.env_begin + blockIdx.x * (blockDim.x >> 5) + warp;                           \
    if (e >= a.env_end) return;                                                                  \
    SMTYPE& s = reinterpret_cast<SMTYPE*>(smem_raw)[warp];                                       \
    const int variant = a.variant ? a.variant[e] : 0;                                            \
    const KM m = c_models[a.slot][variant];                                                      \
    const AvgModelHeader* h = m.h;                                                               \
    float* grec = a.env + (size_t)e * AVG_ENV_STRIDE;                                            \
    float* scr = a.scratch + (size_t)e * AVG_S_STRIDE;                                           \
    (void)lane; (void)s; (void)h; (void)grec; (void)scr;
/* settle / reset paths step only the environments whose mask byte is set */
#define AVG_MASK_CHECK if (a.mask && !a.mask[e]) return;

// L2 prefetch of what the warp that will run `ahead` environments later is going to read first (its blocks are dispatched
// as the resident ones retire, so by then the lines sit in L2 instead of HBM): lane l touches 128-byte line l of the record
// / of the given scratch sections.  A hint only: no register, no scoreboard, dropped if the address is out of range.
#ifndef AVG_PREFETCH
#define AVG_PREFETCH 1
#endif
#ifndef AVG_PF_PCT
#define AVG_PF_PCT 50             /* prefetch distance in percent of the resident environments of the kernel */
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) {
#if AVG_PREFETCH
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// =================================================================================================================
// action -> motor targets, env.py:274-337 (one lane per dof; no shared memory)
// =================================================================================================================
__device__ __forceinline__ void prologue_body(const AvgStepArgs& a, const int e, const int lane, const KM& m, float* grec, float* scr) {
    const AvgModelHeader* h = m.h;
    int* scr_i = reinterpret_cast<int*>(scr);
    const int nj = h->n_jdof;
    const int na = h->n_action_robot + h->n_action_human;
    const float* act = a.actions + (size_t)e * na;
    const bool tremor = grec[AVG_E_TREMOR_ON] != 0.0f;
    const bool human_active = h->human_control || tremor;
    const int iteration = reinterpret_cast<const int*>(grec)[AVG_E_ITERATION];
    if (lane < nj) {
        const AvgDof* D = &m.dof[lane];
        const int ai = D->action, slot = D->human_slot;
        if (ai >= 0 && ai < h->n_action_robot) {
            float av = fminf(fmaxf(act[ai], -1.0f), 1.0f) * h->action_scale;
            float pos = grec[AVG_E_Q + m.body[D->body].qidx];
            for (int f = 0; f < h->substeps; ++f) {
                if (pos + av < D->rep_lower) av = 0.0f;
                if (pos + av > D->rep_upper) av = 0.0f;
                pos += av;
            }
            grec[AVG_E_MTARGET + lane] = pos;
        } else if (slot >= 0 && human_active) {
            float av = 0.0f;
            if (h->human_control) av = fminf(fmaxf(act[h->n_action_robot + slot], -1.0f), 1.0f) * h->action_scale;
            const float sc = grec[AVG_E_LIMIT_SCALE];
            const float lo = D->lower * sc, hi = D->upper * sc;
            float pos = grec[AVG_E_Q + m.body[D->body].qidx];
            float tgt = grec[AVG_E_TARGET_H + slot];
            const float sgn = (iteration % 2 == 0) ? 1.0f : -1.0f;
            for (int f = 0; f < h->substeps; ++f) {
                if (pos + av < lo) av = 0.0f;
                if (pos + av > hi) av = 0.0f;
                if (tremor) { pos = tgt + grec[AVG_E_TREMOR + slot] * sgn; tgt += av; }
                pos += av;
            }
            grec[AVG_E_TARGET_H + slot] = tgt;
            grec[AVG_E_MTARGET + lane] = pos;
        }
    }
    if (lane == 0) {
        if (human_active) grec[AVG_E_HUMAN_KP] = h->task_f[AVG_TF_HUMAN_KP_ACTIVE];
        scr_i[AVG_S_ITERS] = 0; scr_i[AVG_S_OVERFLOW] = 0; scr_i[AVG_S_NCAND] = 0;
    }
}
__global__ void __launch_bounds__(32 * AVG_K_WARPS_PER_BLOCK)
avg_prologue_kernel(AvgStepArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = a.env_begin + blockIdx.x * (blockDim.x >> 5) + warp;
    if (e >= a.env_end) return;
    const int variant = a.variant ? a.variant[e] : 0;
    const KM m = c_models[a.slot][variant];
    prologue_body(a, e, lane, m, a.env + (size_t)e * AVG_ENV_STRIDE, a.scratch + (size_t)e * AVG_S_STRIDE);
}

// =================================================================================================================
// forward kinematics + collision -> contact list in the scratch arena
// =================================================================================================================
// Body of the collide kernel for ONE environment.  The work items go to `np_queue` / `np_count` (the global narrowphase queue).
template <int PF_AHEAD>
__device__ __forceinline__ void collide_body(const AvgStepArgs& a, SmCollide& s, const int e, const int lane, const KM& m, float* grec, float* scr,
                                             AvgNpItem* np_queue, int* np_count, int np_capacity) {
    const AvgModelHeader* h = m.h;
    AVG_MASK_CHECK
    s.q[lane] = grec[AVG_E_Q + lane];
    {   // L2 prefetch for the environment whose warp takes this slot next: positions (1 line), counters (1), certificate cache (3)
        const int ea = e + PF_AHEAD;
        if (PF_AHEAD > 0 && ea < a.env_end) {
            const float* r2 = a.env + (size_t)ea * AVG_ENV_STRIDE; const float* s2 = a.scratch + (size_t)ea * AVG_S_STRIDE;
            const float* p = lane < 1 ? r2 : (lane < 2 ? s2 : (lane < 5 ? s2 + AVG_S_SEP + 4 * AVG_S_NSEPMAX * (lane - 2) : nullptr));   // first 8 certificates of each of the 3 planes
            if (p) prefetch_l2(p);
        }
    }
    __syncwarp();
    fk_warp(m, s, s.q, lane, h->n_body, grec + AVG_E_EBODY);
    if (lane < h->n_body + h->n_ebody) {     // the dynamics / narrowphase / particle kernels of this sub-step reuse the poses
        float4* gp = reinterpret_cast<float4*>(scr + AVG_S_POSE) + 2 * lane;
        gp[0] = make_float4(s.bp[lane][0], s.bp[lane][1], s.bp[lane][2], 0.0f);
        gp[1] = make_float4(s.bq[lane][0], s.bq[lane][1], s.bq[lane][2], s.bq[lane][3]);
    }
    int nc = 0, overflow = 0, ncand = 0, nsep_out = 0;
    int* scr_i = reinterpret_cast<int*>(scr);
    float4* gsep = reinterpret_cast<float4*>(scr + AVG_S_SEP);
    int nsep = (a.dbg & 16) ? 0 : min(max(scr_i[AVG_S_NSEP], 0), AVG_S_NSEPMAX);
    if (lane < nsep) { s.sep[0][lane] = gsep[lane]; s.sep[1][lane] = gsep[AVG_S_NSEPMAX + lane]; s.sep[2][lane] = gsep[2 * AVG_S_NSEPMAX + lane]; }
    if (!(a.dbg & 4)) collide_warp(m, s, lane, e, nc, overflow, ncand, nsep, scr + AVG_S_SEP, nsep_out, np_queue, np_count,
                                   np_capacity, a.dbg);
    if (lane == 0) { scr_i[AVG_S_NQ] = nc; scr_i[AVG_S_NSEP] = nsep_out; scr_i[AVG_S_NCAND] += ncand; if (overflow) scr_i[AVG_S_OVERFLOW] |= overflow; }
}
// OCC = blocks per SM the instance is compiled for: AVG_OCC_COLLIDE (5 blocks of 4 warps, 90 registers, nothing spilled) for handles
// below AVG_COLLIDE_LARGE environments, where the step waits for this kernel's slowest warps; AVG_OCC_COLLIDE_LARGE (7 blocks, 72
// registers, a few spills) above, where 28 instead of 20 resident warps per SM hide more of its load latency.  Measured on a B200,
// 7 against 5 blocks: 393216 envs (staggered episodes) 43.35 -> 42.51 ms per step, policy rollout at 196608 envs +1.3 %,
// ScratchItchPR2 at 131072 envs +1.6 %; 32768 envs -1.4 %, 4096 envs -3.7 % (6 blocks: +0.5 % / 8 blocks: +1.3 % at 393216).
#ifndef AVG_OCC_COLLIDE_LARGE
#define AVG_OCC_COLLIDE_LARGE (28 / AVG_WPB_COLLIDE)
#endif
#ifndef AVG_COLLIDE_LARGE
#define AVG_COLLIDE_LARGE 131072     /* environments of the HANDLE (not of the launch): the batches avg_step runs as one sequence; choosing by the
                                        handle keeps every partition of a batch (stream halves, host chunks, masked resets) on the same instance --
                                        ptxas contracts mul + add differently at the two register budgets, so the instances differ by ulps */
#endif
__global__ void __launch_bounds__(32 * AVG_WPB_COLLIDE, AVG_OCC_COLLIDE)
avg_collide_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmCollide)
    collide_body<148 * AVG_OCC_COLLIDE * AVG_WPB_COLLIDE * AVG_PF_PCT / 100>(a, s, e, lane, m, grec, scr, a.np_queue, a.np_count, a.np_capacity);
}
__global__ void __launch_bounds__(32 * AVG_WPB_COLLIDE, AVG_OCC_COLLIDE_LARGE)
avg_collide_large_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmCollide)
    collide_body<148 * AVG_OCC_COLLIDE_LARGE * AVG_WPB_COLLIDE * AVG_PF_PCT / 100>(a, s, e, lane, m, grec, scr, a.np_queue, a.np_count, a.np_capacity);
}

// =================================================================================================================
// narrowphase: one thread per queued candidate pair (GJK / SAT / plane), results into the environment's result slots
// =================================================================================================================
namespace {
__device__ __forceinline__ void np_load_shape(const KM& m, const float* scr, int si, WShape& w, Q4& q) {
    const AvgShape* S = &m.shape[si];
    w.s = S; w.verts = m.vert + 4 * S->vert_off; w.planes = m.plane + 4 * S->plane_off;
    if (S->body >= 0) {                     // top-level moving shapes and children of compounds: on a dynamic or env-static body
        const float4* gp = reinterpret_cast<const float4*>(scr + AVG_S_POSE) + 2 * S->body;
        const float4 p4 = gp[0], q4 = gp[1];
        const Q4 bq = mkq(q4.x, q4.y, q4.z, q4.w);
        w.p = mk3(p4.x, p4.y, p4.z) + qrot(bq, ld3(S->pos));
        q = qnormalize(qmul(bq, ldq(S->quat)));
    } else { w.p = ld3(S->pos); q = ldq(S->quat); }
    const M3 r = qmat(q);
#pragma unroll
    for (int i = 0; i < 9; ++i) w.R[i] = r.m[i];
}
}  // namespace

// One slab of the narrowphase: lane l < ipw of a converged warp takes item base + l of `queue` (items [0, count)).
__device__ __forceinline__ void narrow_slab(const AvgStepArgs& a, const AvgNpItem* queue, const int base, const int count, const int ipw, const int lane) {
    {
        const int i = base + lane;
        const bool valid = lane < ipw && i < count;
        const long long t_begin = a.dbg_counters ? clock64() : 0;
        AvgNpItem it; it.env = 0; it.pair = 0; it.slot = 0; it.cert = -1;
        if (valid) it = queue[i];
        const int variant = (valid && a.variant) ? a.variant[it.env] : 0;
        const KM m = c_models[a.slot][variant];
        float* scr = a.scratch + (size_t)it.env * AVG_S_STRIDE;
        int sa = it.pair & 0xffff, sb = it.pair >> 16;
        // particle items (Feeding / Drinking): shape a = the sphere template at the particle's position, result into the
        // particle scratch arena
        const bool is_part = valid && (sa & AVG_NP_PARTICLE) != 0;
        const int pidx = sa & 0x7f;
        const int n_all = m.h->n_shape + m.h->n_cshape;
        if (valid && ((!is_part && (sa >= n_all || m.shape[sa].body < 0)) || (is_part && (!a.part || m.h->pshape < 0 || pidx >= m.h->n_particle)) || sb >= n_all ||
                      it.slot < 0 || it.slot >= (is_part ? AVG_MAX_PCAND : AVG_S_NQMAX) || it.cert >= AVG_S_NSEPMAX || it.env < 0 || it.env >= a.n_env)) {
            if (a.dbg & 64) printf("[avg_narrow] bad work item %d of %d: env %d pair %08x slot %d cert %d\n", i, count, it.env, it.pair, it.slot, it.cert);
            sa = 0; sb = m.h->n_mshape; it.slot = 0; it.cert = -1; it.env = a.env_begin;     // never dereference a corrupt item
        }
        WShape A, B; Q4 qa, qb;
        if (is_part) {
            const float* pr = a.part + (size_t)it.env * AVG_P_STRIDE;
            A.s = &m.shape[m.h->pshape]; A.verts = m.vert; A.planes = m.plane;
            A.p = mk3(pr[AVG_P_POS + pidx], pr[AVG_P_POS + 64 + pidx], pr[AVG_P_POS + 128 + pidx]);
            qa = mkq(0, 0, 0, 1);
#pragma unroll
            for (int k = 0; k < 9; ++k) A.R[k] = (k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f;
        } else np_load_shape(m, scr, sa, A, qa);
        np_load_shape(m, scr, sb, B, qb);
        const float thr = fminf(A.s->thr, B.s->thr), ma = A.s->margin, mb = B.s->margin;
        const bool plane = B.s->type == AVG_SHAPE_PLANE;
        float4* gsep = reinterpret_cast<float4*>(scr + AVG_S_SEP);
        V3 pa = mk3(0, 0, 0), pb = pa, n = pa; float d = 0.0f;
        bool hit = false, have_sep = false, done = !valid;
        V3 sepv = mk3(0, 0, 0); float sepgap = 0.0f;
        int gi = 0;
        // (2) support-plane test along the remembered direction; ground-plane pairs: one support along -z
        {
            V3 v = mk3(0, 0, -1);
            bool test = false;
            if (valid && !plane && it.cert >= 0) {
                const float4 hint = gsep[it.cert];
                const V3 va = mk3(hint.x, hint.y, hint.z);
                if (dot(va, va) > 0.0f) { v = qrot(qa, va); test = true; }
            }
            const V3 s1 = support_any(A, plane ? v : -v, (test || plane) && valid, lane);
            const V3 s2 = support_any(B, v, test, lane);
            if (test) {
                const V3 w = s1 - s2;
                const float vv = dot(v, v), vw = dot(v, w), md = thr + ma + mb;
                if (vw > 0.0f && vw * vw > md * md * vv) { done = true; have_sep = true; sepv = v; sepgap = vw * rsqrtf(vv); }
            }
            if (valid && plane) {                                     // narrowphase(): half-space of plane.urdf's box
                done = true;
                d = s1.z - ma;
                if (d < thr) { hit = true; n = mk3(0, 0, 1); pa = mk3(s1.x, s1.y, s1.z - ma); pb = mk3(s1.x, s1.y, 0); }
            }
        }
        // GJK on the cores, margins added afterwards (narrowphase())
        {
            float dist = 0.0f; V3 ca = pa, cb = pb;
            const int g = gjk_lockstep(A, B, !done, lane, thr + ma + mb, dist, ca, cb, sepv, sepgap, gi);
            if (g == 2) have_sep = true;
            else if (g == 0) {
                d = dist - ma - mb;
                if (d >= thr) have_sep = sepgap > 0.0f;
                else { hit = true; n = (ca - cb) * (1.0f / dist); pa = ca - n * ma; pb = cb + n * mb; }
            }
            float best = 3.0e38f; V3 bn = mk3(0, 0, 1), bpa = A.p;
            sat_served(A, B, g == 1, lane, best, bn, bpa);          // cores overlap: face-normal SAT (rare)
            if (g == 1) {
                gi |= 1 << 16;
                if (best > 1.0e38f) best = 0;
                hit = true; n = bn; d = -best - ma - mb;
                pa = bpa - n * ma; pb = bpa + n * best + n * mb;
            }
        }
        if (valid) {
            if (a.dbg_counters) {
                atomicAdd(a.dbg_counters, 1ull); if (have_sep && gi == 0) atomicAdd(a.dbg_counters + 1, 1ull);
                if (gi) { atomicAdd(a.dbg_counters + 2, 1ull); atomicAdd(a.dbg_counters + 3, (unsigned long long)(gi & 0xffff)); if (gi >> 16) atomicAdd(a.dbg_counters + 4, 1ull); }
                if (hit) atomicAdd(a.dbg_counters + 5, 1ull);
                const unsigned long long cyc = (unsigned long long)(clock64() - t_begin);
                atomicAdd(a.dbg_counters + 6, cyc); atomicMax(a.dbg_counters + 7, cyc);
            }
            if (it.cert >= 0 && have_sep && sepgap > 0.0f) {
                const V3 va = qrot_inv(qa, sepv), prel = qrot_inv(qa, B.p - A.p);
                const Q4 qrel = qmul(qconj(qa), qb);
                gsep[it.cert] = make_float4(va.x, va.y, va.z, __uint_as_float(it.pair));
                gsep[AVG_S_NSEPMAX + it.cert] = make_float4(prel.x, prel.y, prel.z, sepgap);
                gsep[2 * AVG_S_NSEPMAX + it.cert] = make_float4(qrel.x, qrel.y, qrel.z, qrel.w);
            }                                                        // else the slot keeps the "no certificate" seed of the collide kernel
            if (is_part) {
                float4* r = reinterpret_cast<float4*>(a.pscratch + (size_t)it.env * AVG_PS_STRIDE + AVG_PS_CAND) + 2 * it.slot;
                r[0] = make_float4(n.x, n.y, n.z, d);
                r[1] = make_float4(__int_as_float(pidx), __int_as_float(sb), hit ? 1.0f : 0.0f, 0.0f);
            } else {
                float4* r = reinterpret_cast<float4*>(scr + AVG_S_NPRES) + 4 * it.slot;
                r[0] = make_float4(pa.x, pa.y, pa.z, pb.x);
                r[1] = make_float4(pb.y, pb.z, n.x, n.y);
                r[2] = make_float4(n.z, d, __int_as_float(sa), __int_as_float(sb));
                r[3] = make_float4(hit ? 1.0f : 0.0f, 0, 0, 0);
            }
        }
        __syncwarp();
    }
}
// OCC = blocks per SM the instance is compiled for: 3 (151 registers, nothing spilled) for small batches, where the step waits
// for the slowest warp of this kernel, AVG_OCC_NARROW (85 registers, a few spills) for large ones, where more resident warps
// hide the latency of the lockstep GJK (+4 % on an in-contact policy rollout at 196608 environments).
template <int OCC>
__global__ void __launch_bounds__(128, OCC)
avg_narrow_kernel(AvgStepArgs a) {
    // one THREAD per work item; the 32 items of a warp advance through the certificate test and the GJK iterations in
    // lockstep so that hull support scans can be served by the whole warp (support_any)
    const int count = min(a.np_count[0], a.np_capacity);
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    // Items per warp.  A full queue gives every warp 32 items (throughput: the lockstep iterations are shared by 32 pairs).  A
    // short queue (small batches: ~1 item per environment and sub-step) is spread over the warps the GPU can hold at once
    // instead -- a warp with few items finishes its lockstep GJK in a fraction of the time (fewer iterations to wait for, every
    // hull scan served by the whole warp), and at a small batch the step waits for the slowest warp of this kernel.
    // (Only at small batches: spreading costs instructions -- fewer pairs share a lockstep iteration -- and at a large batch this
    // kernel overlaps with the other half's kernels, so there its instruction count matters, not its latency: -3 % when spread.)
    const int ipw = (a.env_end - a.env_begin) > 16384 ? 32 : min(32, max(1, (count + AVG_NARROW_WARPS - 1) / AVG_NARROW_WARPS));
    for (int base = wid * ipw; base < count; base += nwarps * ipw) narrow_slab(a, a.np_queue, base, count, ipw, lane);
}

// J.dv over the warp, the inner operation of every dense Gauss-Seidel row.  A float butterfly costs five DEPENDENT shuffle +
// add stages (~150 cycles on the row-to-row dependency chain that bounds this kernel, profiles/ncu_full_r1p.txt).
// AVG_REDUX = 2 (default): two redux.sync.  The first takes the largest |product| of the warp (max over the float bit
//   patterns, monotonic for non-negative floats), which fixes a power-of-two scale such that every lane's product rounds
//   to an integer below 2^25 -- the 24-25 significant bits a float sum keeps, at any magnitude -- and the 32-lane sum stays
//   below 2^30; the second adds the integers, exactly and order-independently (bit-deterministic by construction).
// AVG_REDUX = 1: one redux.sync at a fixed scale of 2^-22 m/s.  3 % faster, but the PR2's finger tips (1.5e-5 kg m^2 under a
//   500 N m motor) reach intermediate J.dv of tens of m/s inside a sweep, so a fixed range is not safe for every model.
// AVG_REDUX = 0: the float butterfly.
__device__ __forceinline__ float dense_dot(float j, float dv) {
#if AVG_REDUX == 1
    return (float)__reduce_add_sync(AVG_FULL, __float2int_rn(j * dv * 4194304.0f)) * (1.0f / 4194304.0f);
#elif AVG_REDUX == 2
    const float x = j * dv;
    const unsigned m = __reduce_max_sync(AVG_FULL, __float_as_uint(x) & 0x7fffffffu);
    const unsigned ee = max(m & 0x7f800000u, 40u << 23);                     // e << 23, e = biased exponent of the largest |x| (|x| < 2^(e - 126) in every lane)
    const float scale = __uint_as_float((278u << 23) - ee);                  // 2^(151 - e)
    const float inv = __uint_as_float(ee - (24u << 23));                     // 2^(e - 151)
    return (float)__reduce_add_sync(AVG_FULL, __float2int_rn(x * scale)) * inv;
#else
    return warp_sum(j * dv);
#endif
}

// =================================================================================================================
// dynamics + constraint rows -> row arena
// =================================================================================================================
// Body of the dynamics kernel for ONE environment (one warp).  PF_AHEAD: distance (in environments) of the L2 prefetch hints.
template <int MAXBLK, int PF_AHEAD>
__device__ __forceinline__ void dynamics_body(const AvgStepArgs& a, SmDyn& s, const int e, const int lane, const KM& m, float* grec, float* scr) {
    const AvgModelHeader* h = m.h;
    const int nb = h->n_body, nj = h->n_jdof, nd = h->n_dof;
    const float dt = h->dt;
    const V3 ref = mk3(h->task_f[16], h->task_f[17], h->task_f[18]);
    int* scr_i = reinterpret_cast<int*>(scr);
    for (int i = lane; i < AVG_E_EBODY; i += 32) s.env[i] = grec[i];
    {   // what this kernel reads first, for the environment whose warp takes this slot next: record (4 lines), counters (1),
        // body poses (6), narrowphase results (4)
        const int ea = e + PF_AHEAD;
        if (PF_AHEAD > 0 && ea < a.env_end) {
            const float* r2 = a.env + (size_t)ea * AVG_ENV_STRIDE; const float* s2 = a.scratch + (size_t)ea * AVG_S_STRIDE;
            const float* p = lane < 4 ? r2 + 32 * lane : (lane < 5 ? s2 : (lane < 11 ? s2 + AVG_S_POSE + 32 * (lane - 5) : (lane < 15 ? s2 + AVG_S_NPRES + 32 * (lane - 11) : nullptr)));   // 24 bodies, 8 results: what is usually there
            if (p) prefetch_l2(p);
        }
    }
    int overflow = 0;
    // the narrowphase queue has been drained by the kernel before this one; empty it for the next sub-step's collide kernel
    if (e == a.env_begin && lane == 0) a.np_count[0] = 0;
    AVG_MASK_CHECK
    // contacts: the hits among the narrowphase results, compacted in pair order; the list is also written to the arena
    // for the solver (impulses) and the epilogue
    int ncontact;
    {
        const int nq = scr_i[AVG_S_NQ];
        ncontact = 0;
#pragma unroll 1
        for (int base = 0; base < nq; base += 32) {
            const float4* r = reinterpret_cast<const float4*>(scr + AVG_S_NPRES) + 4 * (base + lane);
            float4 r0 = make_float4(0, 0, 0, 0), r1 = r0, r2 = r0;
            bool hit = false;
            if (base + lane < nq) { hit = r[3].x != 0.0f; if (hit) { r0 = r[0]; r1 = r[1]; r2 = r[2]; } }
            const unsigned bal = __ballot_sync(AVG_FULL, hit);
            const int slot = ncontact + __popc(bal & ((1u << lane) - 1));
            if (hit && slot < kMaxC) {
                s.c_pa[slot][0] = r0.x; s.c_pa[slot][1] = r0.y; s.c_pa[slot][2] = r0.z;
                s.c_pb[slot][0] = r0.w; s.c_pb[slot][1] = r1.x; s.c_pb[slot][2] = r1.y;
                s.c_n[slot][0] = r1.z; s.c_n[slot][1] = r1.w; s.c_n[slot][2] = r2.x;
                s.c_dist[slot] = r2.y; s.c_sa[slot] = __float_as_int(r2.z); s.c_sb[slot] = __float_as_int(r2.w);
                float* c = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * slot;
                c[0] = r0.x; c[1] = r0.y; c[2] = r0.z; c[3] = r0.w; c[4] = r1.x; c[5] = r1.y; c[6] = r1.z; c[7] = r1.w; c[8] = r2.x;
                c[9] = r2.y; c[10] = r2.z; c[11] = r2.w; c[12] = 0.0f;
                // warm start (btMultiBodyConstraintSolver::setupMultiBodyContactConstraint): the impulse the same shape pair carried
                // in the last internal step (AVG_E_WCACHE), times the warm-starting factor, initialises the normal row in the solver
                float lam0 = 0.0f;
                if (h->warmstart > 0.0f) {
                    const uint32_t key = (uint32_t)__float_as_int(r2.z) | ((uint32_t)__float_as_int(r2.w) << 16);
#pragma unroll
                    for (int w = AVG_WCACHE_N - 1; w >= 0; --w)
                        if (__float_as_uint(grec[AVG_E_WCACHE + 2 * w]) == key) lam0 = h->warmstart * grec[AVG_E_WCACHE + 2 * w + 1];
                }
                c[13] = lam0;
            }
            ncontact += __popc(bal);
        }
        if (ncontact > kMaxC) { overflow |= 1; ncontact = kMaxC; }        // flagged (AVG_E_OVERFLOW, info), never silent
        if (lane == 0) scr_i[AVG_S_NC] = ncontact;
    }
    __syncwarp();
    if (lane < nb + h->n_ebody) {            // body poses: forward kinematics was done by the collide kernel
        const float4* gp = reinterpret_cast<const float4*>(scr + AVG_S_POSE) + 2 * lane;
        const float4 p4 = gp[0], q4 = gp[1];
        s.bp[lane][0] = p4.x; s.bp[lane][1] = p4.y; s.bp[lane][2] = p4.z;
        s.bq[lane][0] = q4.x; s.bq[lane][1] = q4.y; s.bq[lane][2] = q4.z; s.bq[lane][3] = q4.w;
    }
    __syncwarp();
    float* gJ = scr + AVG_S_J; float* gW = scr + AVG_S_W;

    // ---- per-lane body quantities ------------------------------------------------------------------------------
    LaneDyn L; L.S = mksv(mk3(0, 0, 0), mk3(0, 0, 0)); L.anc = 0;
    float qd = (lane < nd) ? s.env[AVG_E_QD + lane] : 0.0f;
    Inertia I; I.m = 0; I.h = mk3(0, 0, 0); I.xx = I.yy = I.zz = I.xy = I.xz = I.yz = 0;
    float Ic[6] = {0, 0, 0, 0, 0, 0};        // world rotational inertia about the COM (xx, yy, zz, xy, xz, yz)
    V3 c = mk3(0, 0, 0);
    int parent = -1; bool is_joint = false, is_free = false;
    V3 grav = mk3(0, 0, 0);
    float mass = 0;
    if (lane < nb) {
        const AvgBody* B = &m.body[lane];
        L.anc = B->anc_mask; parent = B->parent; mass = B->mass; grav = ld3(B->gravity);
        V3 p = ld3(s.bp[lane]); Q4 q = ldq(s.bq[lane]);
        M3 R = qmat(q);
        float i0 = B->inertia[0], i1 = B->inertia[1], i2 = B->inertia[2];
        // changeDynamics(mass=0) per episode (world_creation.py:157-161: the head chain of Feeding / Drinking unless the episode
        // drew a tremor): a massless subtree gives M_kk = 0, which the inverse below turns into a frozen dof
        if ((__float_as_uint(grec[AVG_E_FROZEN]) >> lane) & 1u) { mass = 0.0f; i0 = i1 = i2 = 0.0f; }
        Ic[0] = R.m[0] * R.m[0] * i0 + R.m[1] * R.m[1] * i1 + R.m[2] * R.m[2] * i2;
        Ic[1] = R.m[3] * R.m[3] * i0 + R.m[4] * R.m[4] * i1 + R.m[5] * R.m[5] * i2;
        Ic[2] = R.m[6] * R.m[6] * i0 + R.m[7] * R.m[7] * i1 + R.m[8] * R.m[8] * i2;
        Ic[3] = R.m[0] * R.m[3] * i0 + R.m[1] * R.m[4] * i1 + R.m[2] * R.m[5] * i2;
        Ic[4] = R.m[0] * R.m[6] * i0 + R.m[1] * R.m[7] * i1 + R.m[2] * R.m[8] * i2;
        Ic[5] = R.m[3] * R.m[6] * i0 + R.m[4] * R.m[7] * i1 + R.m[5] * R.m[8] * i2;
        c = p - ref;
        if (B->jtype == AVG_JOINT_FREE) {
            is_free = true;
            // inverse world inertia for the solver
            int fb = (B->dof - nj) / 6;
            float j0 = i0 > 0 ? 1.0f / i0 : 0.0f, j1 = i1 > 0 ? 1.0f / i1 : 0.0f, j2 = i2 > 0 ? 1.0f / i2 : 0.0f;
            float* fi = s.freeInv[fb];
            fi[0] = mass > 0 ? 1.0f / mass : 0.0f;
            fi[1] = R.m[0] * R.m[0] * j0 + R.m[1] * R.m[1] * j1 + R.m[2] * R.m[2] * j2;
            fi[2] = R.m[0] * R.m[3] * j0 + R.m[1] * R.m[4] * j1 + R.m[2] * R.m[5] * j2;
            fi[3] = R.m[0] * R.m[6] * j0 + R.m[1] * R.m[7] * j1 + R.m[2] * R.m[8] * j2;
            fi[4] = fi[2];
            fi[5] = R.m[3] * R.m[3] * j0 + R.m[4] * R.m[4] * j1 + R.m[5] * R.m[5] * j2;
            fi[6] = R.m[3] * R.m[6] * j0 + R.m[4] * R.m[7] * j1 + R.m[5] * R.m[8] * j2;
            fi[7] = fi[3]; fi[8] = fi[6];
            fi[9] = R.m[6] * R.m[6] * j0 + R.m[7] * R.m[7] * j1 + R.m[8] * R.m[8] * j2;
        } else {
            is_joint = true;
            // joint axis / origin from the body pose: axis_body = R(tb)^T axis, origin_body = -R(tb)^T tb_pos
            Q4 tbq = ldq(B->tb_quat);
            V3 ax_w = qrot(q, qrot_inv(tbq, ld3(B->axis)));
            V3 org_w = p - qrot(q, qrot_inv(tbq, ld3(B->tb_pos)));
            if (B->jtype == AVG_JOINT_REVOLUTE) L.S = mksv(ax_w, cross(org_w - ref, ax_w));
            else L.S = mksv(mk3(0, 0, 0), ax_w);
            I.m = mass; I.h = c * mass;
            float cc = dot(c, c);
            I.xx = Ic[0] + mass * (cc - c.x * c.x); I.yy = Ic[1] + mass * (cc - c.y * c.y); I.zz = Ic[2] + mass * (cc - c.z * c.z);
            I.xy = Ic[3] - mass * c.x * c.y; I.xz = Ic[4] - mass * c.x * c.z; I.yz = Ic[5] - mass * c.y * c.z;
        }
    }
    __syncwarp();

    // ---- body velocities and velocity-product accelerations: path sums by pointer jumping -----------------------
    Sv vj = L.S * (is_joint ? qd : 0.0f);
    Sv V = vj;
    {
        int anc = is_joint ? parent : -1;
        while (__any_sync(AVG_FULL, anc >= 0)) {
            int src = anc >= 0 ? anc : lane;
            Sv av = shflsv(V, src);
            int aa = __shfl_sync(AVG_FULL, anc, src);
            if (anc >= 0) { V = V + av; anc = aa; }
        }
    }
    if (h->n_particle > 0) {                 // start-of-step body twists for the one-way particle-vs-link contacts (solve kernel)
        float4* tw = reinterpret_cast<float4*>(scr + AVG_S_TWIST) + 2 * lane;
        tw[0] = make_float4(V.a.x, V.a.y, V.a.z, 0.0f); tw[1] = make_float4(V.l.x, V.l.y, V.l.z, 0.0f);
    }
    Sv Ab = crm(V, vj);
    {
        int anc = is_joint ? parent : -1;
        while (__any_sync(AVG_FULL, anc >= 0)) {
            int src = anc >= 0 ? anc : lane;
            Sv av = shflsv(Ab, src);
            int aa = __shfl_sync(AVG_FULL, anc, src);
            if (anc >= 0) { Ab = Ab + av; anc = aa; }
        }
    }
    // ---- body bias force: I Ab + V x* I V - external (gravity, Bullet velocity damping) -------------------------
    Sv fb = mksv(mk3(0, 0, 0), mk3(0, 0, 0));
    if (is_joint) {
        Sv IV = inertia_mul(I, V);
        fb = inertia_mul(I, Ab) + crf(V, IV);
        V3 w = V.a;
        V3 vc = V.l + cross(w, c);
        V3 f = grav * mass - vc * (mass * (h->lin_damp + h->lin_damp * norm(vc)));
        V3 Iw = mk3(Ic[0] * w.x + Ic[3] * w.y + Ic[4] * w.z, Ic[3] * w.x + Ic[1] * w.y + Ic[5] * w.z, Ic[4] * w.x + Ic[5] * w.y + Ic[2] * w.z);
        V3 n = -(Iw * (h->ang_damp + h->ang_damp * norm(w)));
        fb = fb - mksv(n + cross(c, f), f);
    }
    // ---- subtree sums (composite inertia, bias force).  Bodies are in depth-first order, so the subtree of body i is
    //      the lane range [i, sub_end): every lane stages its 16 values in shared memory and adds up its own range ------
    Inertia Ic_sub = I; Sv f_sub = fb;
    {
        int sub_end = 0;
        if (is_joint) {
            sub_end = m.body[lane].sub_end;
            s.sub[lane][0] = make_float4(I.m, I.h.x, I.h.y, I.h.z);
            s.sub[lane][1] = make_float4(I.xx, I.yy, I.zz, I.xy);
            s.sub[lane][2] = make_float4(I.xz, I.yz, fb.a.x, fb.a.y);
            s.sub[lane][3] = make_float4(fb.a.z, fb.l.x, fb.l.y, fb.l.z);
        }
        __syncwarp();
        for (int j = lane + 1; j < sub_end; ++j) {
            const float4 v0 = s.sub[j][0], v1 = s.sub[j][1], v2 = s.sub[j][2], v3 = s.sub[j][3];
            Ic_sub.m += v0.x; Ic_sub.h.x += v0.y; Ic_sub.h.y += v0.z; Ic_sub.h.z += v0.w;
            Ic_sub.xx += v1.x; Ic_sub.yy += v1.y; Ic_sub.zz += v1.z; Ic_sub.xy += v1.w;
            Ic_sub.xz += v2.x; Ic_sub.yz += v2.y; f_sub.a.x += v2.z; f_sub.a.y += v2.w;
            f_sub.a.z += v3.x; f_sub.l.x += v3.y; f_sub.l.y += v3.z; f_sub.l.z += v3.w;
        }
        __syncwarp();
    }
    Sv F = inertia_mul(Ic_sub, L.S);          // composite inertia times own motion subspace
    float Cb = dot(L.S, f_sub);               // generalized bias force of this lane's joint
    if (lane < nj) Cb = fmaf(m.dof[lane].damping, qd, Cb);   // joint damping torque -damping * qd (URDF <dynamics damping>, PR2)
    // ---- joint-space mass matrix in registers: lane c keeps column c of its articulation's diagonal block,
    //      mc[t] = M[bs+t][c] = S_row . (Ic_deeper S_deeper) (symmetric), one block per articulation -----------------
    int bs = lane, be = lane;
    for (int b = 0; b < h->n_block; ++b) {
        const int b0 = h->block_start[b], b1 = h->block_start[b + 1];
        if (lane >= b0 && lane < b1) { bs = b0; be = b1; }
    }
    float mc[MAXBLK];
#pragma unroll
    for (int t = 0; t < MAXBLK; ++t) {
        const bool valid = bs + t < be;
        const int j = valid ? bs + t : lane;
        const Sv Fj = shflsv(F, j), Sj = shflsv(L.S, j);
        const uint32_t mj = __shfl_sync(AVG_FULL, L.anc, j);
        float v = 0.0f;
        if ((mj >> lane) & 1u) v = dot(L.S, Fj);            // this lane's joint is an ancestor-or-self of j
        else if ((L.anc >> j) & 1u) v = dot(Sj, F);         // j is an ancestor of this lane's joint
        mc[t] = valid ? v : 0.0f;
    }
    // ---- in-place Gauss-Jordan inverse (SPD, no pivoting), all blocks swept concurrently; the pivot column comes from
    //      the pivot's lane by shuffle; frozen dofs (M_kk ~ 0) invert to 0 -------------------------------------------
#pragma unroll
    for (int kk = 0; kk < MAXBLK; ++kk) {
        const bool act = bs + kk < be;
        const int k = act ? bs + kk : lane;
        const float p = __shfl_sync(AVG_FULL, mc[kk], k);
        const float ip = p > 1e-20f ? 1.0f / p : 0.0f;
        const float mkc = mc[kk];
        const float f = mkc * ip;
#pragma unroll
        for (int t = 0; t < MAXBLK; ++t) {
            if (t == kk) continue;
            const float mik = __shfl_sync(AVG_FULL, mc[t], k);
            if (act) mc[t] = (lane != k) ? fmaf(-mik, f, mc[t]) : -mc[t] * ip;
        }
        if (act) mc[kk] = (lane != k) ? f : ip;
    }
    float mdiag = 0.0f;
#pragma unroll
    for (int t = 0; t < MAXBLK; ++t) if (bs + t == lane) mdiag = mc[t];
    // ---- unconstrained velocity update: qd* = qd + dt M^-1 (-C) ---------------------------------------------------
    float qdd = 0.0f;
#pragma unroll
    for (int t = 0; t < MAXBLK; ++t) {
        const float cj = __shfl_sync(AVG_FULL, Cb, (bs + t) & 31);      // slots past the end of the block hold mc[t] = 0 (finite Cb): no select needed
        qdd = fmaf(mc[t], -cj, qdd);
    }
    // free bodies (lane b computes, dof lanes pick up through smem scratch in s.obs)
    if (is_free) {
        const AvgBody* B = &m.body[lane];
        const float* v6 = s.env + AVG_E_QD + B->dof;
        V3 v = ld3(v6), w = ld3(v6 + 3);
        V3 a = mass > 0 ? grav - v * (h->lin_damp + h->lin_damp * norm(v)) : mk3(0, 0, 0);
        V3 Iw = mk3(Ic[0] * w.x + Ic[3] * w.y + Ic[4] * w.z, Ic[3] * w.x + Ic[1] * w.y + Ic[5] * w.z, Ic[4] * w.x + Ic[5] * w.y + Ic[2] * w.z);
        V3 tau = -(Iw * (h->ang_damp + h->ang_damp * norm(w))) - cross(w, Iw);
        const float* fi = s.freeInv[(B->dof - nj) / 6];
        V3 al = mk3(fi[1] * tau.x + fi[2] * tau.y + fi[3] * tau.z, fi[4] * tau.x + fi[5] * tau.y + fi[6] * tau.z, fi[7] * tau.x + fi[8] * tau.y + fi[9] * tau.z);
        float* o = s.tmp + (B->dof - nj);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = al.x; o[4] = al.y; o[5] = al.z;
    }
    __syncwarp();
    if (lane >= nj && lane < nd) qdd = s.tmp[lane - nj];
    __syncwarp();
    qd = qd + dt * qdd;

    // ---- constraint rows ------------------------------------------------------------------------------------------
    // Unit rows (+-e_i) are stored per dof: a motor row and at most one limit row (a dof can violate one side only).
    // Rows of different articulations do not couple (block-diagonal M^-1), so the solver sweeps the blocks
    // concurrently; inside a block the order is Bullet's: motors in dof order, then limits in dof order.
    int nlim;
    {
        float4 ma = make_float4(0, 0, 0, 0), la = ma;
        bool has_lim = false;
        if (lane < nj) {
            const AvgDof* D = &m.dof[lane];
            const float diag = mdiag;
            const float inv = diag > 1e-12f ? 1.0f / diag : 0.0f;
            const float q = s.env[AVG_E_Q + m.body[D->body].qidx];
            if (D->flags & AVG_DOF_MOTOR) {        // btMultiBodyJointMotor: velocity target kp (q*-q)/dt - kd qd, clamp force*dt
                const bool hum = D->flags & AVG_DOF_HUMAN;
                const float kp = hum ? s.env[AVG_E_HUMAN_KP] : D->kp;
                const float maxf = hum ? h->task_f[AVG_TF_HUMAN_FORCE] * s.env[AVG_E_STRENGTH] : D->max_force;
                ma = make_float4(kp * (s.env[AVG_E_MTARGET + lane] - q) / dt - D->kd * qd, inv, maxf * dt, diag);
            }
            if (D->flags & AVG_DOF_LIMIT) {        // btMultiBodyJointLimitConstraint: only while violated, [0, 100]
                const float sc = (D->flags & AVG_DOF_HUMAN) ? s.env[AVG_E_LIMIT_SCALE] : 1.0f;
                const float pen0 = q - D->lower * sc, pen1 = D->upper * sc - q;
                if (pen0 <= 0) { has_lim = true; la = make_float4(-pen0 * h->erp / dt - qd, inv, diag, 1.0f); }
                else if (pen1 <= 0) { has_lim = true; la = make_float4(-pen1 * h->erp / dt + qd, inv, diag, -1.0f); }
            }
        }
        float4* g_um = reinterpret_cast<float4*>(scr + AVG_S_ROWS_M);
        float4* g_ul = reinterpret_cast<float4*>(scr + AVG_S_ROWS_L);
        g_um[lane] = ma; g_ul[lane] = la;
        nlim = __popc(__ballot_sync(AVG_FULL, has_lim));
    }
    float4* g_rows = reinterpret_cast<float4*>(scr + AVG_S_ROWS_D);      // dense rows, indexed by dense row number
    // ---- dense rows in Bullet's order: tool weld (btMultiBodyFixedConstraint, 3 + 3 rows), contact normals, then one
    //      friction row per contact.  One loop, so the M^-1 J^T code exists once. -------------------------------------
    V3 wpa, wpb, perr, rotv;
    {
        Q4 qa, qb;
        frame_pose(m, s, AVG_F_WELD_PARENT, wpa, qa);
        frame_pose(m, s, AVG_F_TOOL_BASE, wpb, qb);
        Q4 dq = qmul(qa, qconj(qb));
        if (dq.w < 0) dq = mkq(-dq.x, -dq.y, -dq.z, -dq.w);
        const float sn = sqrtf(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z);
        rotv = mk3(0, 0, 0);
        if (sn > 1e-9f) { const float ang = 2.0f * atan2f(sn, dq.w) / sn; rotv = mk3(dq.x * ang, dq.y * ang, dq.z * ang); }
        perr = wpa - wpb;
    }
    const float maxi = h->weld_max_force * dt;
    const int nc = ncontact;                   // kMaxDense = 6 + 2 * AVG_MAX_CONTACT always has room
    const int first_contact_row = 6;
    const int ndense = 6 + 2 * nc;
    // free-body lanes: which free body, which component
    int fbi = 0, fk = -1, fbase = lane;
    if (lane >= nj && lane < nd) { fk = lane - nj; while (fk >= 6) { fk -= 6; fbi++; } fbase = lane - fk; }
    // This lane's column of the Jacobian of (linear velocity of world point r on `body`, angular velocity of `body`):
    // every row through that point is a dot product with it, so the six weld rows cost two evaluations and a contact
    // (normal + friction) two more.
    auto jcol = [&](int body, V3 r, V3& cl, V3& cw) {
        cl = mk3(0, 0, 0); cw = mk3(0, 0, 0);
        if (body < 0 || body >= nb) return;              // static world / env-static body
        const AvgBody* B = &m.body[body];
        if (B->jtype == AVG_JOINT_FREE) {
            const int k = lane - B->dof;
            if (k >= 0 && k < 3) cl = mk3(k == 0, k == 1, k == 2);
            else if (k >= 3 && k < 6) { cw = mk3(k == 3, k == 4, k == 5); cl = cross(cw, r - ld3(s.bp[body])); }
        } else if (lane < nj && ((B->anc_mask >> lane) & 1u)) { cw = L.S.a; cl = L.S.l + cross(L.S.a, r - ref); }
    };
    V3 wcl, wcw;
    {
        V3 la, wa, lb, wb;
        jcol(h->weld_body_a, wpa, la, wa); jcol(h->weld_body_b, wpb, lb, wb);
        wcl = la - lb; wcw = wa - wb;
    }
    // The six weld rows are always there and share everything but their Jacobian component, so they are built together:
    // W = M^-1 J^T with one index computation per column of M^-1, and their twelve warp reductions (J.W and J.qd* per
    // row) as ONE packed butterfly (pack2: at every stage two partial sums share a shuffle, half of the lanes keep one,
    // half the other) instead of twelve: 13 + 12 shuffles instead of 120.
#if AVG_WELD_BATCH
    {
        const float jl6[6] = {wcl.x, wcl.y, wcl.z, wcw.x, wcw.y, wcw.z};
        float w6[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int t = 0; t < MAXBLK; ++t) {
            const int src = (bs + t) & 31;                       // mc[t] = 0 past the end of the block
#pragma unroll
            for (int d = 0; d < 6; ++d) w6[d] = fmaf(mc[t], __shfl_sync(AVG_FULL, jl6[d], src), w6[d]);
        }
        {
            const int s3 = (fbase + 3) & 31, s4 = (fbase + 4) & 31, s5 = (fbase + 5) & 31;
            const float* fi = s.freeInv[fbi];
            const int r = fk >= 3 ? fk - 3 : 0;
            const float f0 = fk >= 0 ? fi[0] : 0.0f, f1 = fk >= 3 ? fi[1 + 3 * r] : 0.0f, f2 = fk >= 3 ? fi[2 + 3 * r] : 0.0f, f3 = fk >= 3 ? fi[3 + 3 * r] : 0.0f;
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                const float j3 = __shfl_sync(AVG_FULL, jl6[d], s3), j4 = __shfl_sync(AVG_FULL, jl6[d], s4), j5 = __shfl_sync(AVG_FULL, jl6[d], s5);
                if (fk >= 0) w6[d] = fk < 3 ? f0 * jl6[d] : f1 * j3 + f2 * j4 + f3 * j5;
            }
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) { gJ[d * 32 + lane] = jl6[d]; gW[d * 32 + lane] = w6[d]; }
        auto pack2 = [&](float a, float b, int bit) {           // lanes with `bit` clear end up with a's sum over {lane, lane ^ bit}, the others with b's
            const bool up = (lane & bit) != 0;
            const float keep = up ? b : a, send = up ? a : b;
            return keep + __shfl_xor_sync(AVG_FULL, send, bit);
        };
        // v[k]: k = d -> J_d . W_d, k = 6 + d -> J_d . qd*
        const float a0 = pack2(jl6[0] * w6[0], jl6[1] * w6[1], 16), a1 = pack2(jl6[2] * w6[2], jl6[3] * w6[3], 16), a2 = pack2(jl6[4] * w6[4], jl6[5] * w6[5], 16);
        const float a3 = pack2(jl6[0] * qd, jl6[1] * qd, 16), a4 = pack2(jl6[2] * qd, jl6[3] * qd, 16), a5 = pack2(jl6[4] * qd, jl6[5] * qd, 16);
        const float b0 = pack2(a0, a1, 8), b1 = pack2(a2, a3, 8), b2 = pack2(a4, a5, 8);
        const float c0 = pack2(b0, b1, 4), c1 = b2 + __shfl_xor_sync(AVG_FULL, b2, 4);
        const float e0 = pack2(c0, c1, 2);
        const float red = e0 + __shfl_xor_sync(AVG_FULL, e0, 1);
        // v[k] ends in lane 16 (k & 1) + 8 ((k >> 1) & 1) + 4 ((k >> 2) & 1) for k < 8, and 16 (k & 1) + 8 ((k >> 1) & 1) + 2 for k >= 8
        float diag = 0.0f, u0 = 0.0f;
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            const int kd = d, ku = 6 + d;
            const int ld_ = 16 * (kd & 1) + 8 * ((kd >> 1) & 1) + 4 * ((kd >> 2) & 1);
            const int lu_ = ku < 8 ? 16 * (ku & 1) + 8 * ((ku >> 1) & 1) + 4 * ((ku >> 2) & 1) : 16 * (ku & 1) + 8 * ((ku >> 1) & 1) + 2;
            const float dd = __shfl_sync(AVG_FULL, red, ld_), uu = __shfl_sync(AVG_FULL, red, lu_);
            if (lane == d) { diag = dd; u0 = uu; }
        }
        if (lane < 6) {                                          // lane d writes row d
            const V3 ev = lane < 3 ? perr : rotv;
            const int ax = lane < 3 ? lane : lane - 3;
            const float err = ax == 0 ? ev.x : (ax == 1 ? ev.y : ev.z);
            g_rows[2 * lane] = make_float4(-err * h->erp / dt - u0, diag > 1e-12f ? 1.0f / diag : 0.0f, -maxi, maxi);
            g_rows[2 * lane + 1] = make_float4(diag, 0.0f, __int_as_float(lane), __int_as_float(-1));
        }
    }
#endif
#pragma unroll 1
    for (int d = AVG_WELD_BATCH ? 6 : 0; d < ndense; ++d) {
        float jl, tgt, lo, hi, mu = 0.0f;
        int par = -1;
        if (d < 6) {
            const int ax = d < 3 ? d : d - 3;
            const V3 c = d < 3 ? wcl : wcw, ev = d < 3 ? perr : rotv;
            jl = ax == 0 ? c.x : (ax == 1 ? c.y : c.z);
            const float err = ax == 0 ? ev.x : (ax == 1 ? ev.y : ev.z);
            tgt = -err * h->erp / dt; lo = -maxi; hi = maxi;
        } else {
            const bool fric = d >= 6 + nc;
            const int ci = fric ? d - 6 - nc : d - 6;
            const V3 pa = ld3(s.c_pa[ci]), pb = ld3(s.c_pb[ci]), n = ld3(s.c_n[ci]);
            const int sa = s.c_sa[ci], sb = s.c_sb[ci];
            V3 la, wa, lb, wb;
            jcol(m.shape[sa].body, pa, la, wa); jcol(m.shape[sb].body, pb, lb, wb);
            const V3 cc = la - lb;
            if (!fric) {
                jl = dot(cc, n);
                const float dist = s.c_dist[ci];
                tgt = dist > 0 ? -dist / dt : -dist * h->erp / dt;       // speculative margin / ERP push
                lo = 0.0f; hi = 1e30f;
            } else {
                const V3 vrel = mk3(dense_dot(cc.x, qd), dense_dot(cc.y, qd), dense_dot(cc.z, qd));
                const V3 lat = vrel - n * dot(vrel, n);
                const float ll = norm(lat);
                V3 t;
                if (ll > 1e-6f) t = lat * (1.0f / ll);
                else if (fabsf(n.z) > 0.70710678f) { const float k = rsqrtf(n.y * n.y + n.z * n.z); t = mk3(0, -n.z * k, n.y * k); }
                else { const float k = rsqrtf(n.x * n.x + n.y * n.y); t = mk3(-n.y * k, n.x * k, 0); }
                jl = dot(cc, t);
                tgt = 0.0f; lo = 0.0f; hi = 0.0f; mu = m.shape[sa].friction * m.shape[sb].friction; par = first_contact_row + ci;
            }
        }
        // W = M^-1 J^T: joint lanes use their register column of M^-1, free-body lanes the inverse mass / inertia
        float w = 0.0f;
#pragma unroll
        for (int t = 0; t < MAXBLK; ++t) {
            w = fmaf(mc[t], __shfl_sync(AVG_FULL, jl, (bs + t) & 31), w);   // mc[t] = 0 past the end of the block, jl is finite
        }
        {
            const float j3 = __shfl_sync(AVG_FULL, jl, (fbase + 3) & 31), j4 = __shfl_sync(AVG_FULL, jl, (fbase + 4) & 31), j5 = __shfl_sync(AVG_FULL, jl, (fbase + 5) & 31);
            if (fk >= 0) {
                const float* fi = s.freeInv[fbi];
                if (fk < 3) w = fi[0] * jl;
                else { const int r = fk - 3; w = fi[1 + 3 * r] * j3 + fi[2 + 3 * r] * j4 + fi[3 + 3 * r] * j5; }
            }
        }
        gJ[d * 32 + lane] = jl; gW[d * 32 + lane] = w;
        const float diag = dense_dot(jl, w);
        const float u0 = dense_dot(jl, qd);
        if (lane == 0) {
            g_rows[2 * d] = make_float4(tgt - u0, diag > 1e-12f ? 1.0f / diag : 0.0f, lo, hi);
            g_rows[2 * d + 1] = make_float4(diag, mu, __int_as_float(d), __int_as_float(par));
        }
    }

    // ---- hand-off to the solver -----------------------------------------------------------------------------------
    if (h->n_particle > 0 && lane < 24) scr[AVG_S_FREEINV + lane] = (&s.freeInv[0][0])[lane];
    if (lane < nd) scr[AVG_S_QD + lane] = qd;
#pragma unroll
    for (int t = 0; t < MAXBLK; ++t) scr[AVG_S_MINV + t * 32 + lane] = mc[t];       // [t][lane]: row bs(lane)+t, column lane
    if (lane == 0) {
        scr_i[AVG_S_NR] = ndense; scr_i[AVG_S_NS] = nlim; scr_i[AVG_S_NFR] = first_contact_row + nc; scr_i[AVG_S_FCR] = first_contact_row;
        scr_i[AVG_S_NCS] = nc;
        if (overflow) scr_i[AVG_S_OVERFLOW] |= overflow;
    }
}

template <int MAXBLK>
__global__ void __launch_bounds__(32 * AVG_WPB_DYN, AVG_OCC_DYN)
avg_dynamics_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmDyn)
    dynamics_body<MAXBLK, 148 * AVG_OCC_DYN * AVG_WPB_DYN * AVG_PF_PCT / 100>(a, s, e, lane, m, grec, scr);
}

// enforce_realistic_human_joint_limits, env.py:353-371: Dense 4 -> 64 tanh -> 64 tanh -> 64 tanh -> 1; lane u evaluates
// hidden units u and u + 32, activations are exchanged by shuffles, weight rows are read coalesced (L1-resident).
// Returns the logit (class 1 <=> logit > 0), identical in every lane.
__device__ __forceinline__ void arm_limit_inputs(float tz, float tx, float ty, float qe, float x[4]) {
    const float twopi = 6.28318530717958647692f;                                 // env.py:360-363 (Python % semantics)
    x[0] = fmodf(-tz + twopi, twopi); if (x[0] < 0) x[0] += twopi;
    x[1] = fmodf(tx + twopi, twopi); if (x[1] < 0) x[1] += twopi;
    x[2] = -ty;
    x[3] = fmodf(-qe + twopi, twopi); if (x[3] < 0) x[3] += twopi;
}
__device__ __noinline__ float arm_limit_logit_warp(const float* __restrict__ w, const float x[4], int lane) {
    const float* W1 = w; const float* b1 = W1 + 256; const float* W2 = b1 + 64; const float* b2 = W2 + 4096;
    const float* W3 = b2 + 64; const float* b3 = W3 + 4096; const float* W4 = b3 + 64; const float* b4 = W4 + 64;
    float a0 = __ldg(b1 + lane), a1 = __ldg(b1 + lane + 32);
#pragma unroll
    for (int k = 0; k < 4; ++k) { a0 = fmaf(x[k], __ldg(W1 + k * 64 + lane), a0); a1 = fmaf(x[k], __ldg(W1 + k * 64 + lane + 32), a1); }
    float h0 = tanhf(a0), h1 = tanhf(a1);
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
        const float* W = layer == 0 ? W2 : W3; const float* bb = layer == 0 ? b2 : b3;
        a0 = __ldg(bb + lane); a1 = __ldg(bb + lane + 32);
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const float u = __shfl_sync(AVG_FULL, h0, j), v = __shfl_sync(AVG_FULL, h1, j);
            a0 = fmaf(u, __ldg(W + j * 64 + lane), a0); a1 = fmaf(u, __ldg(W + j * 64 + lane + 32), a1);
            a0 = fmaf(v, __ldg(W + (j + 32) * 64 + lane), a0); a1 = fmaf(v, __ldg(W + (j + 32) * 64 + lane + 32), a1);
        }
        h0 = tanhf(a0); h1 = tanhf(a1);
    }
    return warp_sum(h0 * __ldg(W4 + lane) + h1 * __ldg(W4 + lane + 32)) + __ldg(b4);
}

// =================================================================================================================
// Food / water particles (feeding.py:291-320, drinking.py:291-322) inside the solver kernel.
//
// 8 or 64 free spheres of 5 mm / 1 g.  Their contact rows belong to the same projected Gauss-Seidel sweep as the rows of the
// articulation (a particle on the spoon loads the spoon, the spoon is welded to the gripper), in Bullet's order: ... contact
// normals (articulation, then particles), friction (articulation, then particles).  A particle row touches one particle and,
// on its other side, another particle, the tool (free body, two-way) or something kinematic (static shape, or a robot / human
// link with its start-of-step velocity: one-way).  Inside one block of particle rows (normals / friction) every
// row sees the tool's velocity change as the block began and the block's reactions reach the tool at once (the oracle does the
// same): what a block orders are the rows that share a PARTICLE.  Rows that share no particle commute, so the
// ordered sweep of the oracle is executed here in ROUNDS: a greedy schedule puts every row in the first round after the rows
// before it (canonical contact order) that use one of its particles, <= 32 rows per round, one lane per row -- the same numbers
// as the sequential sweep, up to 32 rows at a time.  Particle velocities and the tool's six velocity components live in shared
// memory during these phases; the per-contact records (20 floats) stay in the particle scratch arena (L1 / L2).
// =================================================================================================================
// NP / NPC: particles / particle contacts the instance holds (Feeding: 8 / 64, Drinking: 64 / AVG_MAX_PCONTACT): the small
// instance keeps the solver kernel's shared memory per warp low enough for the register-limited occupancy.
template <int NP, int NPC>
struct __align__(16) SmPartT {
    static constexpr int kNP = NP, kNPC = NPC;
    float pv[6][NP];                       // unconstrained velocities after gravity / damping (v*, w*), component-major
    float pdv[6][NP];                      // velocity changes accumulated by the solver
    float px[3][NP];                       // centres
    float tdv[8];                          // the tool's six velocity changes while particle rows run
    float tinv[12];                        // tool: 1 / m, world inverse inertia (row-major 3 x 3 from [1])
    float tpos[4];                         // tool body frame origin (its COM)
    float tqd[8];                          // tool velocity after the unconstrained update
    uint32_t touch[4];                     // particles with a contact point on the human [0..1] / on the table or the bowl [2..3]
    float lam[2][NPC];                     // accumulated normal / friction impulses, indexed like the sorted records
    uint16_t order[NPC];                   // contact indices sorted by round
    uint16_t round_of[NPC];
    uint16_t rstart[NPC + 2];              // first entry of each round in `order`
    unsigned long long used[64];           // greedy colouring: rounds already taken by a row of particle p
    uint8_t fill[NPC + 2];
};
using SmPartSmall = SmPartT<8, 64>;
using SmPartLarge = SmPartT<64, AVG_MAX_PCONTACT>;

namespace {
__device__ __forceinline__ V3 plane_space1(V3 n) {       // btPlaneSpace1
    if (fabsf(n.z) > 0.70710678f) { const float k = rsqrtf(n.y * n.y + n.z * n.z); return mk3(0, -n.z * k, n.y * k); }
    const float k = rsqrtf(n.x * n.x + n.y * n.y); return mk3(-n.y * k, n.x * k, 0);
}
template <int NP>
__device__ __forceinline__ V3 sm3(const float (*a)[NP], int base, int p) { return mk3(a[base][p], a[base + 1][p], a[base + 2][p]); }
__device__ __forceinline__ V3 tinv_mul(const float* ti, V3 v) {
    return mk3(ti[1] * v.x + ti[2] * v.y + ti[3] * v.z, ti[4] * v.x + ti[5] * v.y + ti[6] * v.z, ti[7] * v.x + ti[8] * v.y + ti[9] * v.z);
}

// One contact record: row data of the normal and the friction row of a particle contact (see the header comment).
// [0..2] n  [3] target_n  [4] 1/diag_n  [5] lambda_n  [6..8] t  [9] target_t  [10] 1/diag_t  [11] lambda_t  [12] mu
// [13] p | q << 8 | kind << 16  [14..16] lever arm on the tool  [17] diag_n  [18] diag_t
template <class SP>
__device__ void particle_record(const KM& m, const SP& sp, const float* scr, float* rec, V3 n, float dist, int p, int q, int shape_b,
                                float dt, uint32_t* touch) {
    const AvgModelHeader* h = m.h;
    const AvgShape* PS = &m.shape[h->pshape];
    const float r = PS->radius, inv_m = 1.0f / h->p_mass, inv_i = 1.0f / (0.4f * h->p_mass * r * r);
    const V3 xa = sm3(sp.px, 0, p), va = sm3(sp.pv, 0, p), wa = sm3(sp.pv, 3, p);
    const V3 ra = n * (-r);
    V3 vrel = va + cross(wa, ra);
    int kind = 0;
    float mu_b = PS->friction;
    V3 rb = mk3(0, 0, 0);
    float dn = inv_m, dtg = inv_m + r * r * inv_i;                       // (ra x t)^2 = r^2 for a unit tangent
    if (q >= 0) {
        kind = 1;
        vrel = vrel - (sm3(sp.pv, 0, q) + cross(sm3(sp.pv, 3, q), n * r));
        dn += inv_m; dtg += inv_m + r * r * inv_i;
    } else {
        const AvgShape* SB = &m.shape[shape_b];
        mu_b = SB->friction;
        const int body = SB->body;
        const V3 pb = xa - n * (r + dist);
        if (body >= 0 && body == h->tool_body) {
            kind = 2;
            rb = pb - mk3(sp.tpos[0], sp.tpos[1], sp.tpos[2]);
            vrel = vrel - (mk3(sp.tqd[0], sp.tqd[1], sp.tqd[2]) + cross(mk3(sp.tqd[3], sp.tqd[4], sp.tqd[5]), rb));
        } else if (body >= 0 && body < h->n_body) {
            kind = 3;
            const float4* tw = reinterpret_cast<const float4*>(scr + AVG_S_TWIST) + 2 * body;
            const float4 wa4 = tw[0], vl4 = tw[1];
            const V3 ref = mk3(h->task_f[16], h->task_f[17], h->task_f[18]);
            vrel = vrel - (mk3(vl4.x, vl4.y, vl4.z) + cross(mk3(wa4.x, wa4.y, wa4.z), pb - ref));
        }
        const int rbd = SB->ref_body;
        if (rbd == AVG_REF_HUMAN) atomicOr(&touch[p >> 5], 1u << (p & 31));
        if (rbd == AVG_REF_TABLE || rbd == AVG_REF_BOWL) atomicOr(&touch[2 + (p >> 5)], 1u << (p & 31));
    }
    const float vn = dot(vrel, n);
    const V3 lat = vrel - n * vn;
    const float ll = norm(lat);
    const V3 t = ll > 1e-6f ? lat * (1.0f / ll) : plane_space1(n);
    if (kind == 2) {
        const V3 jn = cross(rb, n), jt = cross(rb, t);
        dn += sp.tinv[0] + dot(jn, tinv_mul(sp.tinv, jn));
        dtg += sp.tinv[0] + dot(jt, tinv_mul(sp.tinv, jt));
    }
    rec[0] = n.x; rec[1] = n.y; rec[2] = n.z;
    rec[3] = (dist > 0.0f ? -dist / dt : -dist * h->erp / dt) - vn;
    rec[4] = dn > 1e-12f ? 1.0f / dn : 0.0f; rec[5] = 0.0f;
    rec[6] = t.x; rec[7] = t.y; rec[8] = t.z;
    rec[9] = -dot(vrel, t);
    rec[10] = dtg > 1e-12f ? 1.0f / dtg : 0.0f; rec[11] = 0.0f;
    rec[12] = PS->friction * mu_b;
    rec[13] = __int_as_float(p | ((q >= 0 ? q : 0xff) << 8) | (kind << 16));
    rec[14] = rb.x; rec[15] = rb.y; rec[16] = rb.z; rec[17] = dn; rec[18] = dtg; rec[19] = 0.0f;
}

// Prologue of the particle part of one internal step: unconstrained particle velocities, contact records, touch masks and
// the round schedule.  Returns the number of contacts; nrounds through the reference.
template <class SP>
__device__ int particles_prepare(const KM& m, SP& sp, const AvgStepArgs& a, int e, const float* scr, int lane, float dt, int& nrounds, int& overflow) {
    const AvgModelHeader* h = m.h;
    const int np = h->n_particle;
    float* ps = a.pscratch + (size_t)e * AVG_PS_STRIDE;
    const int* ps_i = reinterpret_cast<const int*>(ps);
    const float* prec = a.part + (size_t)e * AVG_P_STRIDE;
    const uint32_t alive0 = __float_as_uint(prec[AVG_P_ALIVE]), alive1 = __float_as_uint(prec[AVG_P_ALIVE + 1]);
    const V3 g = mk3(h->p_gravity[0], h->p_gravity[1], h->p_gravity[2]);
    for (int p = lane; p < SP::kNP; p += 32) {
        const bool live = p < np && (((p < 32 ? alive0 : alive1) >> (p & 31)) & 1u);
        V3 x = mk3(0, 0, 0), v = x, w = x;
        if (live) {
            x = mk3(prec[AVG_P_POS + p], prec[AVG_P_POS + 64 + p], prec[AVG_P_POS + 128 + p]);
            v = mk3(prec[AVG_P_VEL + p], prec[AVG_P_VEL + 64 + p], prec[AVG_P_VEL + 128 + p]);
            w = mk3(prec[AVG_P_ANG + p], prec[AVG_P_ANG + 64 + p], prec[AVG_P_ANG + 128 + p]);
            // gravity and the multibody base damping (0.04, as for every other body of the step)
            v = v + (g - v * (h->lin_damp + h->lin_damp * norm(v))) * dt;
            w = w - w * ((h->ang_damp + h->ang_damp * norm(w)) * dt);
        }
        sp.px[0][p] = x.x; sp.px[1][p] = x.y; sp.px[2][p] = x.z;
        sp.pv[0][p] = v.x; sp.pv[1][p] = v.y; sp.pv[2][p] = v.z; sp.pv[3][p] = w.x; sp.pv[4][p] = w.y; sp.pv[5][p] = w.z;
#pragma unroll
        for (int k = 0; k < 6; ++k) sp.pdv[k][p] = 0.0f;
    }
    const int tb = h->tool_body;
    if (tb >= 0) {
        const int td = m.body[tb].dof, fbi = (td - h->n_jdof) / 6;
        if (lane < 12) sp.tinv[lane] = scr[AVG_S_FREEINV + 12 * fbi + lane];
        if (lane < 6) sp.tqd[lane] = scr[AVG_S_QD + td + lane];
        if (lane < 3) sp.tpos[lane] = scr[AVG_S_POSE + 8 * tb + lane];
    }
    if (lane < 4) sp.touch[lane] = 0u;
    if (lane < 6) sp.tdv[lane] = 0.0f;
    __syncwarp();
    // contact records in canonical order: hits among the particle-vs-shape candidates (particle-major, shape index ascending),
    // then the particle-particle contacts
    float* rec = ps + AVG_PS_REC;
    const int ncand = min(ps_i[AVG_PS_NCAND], AVG_MAX_PCAND), npp = min(ps_i[AVG_PS_NPP], AVG_PS_MAXPP);
    int nc = 0;
    const float4* cand = reinterpret_cast<const float4*>(ps + AVG_PS_CAND);
#pragma unroll 1
    for (int base = 0; base < ncand; base += 32) {
        const int i = base + lane;
        float4 r0 = make_float4(0, 0, 0, 0), r1 = r0;
        bool hit = false;
        if (i < ncand) { r1 = cand[2 * i + 1]; hit = r1.z != 0.0f; if (hit) r0 = cand[2 * i]; }
        const unsigned bal = __ballot_sync(AVG_FULL, hit);
        const int slot = nc + __popc(bal & ((1u << lane) - 1));
        if (hit && slot < SP::kNPC)
            particle_record(m, sp, scr, rec + AVG_PS_REC_STRIDE * slot, mk3(r0.x, r0.y, r0.z), r0.w, __float_as_int(r1.x), -1, __float_as_int(r1.y), dt, sp.touch);
        nc += __popc(bal);
    }
    const float4* pp = reinterpret_cast<const float4*>(ps + AVG_PS_PP);
#pragma unroll 1
    for (int base = 0; base < npp; base += 32) {
        const int i = base + lane, slot = nc + i;
        if (i < npp && slot < SP::kNPC) {
            const float4 r0 = pp[2 * i], r1 = pp[2 * i + 1];
            particle_record(m, sp, scr, rec + AVG_PS_REC_STRIDE * slot, mk3(r0.x, r0.y, r0.z), r0.w, __float_as_int(r1.x), __float_as_int(r1.y), -1, dt, sp.touch);
        }
    }
    nc += npp;
    if (nc > SP::kNPC) { overflow |= 8; nc = SP::kNPC; }
    if (ps_i[AVG_PS_OVERFLOW]) overflow |= ps_i[AVG_PS_OVERFLOW];
    __syncwarp();
    // round schedule = greedy colouring of the contacts in canonical order (serial: ~15 instructions per contact, once per
    // internal step): a contact takes the lowest round not yet used by a contact of its particle(s), at most 32 per round.
    // Rows of one round share no particle, so they commute; the oracle sweeps in the same order (particle_row_order).
    for (int i = lane; i < 64; i += 32) sp.used[i] = 0ull;
    for (int i = lane; i < nc + 2; i += 32) sp.fill[i] = 0;
    __syncwarp();
    int nr = 0;
    if (lane == 0) {
        unsigned long long full = 0ull;
        for (int c = 0; c < nc; ++c) {
            const int pk = __float_as_int(rec[AVG_PS_REC_STRIDE * c + 13]);
            const int p = pk & 0xff, q = (pk >> 8) & 0xff, kind = pk >> 16;
            const unsigned long long mk = sp.used[p] | (kind == 1 ? sp.used[q] : 0ull) | full;
            int r = 63;
            if (mk != ~0ull) r = __ffsll((long long)~mk) - 1; else overflow |= 8;
            sp.round_of[c] = (uint16_t)r;
            sp.used[p] |= 1ull << r;
            if (kind == 1) sp.used[q] |= 1ull << r;
            if (++sp.fill[r] >= 32) full |= 1ull << r;
            nr = max(nr, r + 1);
        }
        int acc = 0;
        for (int r = 0; r < nr; ++r) { sp.rstart[r] = (uint16_t)acc; acc += sp.fill[r]; sp.fill[r] = 0; }
        sp.rstart[nr] = (uint16_t)acc;
        for (int c = 0; c < nc; ++c) { const int r = sp.round_of[c]; sp.order[sp.rstart[r] + sp.fill[r]] = (uint16_t)c; sp.fill[r]++; }
    }
    nrounds = __shfl_sync(AVG_FULL, nr, 0);
    __syncwarp();
    // the records in round order, so that the rows of a round are adjacent (one vector load per lane and float4)
    {
        const float4* src = reinterpret_cast<const float4*>(rec);
        float4* dst = reinterpret_cast<float4*>(ps + AVG_PS_SORTED);
        for (int i = lane; i < 5 * nc; i += 32) { const int row = i / 5, k = i - 5 * row; dst[i] = src[5 * sp.order[row] + k]; }
        for (int i = lane; i < nc; i += 32) { sp.lam[0][i] = 0.0f; sp.lam[1][i] = 0.0f; }
    }
    __syncwarp();
    return nc;
}

// One sweep over the particle rows (normal rows, or friction rows), round by round; `srec` = records in round order.  The
// records of the next round are fetched while the current one is computed (they do not depend on it; only the impulses and
// velocities in shared memory do).  Returns the largest |delta| * diag.
template <bool FRICTION, class SP>
__device__ __forceinline__ float particles_sweep(const KM& m, SP& sp, const float* srec, int nrounds, int lane) {
    const AvgModelHeader* h = m.h;
    const float r = m.shape[h->pshape].radius, inv_m = 1.0f / h->p_mass, inv_i = 1.0f / (0.4f * h->p_mass * r * r);
    const float4* R4 = reinterpret_cast<const float4*>(srec);
    float resid = 0.0f;
    // the tool's velocity change as the block began (read by every row of the block) and this lane's share of the block's
    // reactions on the tool (summed over the warp after the last round: a fixed order, so the result is bit-deterministic)
    const V3 tdl = mk3(sp.tdv[0], sp.tdv[1], sp.tdv[2]), tda = mk3(sp.tdv[3], sp.tdv[4], sp.tdv[5]);
    V3 accl = mk3(0, 0, 0), acca = mk3(0, 0, 0);
    float4 n0 = make_float4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0, n4 = n0;
    {
        const int cnt = nrounds > 0 ? sp.rstart[1] - sp.rstart[0] : 0;
        if (lane < cnt) { const float4* q = R4 + 5 * lane; n0 = q[0]; n1 = q[1]; if (FRICTION) n2 = q[2]; n3 = q[3]; n4 = q[4]; }
    }
#pragma unroll 1
    for (int rd = 0; rd < nrounds; ++rd) {
        const int s0 = sp.rstart[rd], s1 = sp.rstart[rd + 1], cnt = s1 - s0;
        const float4 c0 = n0, c1 = n1, c2 = n2, c3 = n3, c4 = n4;
        if (rd + 1 < nrounds) {
            const int cn = sp.rstart[rd + 2] - s1;
            if (lane < cn) { const float4* q = R4 + 5 * (s1 + lane); n0 = q[0]; n1 = q[1]; if (FRICTION) n2 = q[2]; n3 = q[3]; n4 = q[4]; }
        }
        if (lane < cnt) {
            const int idx = s0 + lane;
            // c0 = n, target_n | c1 = 1/diag_n, lambda_n (unused), t.x, t.y | c2 = t.z, target_t, 1/diag_t, lambda_t (unused)
            // c3 = mu, packed, rb.x, rb.y | c4 = rb.z, diag_n, diag_t, pad
            const int pk = __float_as_int(c3.y);
            const int p = pk & 0xff, q = (pk >> 8) & 0xff, kind = pk >> 16;
            const V3 n = mk3(c0.x, c0.y, c0.z);
            const V3 d = FRICTION ? mk3(c1.z, c1.w, c2.x) : n;
            const V3 ja = FRICTION ? cross(n, d) * (-r) : mk3(0, 0, 0);        // A's angular Jacobian: lever -r n (zero for the normal row)
            float jdv = dot(d, sm3(sp.pdv, 0, p));
            if (FRICTION) jdv += dot(ja, sm3(sp.pdv, 3, p));
            V3 jb = mk3(0, 0, 0);
            if (kind == 1) {
                jdv -= dot(d, sm3(sp.pdv, 0, q));
                if (FRICTION) { jb = cross(n, d) * r; jdv -= dot(jb, sm3(sp.pdv, 3, q)); }
            } else if (kind == 2) {
                jb = cross(mk3(c3.z, c3.w, c4.x), d);
                jdv -= dot(d, tdl) + dot(jb, tda);
            }
            const float lam = sp.lam[FRICTION ? 1 : 0][idx];
            float sum = fmaf((FRICTION ? c2.y : c0.w) - jdv, FRICTION ? c2.z : c1.x, lam);
            if (FRICTION) { const float lim = c3.x * sp.lam[0][idx]; sum = fminf(fmaxf(sum, -lim), lim); }
            else sum = fmaxf(sum, 0.0f);
            const float delta = sum - lam;
            sp.lam[FRICTION ? 1 : 0][idx] = sum;
            const float dm = delta * inv_m;
            sp.pdv[0][p] += d.x * dm; sp.pdv[1][p] += d.y * dm; sp.pdv[2][p] += d.z * dm;
            if (FRICTION) { const float di = delta * inv_i; sp.pdv[3][p] += ja.x * di; sp.pdv[4][p] += ja.y * di; sp.pdv[5][p] += ja.z * di; }
            if (kind == 1) {
                sp.pdv[0][q] -= d.x * dm; sp.pdv[1][q] -= d.y * dm; sp.pdv[2][q] -= d.z * dm;
                if (FRICTION) { const float di = delta * inv_i; sp.pdv[3][q] -= jb.x * di; sp.pdv[4][q] -= jb.y * di; sp.pdv[5][q] -= jb.z * di; }
            } else if (kind == 2) {
                accl = accl - d * delta; acca = acca - jb * delta;          // impulse and angular impulse on the tool
            }
            resid = fmaxf(resid, fabsf(delta) * (FRICTION ? c4.z : c4.y));
        }
        __syncwarp();
    }
    // the block's reactions reach the tool at once: dv_tool += M_tool^-1 (sum of the impulses)
    if (__any_sync(AVG_FULL, accl.x != 0.0f || accl.y != 0.0f || accl.z != 0.0f || acca.x != 0.0f || acca.y != 0.0f || acca.z != 0.0f)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            accl.x += __shfl_xor_sync(AVG_FULL, accl.x, o); accl.y += __shfl_xor_sync(AVG_FULL, accl.y, o); accl.z += __shfl_xor_sync(AVG_FULL, accl.z, o);
            acca.x += __shfl_xor_sync(AVG_FULL, acca.x, o); acca.y += __shfl_xor_sync(AVG_FULL, acca.y, o); acca.z += __shfl_xor_sync(AVG_FULL, acca.z, o);
        }
        if (lane == 0) {
            const V3 wj = tinv_mul(sp.tinv, acca);
            sp.tdv[0] += accl.x * sp.tinv[0]; sp.tdv[1] += accl.y * sp.tinv[0]; sp.tdv[2] += accl.z * sp.tinv[0];
            sp.tdv[3] += wj.x; sp.tdv[4] += wj.y; sp.tdv[5] += wj.z;
        }
    }
    __syncwarp();
    return resid;
}

// End of the internal step: integrate the live particles and publish the touch masks (feeding.py:111,116 read them).
template <class SP>
__device__ void particles_finish(const KM& m, const SP& sp, const AvgStepArgs& a, int e, int lane, float dt, int nc, int overflow) {
    const AvgModelHeader* h = m.h;
    const int np = h->n_particle;
    float* prec = a.part + (size_t)e * AVG_P_STRIDE;
    const uint32_t alive0 = __float_as_uint(prec[AVG_P_ALIVE]), alive1 = __float_as_uint(prec[AVG_P_ALIVE + 1]);
    for (int p = lane; p < np; p += 32) {
        if (!(((p < 32 ? alive0 : alive1) >> (p & 31)) & 1u)) continue;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = sp.pv[k][p] + sp.pdv[k][p];
            prec[AVG_P_VEL + 64 * k + p] = v;
            prec[AVG_P_ANG + 64 * k + p] = sp.pv[3 + k][p] + sp.pdv[3 + k][p];
            prec[AVG_P_POS + 64 * k + p] = sp.px[k][p] + dt * v;
        }
    }
    if (lane < 2) prec[AVG_P_TOUCH_HUMAN + lane] = __uint_as_float(sp.touch[lane]);
    else if (lane < 4) prec[AVG_P_TOUCH_SPILL + lane - 2] = __uint_as_float(sp.touch[lane]);
    if (lane == 0) {
        reinterpret_cast<int*>(prec)[AVG_P_NCONTACT] = nc;
        if (overflow) reinterpret_cast<int*>(prec)[AVG_P_NCONTACT + 1] |= overflow;
    }
}
}  // namespace

// =================================================================================================================
// projected Gauss-Seidel + integration + human hard limits
// =================================================================================================================
// PART: 0 no particles, 1 Feeding (8 particles), 2 Drinking (64 particles)
// Body of the solve kernel for ONE environment (one warp).  PF_AHEAD: distance of the L2 prefetch hints (0: none -- the fused
// dynamics + solve kernel reads rows its own warp has just written); part_smem: the particle block behind SmSolve (PART > 0).
template <int MAXBLK, int PART, int PF_AHEAD>
__device__ __forceinline__ void solve_body(const AvgStepArgs& a, SmSolve& s, unsigned char* part_smem, const int e, const int lane, const KM& m, float* grec, float* scr) {
    const AvgModelHeader* h = m.h;
    AVG_MASK_CHECK
    const int nb = h->n_body, nj = h->n_jdof, nd = h->n_dof;
    const float dt = h->dt;
    int* scr_i = reinterpret_cast<int*>(scr);
    const int ndense = scr_i[AVG_S_NR], nlim = scr_i[AVG_S_NS], nfr = scr_i[AVG_S_NFR], first_contact_row = scr_i[AVG_S_FCR];
    const int nc = scr_i[AVG_S_NCS];
    const float* gJ = scr + AVG_S_J; const float* gW = scr + AVG_S_W;
    {   // L2 prefetch for the environment whose warp takes this slot next: counters + velocities (1 line), motor / limit rows (4),
        // weld rows (2), M^-1 columns (10), J and W of the weld rows (6 + 6)
        const int ea = e + PF_AHEAD;
        if (PF_AHEAD > 0 && ea < a.env_end) {
            const float* s2 = a.scratch + (size_t)ea * AVG_S_STRIDE;
            const float* p = lane < 1 ? s2 : (lane < 5 ? s2 + AVG_S_ROWS_M + 32 * (lane - 1) : (lane < 7 ? s2 + AVG_S_ROWS_D + 32 * (lane - 5)
                           : (lane < 17 ? s2 + AVG_S_MINV + 32 * (lane - 7) : (lane < 23 ? s2 + AVG_S_J + 32 * (lane - 17) : (lane < 29 ? s2 + AVG_S_W + 32 * (lane - 23) : nullptr)))));
            if (p) prefetch_l2(p);
        }
    }
    // stage rows: coalesced loads from the arena
    for (int d = 0; d < min(ndense, kSmDense); ++d) { if (d >= 6) s.J[d][lane] = gJ[d * 32 + lane]; s.W[d][lane] = gW[d * 32 + lane]; }
    {
        const float4* g_rd = reinterpret_cast<const float4*>(scr + AVG_S_ROWS_D);
        for (int i = lane; i < 2 * ndense; i += 32) s.rd[i >> 1][i & 1] = g_rd[i];
        __syncwarp();
        if (lane < 6) s.rd[lane][0].z = s.rd[lane][1].x;             // weld rows: lo = -hi, so the slot carries the row's diagonal instead (one load per row and sweep)
    }
    const float qd = lane < nd ? scr[AVG_S_QD + lane] : 0.0f;
    // block of this lane's dof, and this lane's column of the block of M^-1 (registers).  Unit rows are staged per
    // block (group 3 = lanes without a joint dof: all-zero rows), so the sweep addresses them with immediates.
    int bs = lane, grp = 3;
    for (int b = 0; b < h->n_block; ++b) {
        const int b0 = h->block_start[b], b1 = h->block_start[b + 1];
        if (lane >= b0 && lane < b1) { bs = b0; grp = b; }
    }
    {
        const float4 z4 = make_float4(0, 0, 0, 0);
        float4* um0 = &s.um[0][0]; float4* ul0 = &s.ul[0][0];
        for (int i = lane; i < 4 * kMaxBlk; i += 32) { um0[i] = z4; ul0[i] = z4; }
        __syncwarp();
        if (lane < nj) {
            s.um[grp][lane - bs] = reinterpret_cast<const float4*>(scr + AVG_S_ROWS_M)[lane];
            if (nlim) s.ul[grp][lane - bs] = reinterpret_cast<const float4*>(scr + AVG_S_ROWS_L)[lane];
        }
    }
    const float4* umg = s.um[grp]; const float4* ulg = s.ul[grp];
    const int lt = lane - bs;
    // block slots that hold a limit row in ANY articulation: the sweep skips the others (a uniform branch per slot)
    unsigned limslots = 0;
#if AVG_LIM_SKIP
    if (nlim) limslots = __reduce_or_sync(AVG_FULL, (lane < nj && s.ul[grp][lt].y != 0.0f) ? 1u << lt : 0u);
#else
    limslots = 0xffffffffu;
#endif
    float mcol[MAXBLK];
#pragma unroll
    for (int t = 0; t < MAXBLK; ++t) mcol[t] = scr[AVG_S_MINV + t * 32 + lane];
    __syncwarp();

    // ---- projected Gauss-Seidel.  dv lives in one register per lane.  Unit rows: the articulation blocks are swept
    //      concurrently (they do not couple), in Bullet's order inside a block.  Every lane of a block evaluates the
    //      block's current row (same inputs, same result), so motor impulses are replicated in registers and need
    //      neither an owner lane nor a shuffle; the lane's column of M^-1 sits in registers.  Slots past the end of a
    //      block hold all-zero rows (1/diag = 0), which makes their delta exactly 0.
    //      Dense rows follow in strict order: the six weld rows unrolled with J, W and impulses in registers, then
    //      the contact rows (rare) from shared memory / the arena with the impulse of row d in lane d.
    float dv = 0.0f, lamL = 0.0f;
    for (int d = 6 + lane; d < ndense; d += 32) s.lam[d] = 0.0f;                 // impulses of the contact rows live in shared memory
    __syncwarp();
    // particles (Feeding / Drinking): records, schedule and unconstrained velocities for this internal step
    using SP = typename std::conditional<PART == 2, SmPartLarge, SmPartSmall>::type;
    SP* spp = PART ? reinterpret_cast<SP*>(part_smem) : nullptr;
    int npc = 0, nrounds = 0, p_overflow = 0, tool_dof = -1;
    float* prec_rows = nullptr;
    if (PART) {
        npc = particles_prepare(m, *spp, a, e, scr, lane, dt, nrounds, p_overflow);
        prec_rows = a.pscratch + (size_t)e * AVG_PS_STRIDE + AVG_PS_SORTED;
        tool_dof = h->tool_body >= 0 ? m.body[h->tool_body].dof : -1;
    }
    // warm-started contact normal rows: the row starts from its impulse and the velocities from that impulse's effect
    if (h->warmstart > 0.0f) {
#pragma unroll 1
        for (int c = 0; c < nc; ++c) {
            const float lam0 = scr[AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * c + 13];
            if (lam0 != 0.0f) {
                const int d = first_contact_row + c;
                s.lam[d] = lam0;
                dv = fmaf(d < kSmDense ? s.W[d][lane] : gW[d * 32 + lane], lam0, dv);
            }
        }
        __syncwarp();
    }
    float lamM[MAXBLK], lamW[6], jw[6];
#pragma unroll
    for (int t = 0; t < MAXBLK; ++t) lamM[t] = 0.0f;
#pragma unroll
    for (int d = 0; d < 6; ++d) { lamW[d] = 0.0f; jw[d] = gJ[d * 32 + lane]; }
    const float thr = sqrtf(h->residual_thr);                        // on |delta impulse| * diag (Bullet squares both sides)
    int iters = 0;
    for (int it = 0; it < h->solver_iters; ++it) {
        float resid = 0.0f;
#pragma unroll
        for (int t = 0; t < MAXBLK; ++t) {
            const float4 ra = umg[t];
            const float jdv = __shfl_sync(AVG_FULL, dv, bs + t);
            const float sum = fminf(fmaxf(fmaf(ra.x - jdv, ra.y, lamM[t]), -ra.z), ra.z);
            const float delta = sum - lamM[t];
            lamM[t] = sum;
            dv = fmaf(mcol[t], delta, dv);
            resid = fmaxf(resid, fabsf(delta) * ra.w);
        }
        if (nlim) {
#pragma unroll
            for (int t = 0; t < MAXBLK; ++t) {
                if (!((limslots >> t) & 1u)) continue;
                const float4 ra = ulg[t];
                const float sg = ra.w;
                const float jdv = sg * __shfl_sync(AVG_FULL, dv, bs + t);
                const float lam = __shfl_sync(AVG_FULL, lamL, bs + t);
                const float sum = fminf(fmaxf(fmaf(ra.x - jdv, ra.y, lam), 0.0f), 100.0f);
                const float delta = sum - lam;
                dv = fmaf(sg * mcol[t], delta, dv);
                if (lt == t) lamL = sum;
                resid = fmaxf(resid, fabsf(delta) * ra.z);
            }
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            const float4 ra = s.rd[d][0];                            // {target, 1/diag, diag, max}: the weld clamp is symmetric (+-max)
            const float jdv = dense_dot(jw[d], dv);
            const float sum = fminf(fmaxf(fmaf(ra.x - jdv, ra.y, lamW[d]), -ra.w), ra.w);
            const float delta = sum - lamW[d];
            lamW[d] = sum;
            dv = fmaf(s.W[d][lane], delta, dv);
            resid = fmaxf(resid, fabsf(delta) * ra.z);
        }
        // every lane holds the same jdv (an exact integer sum) and the same row data, hence the same impulse: all lanes store it
        // to the same word -- no owner lane, no shuffle.  A row reads its words before its warp-wide reduction (redux.sync is a
        // convergence point) and writes after it, so a write never overtakes another lane's read of the old value.
        auto lam_get = [&](int d) { return s.lam[d]; };
#if AVG_REDUX == 0
        auto lam_set = [&](int d, float v) { s.lam[d] = __shfl_sync(AVG_FULL, v, 0); };    // float butterflies may differ in the last bit between lanes
#else
        auto lam_set = [&](int d, float v) { s.lam[d] = v; };
#endif
#pragma unroll 1
        for (int d = 6; d < nfr; ++d) {
            const float4 ra = s.rd[d][0]; const float4 rb = s.rd[d][1];
            const float* jp = d < kSmDense ? &s.J[d][lane] : &gJ[d * 32 + lane];
            const float* wp = d < kSmDense ? &s.W[d][lane] : &gW[d * 32 + lane];
            const float lam = lam_get(d);                            // read BEFORE the warp-wide reduction, written after it: no lane can
            const float jdv = dense_dot(*jp, dv);                    // overwrite the word while another still has to read the old value
            const float sum = fminf(fmaxf(fmaf(ra.x - jdv, ra.y, lam), ra.z), ra.w);
            const float delta = sum - lam;
            lam_set(d, sum);
            dv = fmaf(*wp, delta, dv);
            resid = fmaxf(resid, fabsf(delta) * rb.x);
        }
        if (PART && npc > 0) {                                       // particle contact normals (after the articulation's, Bullet's order)
            if (tool_dof >= 0 && lane >= tool_dof && lane < tool_dof + 6) spp->tdv[lane - tool_dof] = dv;
            __syncwarp();
            resid = fmaxf(resid, particles_sweep<false, SP>(m, *spp, prec_rows, nrounds, lane));
            if (tool_dof >= 0 && lane >= tool_dof && lane < tool_dof + 6) dv = spp->tdv[lane - tool_dof];
        }
#pragma unroll 1
        for (int d = nfr; d < ndense; ++d) {
            const float4 ra = s.rd[d][0]; const float4 rb = s.rd[d][1];
            const int par = __float_as_int(rb.w);
            const float* jp = d < kSmDense ? &s.J[d][lane] : &gJ[d * 32 + lane];
            const float* wp = d < kSmDense ? &s.W[d][lane] : &gW[d * 32 + lane];
            const float lam = lam_get(d);
            const float lim = rb.y * lam_get(par);
            const float jdv = dense_dot(*jp, dv);
            const float sum = fminf(fmaxf(fmaf(ra.x - jdv, ra.y, lam), -lim), lim);
            const float delta = sum - lam;
            lam_set(d, sum);
            dv = fmaf(*wp, delta, dv);
            resid = fmaxf(resid, fabsf(delta) * rb.x);
        }
        if (PART && npc > 0) {                                       // particle friction rows
            if (tool_dof >= 0 && lane >= tool_dof && lane < tool_dof + 6) spp->tdv[lane - tool_dof] = dv;
            __syncwarp();
            resid = fmaxf(resid, particles_sweep<true, SP>(m, *spp, prec_rows, nrounds, lane));
            if (tool_dof >= 0 && lane >= tool_dof && lane < tool_dof + 6) dv = spp->tdv[lane - tool_dof];
        }
        iters++;
        if (!__any_sync(AVG_FULL, resid > thr)) break;              // blocks ran in different lanes
    }

    // contact impulses (getContactPoints()[9] = impulse / dt), read by the epilogue after the last sub-step
    for (int c = lane; c < nc; c += 32) scr[AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * c + 12] = s.lam[first_contact_row + c];
    if (lane < AVG_WCACHE_N) {                                   // what the next internal step warm-starts from
        const bool have = lane < nc;
        const float* cc = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * lane;
        const uint32_t key = have ? ((uint32_t)__float_as_int(cc[10]) | ((uint32_t)__float_as_int(cc[11]) << 16)) : 0u;
        grec[AVG_E_WCACHE + 2 * lane] = __uint_as_float(key);
        grec[AVG_E_WCACHE + 2 * lane + 1] = have ? s.lam[first_contact_row + lane] : 0.0f;
    }
    if (lane == 0) scr_i[AVG_S_ITERS] += iters;
    if (PART) particles_finish(m, *spp, a, e, lane, dt, npc, p_overflow);

    // ---- integrate (semi-implicit Euler), enforce_realistic_human_joint_limits (env.py:353-371, human-active ids),
    //      enforce_hard_human_joint_limits (env.py:389-410) -----------------------------------------------------------
    float v = qd + dv;
    if (lane < nj) v = fminf(fmaxf(v, -h->max_vel), h->max_vel);
    const bool is_jlane = lane < nb && m.body[lane].jtype != AVG_JOINT_FREE;
    float qn = 0.0f;
    if (is_jlane) qn = grec[AVG_E_Q + m.body[lane].qidx] + dt * v;
    // the reference's per-frame hooks run after every p.stepSimulation call (env.py:343-349), i.e. after the LAST internal
    // step of a frame when numSubSteps > 1
    if (a.post && h->human_control && m.mlp && h->mlp_dof[0] >= 0) {
        float q4[4], x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q4[k] = __shfl_sync(AVG_FULL, qn, h->mlp_dof[k]);
        arm_limit_inputs(q4[0], q4[1], q4[2], q4[3], x);
        const float logit = arm_limit_logit_warp(m.mlp, x, lane);
        int* grec_i = reinterpret_cast<int*>(grec);
        if (logit > 0.0f) {
            if (lane < 4) grec[AVG_E_VALID_POSE + lane] = q4[0] * (lane == 0) + q4[1] * (lane == 1) + q4[2] * (lane == 2) + q4[3] * (lane == 3);
            if (lane == 0) grec_i[AVG_E_HAS_VALID] = 1;
        } else if (grec_i[AVG_E_HAS_VALID] != 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (lane == h->mlp_dof[k]) { qn = grec[AVG_E_VALID_POSE + k]; v = 0.0f; }
        }
    }
    if (is_jlane) {
        const AvgDof* D = &m.dof[lane];
        if (a.post && (D->flags & AVG_DOF_HARD_LIMIT)) {
            const float sc = grec[AVG_E_LIMIT_SCALE];
            const float lo = D->lower * sc, hi = D->upper * sc;
            if (qn < lo) { qn = lo; v = 0.0f; }
            else if (qn > hi) { qn = hi; v = 0.0f; }
        }
        grec[AVG_E_Q + m.body[lane].qidx] = qn;
    }
    if (lane < nd) grec[AVG_E_QD + lane] = v;
    // free bodies: the body's lane gathers its six velocity components from the dof lanes
    {
        int fdof = -1, fq = 0;
        if (lane < nb && m.body[lane].jtype == AVG_JOINT_FREE) { fdof = m.body[lane].dof; fq = m.body[lane].qidx; }
        const int src = fdof >= 0 ? fdof : 0;
        const float v0 = __shfl_sync(AVG_FULL, v, src), v1 = __shfl_sync(AVG_FULL, v, (src + 1) & 31), v2 = __shfl_sync(AVG_FULL, v, (src + 2) & 31);
        const float w0 = __shfl_sync(AVG_FULL, v, (src + 3) & 31), w1 = __shfl_sync(AVG_FULL, v, (src + 4) & 31), w2 = __shfl_sync(AVG_FULL, v, (src + 5) & 31);
        if (fdof >= 0) {
            float* q = grec + AVG_E_Q + fq;
            q[0] += dt * v0; q[1] += dt * v1; q[2] += dt * v2;
            V3 w = mk3(w0, w1, w2); float wn = norm(w);
            Q4 cur = ldq(q + 3);
            if (wn * dt > 1e-9f) cur = qmul(qaxis(w * (1.0f / wn), wn * dt), cur);
            cur = qnormalize(cur);
            q[3] = cur.x; q[4] = cur.y; q[5] = cur.z; q[6] = cur.w;
        }
    }
}

template <int MAXBLK, int PART>
__global__ void __launch_bounds__(32, PART ? 20 : AVG_OCC_SOLVE)
avg_solve_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmSolve)
    solve_body<MAXBLK, PART, 148 * AVG_OCC_SOLVE * AVG_PF_PCT / 100>(a, s, smem_raw + sizeof(SmSolve), e, lane, m, grec, scr);
}

// Dynamics + solve of one internal step in ONE kernel (no particles): the warp that built an environment's rows sweeps them,
// reading them back from the arena lines it has just written (L2 hits instead of an HBM round trip across two kernels), and
// the sub-step is one launch shorter.  Shared memory is the union of the two bodies' blocks.
#ifndef AVG_WPB_FUSED
#define AVG_WPB_FUSED 1
#endif
#ifndef AVG_OCC_FUSED
#define AVG_OCC_FUSED (16 / AVG_WPB_FUSED)      /* 126 registers, nothing spilled: the kernel only runs launches of <= AVG_FUSE_MAX environments (at most 14 warps per SM),
                                                  where one warp's latency is the step time -- 4096 envs 1.251 -> 1.180 ms per step, 1024 envs 0.987 -> 0.916 against 24 blocks / 80 registers */
#endif
union __align__(16) SmDynSolve { SmDyn d; SmSolve s; };
template <int MAXBLK>
__global__ void __launch_bounds__(32 * AVG_WPB_FUSED, AVG_OCC_FUSED)
avg_dynsolve_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmDynSolve)
    dynamics_body<MAXBLK, 148 * AVG_OCC_FUSED * AVG_WPB_FUSED * AVG_PF_PCT / 100>(a, s.d, e, lane, m, grec, scr);
    __syncwarp();
    solve_body<MAXBLK, 0, 0>(a, s.s, nullptr, e, lane, m, grec, scr);
}

// =================================================================================================================
// forces, reward, observation, info (scratch_itch.py:53-128)
// =================================================================================================================
namespace {
// ScratchItchEnv._get_obs, scratch_itch.py:104-128 (lane 0 fills s.obs)
template <class SM>
__device__ void fill_obs(const KM& m, SM& s, V3 tgt, float tool_force, float total_force_on_human, float tool_force_at_target) {
    const AvgModelHeader* h = m.h;
    const int nj = h->n_jdof;
    V3 torso, tool, sh, el, wr, chest; Q4 tq, dq;
    frame_cached(s, AVG_F_TORSO, torso, dq);
    frame_cached(s, AVG_F_TOOL_TIP, tool, tq);
    frame_cached(s, AVG_F_SHOULDER, sh, dq);
    frame_cached(s, AVG_F_ELBOW, el, dq);
    frame_cached(s, AVG_F_WRIST, wr, dq);
    frame_cached(s, AVG_F_CHEST, chest, dq);
    float* o = s.obs; int k = 0;
    V3 t;
    t = tool - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    o[k++] = tq.x; o[k++] = tq.y; o[k++] = tq.z; o[k++] = tq.w;
    t = tool - tgt; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    t = tgt - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    for (int i = 0; i < nj; ++i) if (m.dof[i].action >= 0 && m.dof[i].action < h->n_action_robot) o[k++] = s.env[AVG_E_Q + m.body[m.dof[i].body].qidx];
    t = sh - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    t = el - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    t = wr - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    o[k++] = tool_force;
    if (h->human_control) {
        t = tool - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        o[k++] = tq.x; o[k++] = tq.y; o[k++] = tq.z; o[k++] = tq.w;
        t = tool - tgt; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        t = tgt - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        const int hq0 = k;
        for (int i = 0; i < 10; ++i) o[k++] = 0.0f;
        for (int i = 0; i < nj; ++i) if (m.dof[i].human_slot >= 0) o[hq0 + m.dof[i].human_slot] = s.env[AVG_E_Q + m.body[m.dof[i].body].qidx];
        t = sh - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        t = el - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        t = wr - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        o[k++] = total_force_on_human; o[k++] = tool_force_at_target;
    }
}
}  // namespace

__device__ __forceinline__ void epilogue_body(const AvgStepArgs& a, SmEpi& s, const int e, const int lane, const KM& m, float* grec, float* scr) {
    const AvgModelHeader* h = m.h;
    for (int i = lane; i < AVG_ENV_STRIDE; i += 32) s.env[i] = grec[i];
    __syncwarp();
    int* env_i = reinterpret_cast<int*>(s.env);
    const int* scr_i = reinterpret_cast<const int*>(scr);
    const int na = h->n_action_robot + h->n_action_human;
    const float* act = a.actions + (size_t)e * na;
    float raw_sq;
    {
        const float av = lane < na ? act[lane] : 0.0f;
        raw_sq = warp_sum(av * av);                          // reward_action uses the raw action, scratch_itch.py:64
    }
    fk_warp(m, s, s.env + AVG_E_Q, lane, h->n_body, s.env + AVG_E_EBODY);
    frames_warp(m, s, lane);
    V3 tgt; { V3 lp; Q4 lq; frame_cached(s, env_i[AVG_E_LIMB_FRAME], lp, lq); tgt = lp + qrot(lq, ld3(s.env + AVG_E_TARGET_ON_ARM)); }
    const float dt = h->dt;
    const float* tf = h->task_f;
    const int ncontact = scr_i[AVG_S_NCS];
    // get_total_force, scratch_itch.py:84-102
    float total_force_on_human = 0, tool_force = 0, tool_force_at_target = 0;
    bool have_tcp = false; V3 tcp = mk3(0, 0, 0);
    for (int ci = 0; ci < ncontact; ++ci) {
        const float* c = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * ci;
        const AvgShape* sa = &m.shape[__float_as_int(c[10])]; const AvgShape* sb = &m.shape[__float_as_int(c[11])];
        const float force = c[12] / dt;
        const bool a_tool = sa->ref_body == AVG_REF_TOOL, b_tool = sb->ref_body == AVG_REF_TOOL;
        const bool a_hum = sa->ref_body == AVG_REF_HUMAN, b_hum = sb->ref_body == AVG_REF_HUMAN;
        const bool a_rob = sa->ref_body == AVG_REF_ROBOT, b_rob = sb->ref_body == AVG_REF_ROBOT;
        if (a_tool || b_tool) tool_force += force;
        if ((a_tool && b_hum) || (b_tool && a_hum)) {
            total_force_on_human += force;
            const int link_tool = a_tool ? sa->ref_link : sb->ref_link;
            const V3 pos_h = a_tool ? ld3(c + 3) : ld3(c);
            if ((link_tool == 0 || link_tool == 1) && norm(pos_h - tgt) < tf[AVG_TF_TARGET_RADIUS]) {
                tool_force_at_target += force; tcp = pos_h; have_tcp = true;
            }
        }
        if ((a_rob && b_hum) || (b_rob && a_hum)) total_force_on_human += force;
    }
    if (lane == 0) {
        fill_obs(m, s, tgt, tool_force, total_force_on_human, tool_force_at_target);
        V3 tool; Q4 tq; frame_cached(s, AVG_F_TOOL_TIP, tool, tq);
        const int tb = m.frame[AVG_F_TOOL_TIP].body;
        const float* tv = s.env + AVG_E_QD + m.body[tb].dof;
        const float ee_vel = norm(ld3(tv) + cross(ld3(tv + 3), tool - ld3(s.bp[tb])));       // scratch_itch.py:54
        // human_preferences (env.py:412-448) and reward (scratch_itch.py:62-72)
        const float pref = tf[AVG_TF_C_V] * (-ee_vel) + tf[AVG_TF_C_F] * (-(total_force_on_human - tool_force_at_target))
                         + tf[AVG_TF_C_HF] * (tool_force_at_target < tf[AVG_TF_FORCE_CAP] ? 0.0f : -tool_force_at_target);
        const float reward_distance = -norm(tgt - tool);
        const float reward_action = -raw_sq;
        float reward_force_scratch = 0.0f;
        float task_success = s.env[AVG_E_TASK_SUCCESS];
        const V3 prev = ld3(s.env + AVG_E_PREV_CONTACT);
        if (have_tcp && norm(tcp - prev) > tf[AVG_TF_SCRATCH_MOVE] && tool_force_at_target < tf[AVG_TF_FORCE_CAP]) {
            reward_force_scratch = tool_force_at_target;
            st3(grec + AVG_E_PREV_CONTACT, tcp);
            task_success += 1.0f;
            grec[AVG_E_TASK_SUCCESS] = task_success;
        }
        const float reward = tf[AVG_TF_DISTANCE_W] * reward_distance + tf[AVG_TF_ACTION_W] * reward_action
                           + tf[AVG_TF_TOOL_FORCE_W] * tool_force_at_target + tf[AVG_TF_SCRATCH_W] * reward_force_scratch + pref;
        int* grec_i = reinterpret_cast<int*>(grec);
        grec[AVG_E_EPISODE_RETURN] = s.env[AVG_E_EPISODE_RETURN] + reward;
        st3(grec + AVG_E_TARGET_POS, tgt);
        grec_i[AVG_E_ITERATION] = env_i[AVG_E_ITERATION] + 1;                               // env.py:351
        grec_i[AVG_E_OVERFLOW] = env_i[AVG_E_OVERFLOW] | scr_i[AVG_S_OVERFLOW];
        grec_i[AVG_E_SOLVER_ITERS] = scr_i[AVG_S_ITERS];
        grec_i[AVG_E_NCAND] = scr_i[AVG_S_NCAND];
        a.reward[e] = reward;
        const float success = task_success >= tf[AVG_TF_SUCCESS_THR] ? 1.0f : 0.0f;
        a.info[2 * e] = total_force_on_human;
        a.info[2 * e + 1] = success;
        // the env itself never terminates (scratch_itch.py:78); with a time limit set, `done` is gym's TimeLimit wrapper
        // (__init__.py:21) evaluated per environment on its own step counter, so staggered episodes need no host bookkeeping
        if (a.done) a.done[e] = (a.time_limit > 0 && env_i[AVG_E_ITERATION] + 1 >= a.time_limit) ? 1 : 0;
        if (a.terms) {
            float* tr = a.terms + 8 * (size_t)e;
            tr[0] = total_force_on_human; tr[1] = success; tr[2] = tool_force; tr[3] = tool_force_at_target;
            tr[4] = reward_distance; tr[5] = reward_action; tr[6] = reward_force_scratch; tr[7] = pref;
        }
    }
    __syncwarp();
    const int nobs = h->n_obs_robot + h->n_obs_human;
    for (int i = lane; i < nobs; i += 32) a.obs[(size_t)e * nobs + i] = s.obs[i];
    if (a.contacts) {
        AvgContact* co = a.contacts + (size_t)e * kMaxC;
        for (int ci = lane; ci < kMaxC; ci += 32) {
            AvgContact c;
            if (ci < ncontact) {
                const float* g = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * ci;
                c.shape_a = __float_as_int(g[10]); c.shape_b = __float_as_int(g[11]);
                for (int k = 0; k < 3; ++k) { c.pos_a[k] = g[k]; c.pos_b[k] = g[3 + k]; c.normal[k] = g[6 + k]; }
                c.dist = g[9]; c.force = g[12] / dt;
            } else { c.shape_a = -1; c.shape_b = -1; for (int k = 0; k < 3; ++k) { c.pos_a[k] = c.pos_b[k] = c.normal[k] = 0; } c.dist = 0; c.force = 0; }
            c.pad[0] = c.pad[1] = c.pad[2] = 0;
            co[ci] = c;
        }
        if (lane == 0) a.ncontacts[e] = ncontact;
    }
}
__global__ void __launch_bounds__(32 * AVG_K_WARPS_PER_BLOCK)
avg_epilogue_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmEpi)
    epilogue_body(a, s, e, lane, m, grec, scr);
}

// =================================================================================================================
// BedBathing epilogue (bed_bathing.py:53-153): get_total_force with the wiping targets, closest tool-human distance,
// human_preferences, reward, observation, info
// =================================================================================================================
namespace {
// bed_bathing.py:129-153
template <class SM>
__device__ void fill_obs_bb(const KM& m, SM& s, float tool_force, float total_force_on_human, float tool_force_on_human) {
    const AvgModelHeader* h = m.h;
    const int nj = h->n_jdof;
    V3 torso, tool, sh, el, wr; Q4 tq, dq;
    frame_cached(s, AVG_F_TORSO, torso, dq);
    frame_cached(s, AVG_F_TOOL_TIP, tool, tq);
    frame_cached(s, AVG_F_SHOULDER, sh, dq);
    frame_cached(s, AVG_F_ELBOW, el, dq);
    frame_cached(s, AVG_F_WRIST, wr, dq);
    float* o = s.obs; int k = 0;
    V3 t;
    t = tool - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    o[k++] = tq.x; o[k++] = tq.y; o[k++] = tq.z; o[k++] = tq.w;
    for (int i = 0; i < nj; ++i) if (m.dof[i].action >= 0 && m.dof[i].action < h->n_action_robot) o[k++] = s.env[AVG_E_Q + m.body[m.dof[i].body].qidx];
    t = sh - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    t = el - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    t = wr - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    o[k++] = tool_force;
    if (h->human_control) {                                  // :136-139,149: positions relative to human link 3
        V3 chest; frame_cached(s, AVG_F_CHEST, chest, dq);
        t = tool - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        o[k++] = tq.x; o[k++] = tq.y; o[k++] = tq.z; o[k++] = tq.w;
        const int hq0 = k;
        for (int i = 0; i < 10; ++i) o[k++] = 0.0f;
        for (int i = 0; i < nj; ++i) if (m.dof[i].human_slot >= 0) o[hq0 + m.dof[i].human_slot] = s.env[AVG_E_Q + m.body[m.dof[i].body].qidx];
        t = sh - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        t = el - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        t = wr - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        o[k++] = total_force_on_human; o[k++] = tool_force_on_human;
    }
}

// world pose of any shape from the body poses of an epilogue's forward kinematics
template <class SM>
__device__ __forceinline__ void epi_load_shape(const KM& m, const SM& s, int si, WShape& w) {
    const AvgShape* S = &m.shape[si];
    w.s = S; w.verts = m.vert + 4 * S->vert_off; w.planes = m.plane + 4 * S->plane_off;
    Q4 q;
    if (S->body >= 0) {
        const Q4 bq = ldq(s.bq[S->body]);
        w.p = ld3(s.bp[S->body]) + qrot(bq, ld3(S->pos));
        q = qnormalize(qmul(bq, ldq(S->quat)));
    } else { w.p = ld3(S->pos); q = ldq(S->quat); }
    const M3 r = qmat(q);
#pragma unroll
    for (int i = 0; i < 9; ++i) w.R[i] = r.m[i];
}
// bounding capsule of a shape in the world frame (the blob stores it in the shape frame for moving shapes)
template <class SM>
__device__ __forceinline__ void epi_world_capsule(const KM& m, const SM& s, int si, float4& c0, float4& c1) {
    c0 = __ldg(&m.bcap[2 * si]); c1 = __ldg(&m.bcap[2 * si + 1]);
    const AvgShape* S = &m.shape[si];
    if (S->body >= 0) {
        const Q4 bq = ldq(s.bq[S->body]);
        const V3 sp = ld3(s.bp[S->body]) + qrot(bq, ld3(S->pos));
        const Q4 sq = qnormalize(qmul(bq, ldq(S->quat)));
        const V3 p0 = sp + qrot(sq, mk3(c0.x, c0.y, c0.z)), p1 = sp + qrot(sq, mk3(c1.x, c1.y, c1.z));
        c0 = make_float4(p0.x, p0.y, p0.z, c0.w); c1 = make_float4(p1.x, p1.y, p1.z, 0.0f);
    }
}
// closest distance between segments [p1, p1 + d1] and [p2, p2 + d2] (Ericson 5.1.9)
__device__ __forceinline__ float seg_seg_distance(V3 p1, V3 d1, V3 p2, V3 d2) {
    const V3 r = p1 - p2;
    const float aa = dot(d1, d1), ee = dot(d2, d2), ff = dot(d2, r);
    float sc = 0.0f, tc = 0.0f;
    if (aa <= 1e-12f && ee <= 1e-12f) { }
    else if (aa <= 1e-12f) tc = fminf(fmaxf(ff / ee, 0.0f), 1.0f);
    else {
        const float cc = dot(d1, r);
        if (ee <= 1e-12f) sc = fminf(fmaxf(-cc / aa, 0.0f), 1.0f);
        else {
            const float bb = dot(d1, d2), den = aa * ee - bb * bb;
            sc = den > 1e-12f ? fminf(fmaxf((bb * ff - cc * ee) / den, 0.0f), 1.0f) : 0.0f;
            tc = (bb * sc + ff) / ee;
            if (tc < 0.0f) { tc = 0.0f; sc = fminf(fmaxf(-cc / aa, 0.0f), 1.0f); }
            else if (tc > 1.0f) { tc = 1.0f; sc = fminf(fmaxf((bb - cc) / aa, 0.0f), 1.0f); }
        }
    }
    return norm((p1 + d1 * sc) - (p2 + d2 * tc));
}
// lower bound on the distance of (tool shape ti, human shape hi) from their bounding capsules
template <class SM>
__device__ __forceinline__ float bb_pair_bound(const KM& m, const SM& s, int ti, int hi) {
    const float4 a0 = s.tcap[ti][0], a1 = s.tcap[ti][1];
    float4 b0, b1; epi_world_capsule(m, s, s.hum_idx[hi], b0, b1);
    const V3 p1 = mk3(a0.x, a0.y, a0.z), p2 = mk3(b0.x, b0.y, b0.z);
    return seg_seg_distance(p1, mk3(a1.x, a1.y, a1.z) - p1, p2, mk3(b1.x, b1.y, b1.z) - p2) - a0.w - b0.w - 1e-5f;
}
// exact narrowphase distance (narrowphase() of the oracle: GJK on the cores minus the margins, face-normal SAT when the
// cores overlap) of one pair per lane; lanes with active == false only serve the hull scans.  3e38 when the pair is
// farther apart than `bound`.
template <class SM>
__device__ float bb_pair_distance(const KM& m, const SM& s, int ti, int hi, bool active, float bound, int lane) {
    WShape A, B;
    epi_load_shape(m, s, s.tool_idx[ti], A); epi_load_shape(m, s, s.hum_idx[hi], B);
    const float ma = A.s->margin, mb = B.s->margin;
    float dist = 0.0f, gap = 0.0f; V3 ca = A.p, cb = B.p, vout = mk3(0, 0, 0); int gi = 0;
    // cores farther apart than bound + margins cannot beat `bound`; with a negative bound any separated cores qualify
    const int g = gjk_lockstep(A, B, active, lane, fmaxf(bound + ma + mb, 0.0f), dist, ca, cb, vout, gap, gi);
    float best = 3.0e38f; V3 bn = mk3(0, 0, 1), bpa = A.p;
    sat_served(A, B, g == 1, lane, best, bn, bpa);
    if (g == 0) return dist - ma - mb;
    if (g == 1) return -(best > 1.0e38f ? 0.0f : best) - ma - mb;
    return 3.0e38f;
}

__device__ void dump_contacts(const AvgStepArgs& a, int e, const float* scr, int ncontact, float dt, int lane) {
    AvgContact* co = a.contacts + (size_t)e * kMaxC;
    for (int ci = lane; ci < kMaxC; ci += 32) {
        AvgContact c;
        if (ci < ncontact) {
            const float* g = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * ci;
            c.shape_a = __float_as_int(g[10]); c.shape_b = __float_as_int(g[11]);
            for (int k = 0; k < 3; ++k) { c.pos_a[k] = g[k]; c.pos_b[k] = g[3 + k]; c.normal[k] = g[6 + k]; }
            c.dist = g[9]; c.force = g[12] / dt;
        } else { c.shape_a = -1; c.shape_b = -1; for (int k = 0; k < 3; ++k) { c.pos_a[k] = c.pos_b[k] = c.normal[k] = 0; } c.dist = 0; c.force = 0; }
        c.pad[0] = c.pad[1] = c.pad[2] = 0;
        co[ci] = c;
    }
    if (lane == 0) a.ncontacts[e] = ncontact;
}
}  // namespace

__device__ __forceinline__ void epilogue_bb_body(const AvgStepArgs& a, SmEpiBB& s, const int e, const int lane, const KM& m, float* grec, float* scr) {
    const AvgModelHeader* h = m.h;
    for (int i = lane; i < AVG_ENV_STRIDE; i += 32) s.env[i] = grec[i];
    __syncwarp();
    int* env_i = reinterpret_cast<int*>(s.env);
    const int* scr_i = reinterpret_cast<const int*>(scr);
    const int na = h->n_action_robot + h->n_action_human;
    const float* act = a.actions + (size_t)e * na;
    float raw_sq;
    {
        const float av = lane < na ? act[lane] : 0.0f;
        raw_sq = warp_sum(av * av);                          // reward_action uses the raw action, bed_bathing.py:62
    }
    fk_warp(m, s, s.env + AVG_E_Q, lane, h->n_body, s.env + AVG_E_EBODY);
    frames_warp(m, s, lane);
    const float dt = h->dt;
    const float* tf = h->task_f;
    const int ncontact = scr_i[AVG_S_NCS];
    // ---- get_total_force, bed_bathing.py:77-127.  Lane w holds word w of the alive-target bitmap; in round r of a
    //      sweep over the targets lane l looks at target 32 r + l, i.e. bit l of word r.
    uint32_t mask_w = lane < 5 ? (uint32_t)env_i[AVG_E_TARGET_MASK + lane] : 0u;
    V3 up_p, fo_p; Q4 up_q, fo_q;
    frame_cached(s, AVG_F_SHOULDER, up_p, up_q);             // human link 9 / 11 COM frames, bed_bathing.py:383,389
    frame_cached(s, AVG_F_ELBOW, fo_p, fo_q);
    const int n_target = h->n_target, n_upper = h->n_target_upper;
    const float radius = tf[AVG_TF_TARGET_RADIUS];
    float tool_force = 0, tool_force_on_human = 0, total_force_on_human = 0;
    int new_points = 0;
    for (int ci = 0; ci < ncontact; ++ci) {
        const float* c = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * ci;
        const AvgShape* sa = &m.shape[__float_as_int(c[10])]; const AvgShape* sb = &m.shape[__float_as_int(c[11])];
        const float force = c[12] / dt;
        const bool a_tool = sa->ref_body == AVG_REF_TOOL, b_tool = sb->ref_body == AVG_REF_TOOL;
        const bool a_hum = sa->ref_body == AVG_REF_HUMAN, b_hum = sb->ref_body == AVG_REF_HUMAN;
        const bool a_rob = sa->ref_body == AVG_REF_ROBOT, b_rob = sb->ref_body == AVG_REF_ROBOT;
        if (a_tool || b_tool) tool_force += force;                                           // :83-85
        if ((a_rob && b_hum) || (b_rob && a_hum)) total_force_on_human += force;             // :90-91
        if ((a_tool && b_hum) || (b_tool && a_hum)) {                                        // :92-125
            total_force_on_human += force;
            const int link_tool = a_tool ? sa->ref_link : sb->ref_link;
            const int link_hum = a_tool ? sb->ref_link : sa->ref_link;
            if (link_tool != 1) continue;
            tool_force_on_human += force;
            if (link_hum < 0) continue;                                                      // :100-101
            const V3 pos_h = a_tool ? ld3(c + 3) : ld3(c);                                   // positionOnB, B = human
            for (int r = 0; r * 32 < n_target; ++r) {
                const int t = r * 32 + lane;
                const uint32_t word = __shfl_sync(AVG_FULL, mask_w, r);
                bool wiped = false;
                if (t < n_target && ((word >> lane) & 1u)) {
                    const float4 tp = __ldg(&m.target[t]);
                    const V3 tw = t < n_upper ? up_p + qrot(up_q, mk3(tp.x, tp.y, tp.z)) : fo_p + qrot(fo_q, mk3(tp.x, tp.y, tp.z));
                    wiped = norm(pos_h - tw) < radius;
                }
                const unsigned bal = __ballot_sync(AVG_FULL, wiped);
                if (lane == r) mask_w &= ~bal;
                new_points += __popc(bal);
            }
        }
    }
    // ---- reward_distance: min closest-point distance over (tool link, human link) pairs, bed_bathing.py:61.  Every pair
    //      has a lower bound from its bounding capsules; the pair with the smallest bound is measured first and only
    //      pairs whose bound beats the best exact distance so far go through GJK (one pair per lane, lockstep).
    int nt = 0, nh = 0;
    for (int base = 0; base < h->n_shape; base += 32) {
        const int si = base + lane;
        const int rb = si < h->n_shape ? m.shape[si].ref_body : -1;
        const unsigned bt = __ballot_sync(AVG_FULL, rb == AVG_REF_TOOL), bh = __ballot_sync(AVG_FULL, rb == AVG_REF_HUMAN);
        if (rb == AVG_REF_TOOL) { const int k = nt + __popc(bt & ((1u << lane) - 1)); if (k < 8) s.tool_idx[k] = (uint8_t)si; }
        if (rb == AVG_REF_HUMAN) { const int k = nh + __popc(bh & ((1u << lane) - 1)); if (k < 120) s.hum_idx[k] = (uint8_t)si; }
        nt += __popc(bt); nh += __popc(bh);
    }
    nt = min(nt, 8); nh = min(nh, 120);
    __syncwarp();
    if (lane < nt) { float4 c0, c1; epi_world_capsule(m, s, s.tool_idx[lane], c0, c1); s.tcap[lane][0] = c0; s.tcap[lane][1] = c1; }
    __syncwarp();
    const int npair = nt * nh;
    float closest = tf[AVG_TF_CLOSEST_RANGE];
    if (npair > 0) {
        float lb_min = 3.0e38f; int p_min = 0x7fffffff;
        for (int p = lane; p < npair; p += 32) {
            const float lb = bb_pair_bound(m, s, p / nh, p % nh);
            if (lb < lb_min) { lb_min = lb; p_min = p; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ol = __shfl_xor_sync(AVG_FULL, lb_min, o); const int op = __shfl_xor_sync(AVG_FULL, p_min, o);
            if (ol < lb_min || (ol == lb_min && op < p_min)) { lb_min = ol; p_min = op; }
        }
        const int p0 = p_min;
        float best = bb_pair_distance(m, s, p0 / nh, p0 % nh, lane == 0, 3.0e37f, lane);
        best = __shfl_sync(AVG_FULL, best, 0);
        int p = lane;
        while (true) {
            // the capsule bound is a bound on separation; once the best pair penetrates (face-normal SAT depths are not
            // minimal translations) only pairs whose bounding capsules are apart can be skipped
            while (p < npair && (p == p0 || bb_pair_bound(m, s, p / nh, p % nh) >= fmaxf(best, 0.0f))) p += 32;
            const bool active = p < npair;
            if (!__any_sync(AVG_FULL, active)) break;
            const int pp = active ? p : 0;
            const float d = bb_pair_distance(m, s, pp / nh, pp % nh, active, best, lane);
            float mine = active ? fminf(best, d) : best;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine = fminf(mine, __shfl_xor_sync(AVG_FULL, mine, o));
            best = mine;
            p += 32;
        }
        closest = fminf(closest, best);
    }
    if (lane < 5) reinterpret_cast<uint32_t*>(grec)[AVG_E_TARGET_MASK + lane] = mask_w;
    if (lane == 0) {
        fill_obs_bb(m, s, tool_force, total_force_on_human, tool_force_on_human);
        V3 tool; Q4 tq; frame_cached(s, AVG_F_TOOL_TIP, tool, tq);
        const int tb = m.frame[AVG_F_TOOL_TIP].body;
        const float* tv = s.env + AVG_E_QD + m.body[tb].dof;
        const float ee_vel = norm(ld3(tv) + cross(ld3(tv + 3), tool - ld3(s.bp[tb])));       // bed_bathing.py:54
        const float pref = tf[AVG_TF_C_V] * (-ee_vel) + tf[AVG_TF_C_F] * (-(total_force_on_human - tool_force_on_human))
                         + tf[AVG_TF_C_HF] * (tool_force_on_human < tf[AVG_TF_FORCE_CAP] ? 0.0f : -tool_force_on_human);   // env.py:412-448
        const float reward_distance = -closest;
        const float reward_action = -raw_sq;
        const float task_success = s.env[AVG_E_TASK_SUCCESS] + (float)new_points;
        grec[AVG_E_TASK_SUCCESS] = task_success;
        const float reward = tf[AVG_TF_DISTANCE_W] * reward_distance + tf[AVG_TF_ACTION_W] * reward_action
                           + tf[AVG_TF_SCRATCH_W] * (float)new_points + pref;                // bed_bathing.py:65
        int* grec_i = reinterpret_cast<int*>(grec);
        grec[AVG_E_EPISODE_RETURN] = s.env[AVG_E_EPISODE_RETURN] + reward;
        grec_i[AVG_E_ITERATION] = env_i[AVG_E_ITERATION] + 1;                               // env.py:351
        grec_i[AVG_E_OVERFLOW] = env_i[AVG_E_OVERFLOW] | scr_i[AVG_S_OVERFLOW];
        grec_i[AVG_E_SOLVER_ITERS] = scr_i[AVG_S_ITERS];
        grec_i[AVG_E_NCAND] = scr_i[AVG_S_NCAND];
        a.reward[e] = reward;
        const float success = task_success >= tf[AVG_TF_SUCCESS_THR] ? 1.0f : 0.0f;          // bed_bathing.py:72
        a.info[2 * e] = total_force_on_human;
        a.info[2 * e + 1] = success;
        if (a.done) a.done[e] = (a.time_limit > 0 && env_i[AVG_E_ITERATION] + 1 >= a.time_limit) ? 1 : 0;     // TimeLimit per environment, see avg_epilogue_kernel
        if (a.terms) {
            float* tr = a.terms + 8 * (size_t)e;
            tr[0] = total_force_on_human; tr[1] = success; tr[2] = tool_force; tr[3] = tool_force_on_human;
            tr[4] = reward_distance; tr[5] = reward_action; tr[6] = (float)new_points; tr[7] = pref;
        }
    }
    __syncwarp();
    const int nobs = h->n_obs_robot + h->n_obs_human;
    for (int i = lane; i < nobs; i += 32) a.obs[(size_t)e * nobs + i] = s.obs[i];
    if (a.contacts) dump_contacts(a, e, scr, ncontact, dt, lane);
}
__global__ void __launch_bounds__(32 * AVG_K_WARPS_PER_BLOCK)
avg_epilogue_bb_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmEpiBB)
    epilogue_bb_body(a, s, e, lane, m, grec, scr);
}

// =================================================================================================================
// Feeding / Drinking epilogue (feeding.py:56-142, drinking.py:57-157): get_total_force, get_food_rewards /
// get_water_rewards with one lane per particle, human_preferences (env.py:412-448, feeding branch), reward, observation, info
// =================================================================================================================
namespace {
// feeding.py:123-142 / drinking.py:138-157 (lane 0 fills s.obs)
template <class SM>
__device__ void fill_obs_fd(const KM& m, SM& s, V3 mouth, float tool_force_on_human, float robot_force_on_human) {
    const AvgModelHeader* h = m.h;
    const int nj = h->n_jdof;
    V3 torso, tool, head, chest; Q4 tq, hq, dq;
    frame_cached(s, AVG_F_TORSO, torso, dq);
    frame_cached(s, AVG_F_TOOL_TIP, tool, tq);
    frame_cached(s, AVG_F_HEAD, head, hq);
    frame_cached(s, AVG_F_CHEST, chest, dq);
    float* o = s.obs; int k = 0;
    V3 t;
    t = tool - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    o[k++] = tq.x; o[k++] = tq.y; o[k++] = tq.z; o[k++] = tq.w;
    t = tool - mouth; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    for (int i = 0; i < nj; ++i) if (m.dof[i].action >= 0 && m.dof[i].action < h->n_action_robot) o[k++] = s.env[AVG_E_Q + m.body[m.dof[i].body].qidx];
    t = head - torso; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
    o[k++] = hq.x; o[k++] = hq.y; o[k++] = hq.z; o[k++] = hq.w;
    o[k++] = tool_force_on_human;
    if (h->human_control) {
        t = tool - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        o[k++] = tq.x; o[k++] = tq.y; o[k++] = tq.z; o[k++] = tq.w;
        t = tool - mouth; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        const int hq0 = k;
        for (int i = 0; i < 4; ++i) o[k++] = 0.0f;
        for (int i = 0; i < nj; ++i) { const int sl = m.dof[i].human_slot; if (sl >= 0 && sl < 4) o[hq0 + sl] = s.env[AVG_E_Q + m.body[m.dof[i].body].qidx]; }
        t = head - chest; o[k++] = t.x; o[k++] = t.y; o[k++] = t.z;
        o[k++] = hq.x; o[k++] = hq.y; o[k++] = hq.z; o[k++] = hq.w;
        o[k++] = robot_force_on_human; o[k++] = tool_force_on_human;
    }
}
// btQuaternion::getEulerZYX roll, what p.getEulerFromQuaternion(q)[0] returns (drinking.py:71)
__device__ __forceinline__ float quat_roll(Q4 q) {
    const float sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
    const float sarg = -2.0f * (q.x * q.z - q.w * q.y) / (sqx + sqy + sqz + sqw);
    if (sarg <= -0.99999f || sarg >= 0.99999f) return 0.0f;
    return atan2f(2.0f * (q.y * q.z + q.w * q.x), sqw - sqx - sqy + sqz);
}
template <class SM>
__device__ __forceinline__ V3 fd_mouth(const KM& m, const SM& s) {
    V3 hp; Q4 hq; frame_cached(s, AVG_F_HEAD, hp, hq);
    const float* tf = m.h->task_f;
    return hp + qrot(hq, mk3(tf[AVG_TF_MOUTH], tf[AVG_TF_MOUTH + 1], tf[AVG_TF_MOUTH + 2]));
}
}  // namespace

__global__ void __launch_bounds__(32 * AVG_K_WARPS_PER_BLOCK)
avg_epilogue_fd_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmEpi)
    for (int i = lane; i < AVG_ENV_STRIDE; i += 32) s.env[i] = grec[i];
    __syncwarp();
    int* env_i = reinterpret_cast<int*>(s.env);
    const int* scr_i = reinterpret_cast<const int*>(scr);
    const int na = h->n_action_robot + h->n_action_human;
    const float* act = a.actions + (size_t)e * na;
    float raw_sq;
    {
        const float av = lane < na ? act[lane] : 0.0f;
        raw_sq = warp_sum(av * av);                          // reward_action uses the raw action, feeding.py:69
    }
    fk_warp(m, s, s.env + AVG_E_Q, lane, h->n_body, s.env + AVG_E_EBODY);
    frames_warp(m, s, lane);
    const float dt = h->dt;
    const float* tf = h->task_f;
    const bool drinking = h->task == AVG_TASK_DRINKING;
    const int ncontact = scr_i[AVG_S_NCS];
    // get_total_force, feeding.py:83-90
    float robot_force_on_human = 0, tool_force_on_human = 0;
    for (int ci = 0; ci < ncontact; ++ci) {
        const float* c = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * ci;
        const AvgShape* sa = &m.shape[__float_as_int(c[10])]; const AvgShape* sb = &m.shape[__float_as_int(c[11])];
        const float force = c[12] / dt;
        const bool a_tool = sa->ref_body == AVG_REF_TOOL, b_tool = sb->ref_body == AVG_REF_TOOL;
        const bool a_hum = sa->ref_body == AVG_REF_HUMAN, b_hum = sb->ref_body == AVG_REF_HUMAN;
        const bool a_rob = sa->ref_body == AVG_REF_ROBOT, b_rob = sb->ref_body == AVG_REF_ROBOT;
        if ((a_rob && b_hum) || (b_rob && a_hum)) robot_force_on_human += force;
        if ((a_tool && b_hum) || (b_tool && a_hum)) tool_force_on_human += force;
    }
    const V3 mouth = fd_mouth(m, s);
    V3 tool; Q4 tq; frame_cached(s, AVG_F_TOOL_TIP, tool, tq);
    // the cup's frame and the ends of its cylinder, drinking.py:97-102
    V3 top = tool, bottom = tool; Q4 cup_q = tq;
    if (drinking) {
        const V3 cup_p = tool + qrot(tq, mk3(0.0f, 0.06f, 0.0f));
        cup_q = qnormalize(qmul(tq, qaxis(mk3(1, 0, 0), 1.57079632679f)));
        top = cup_p + qrot(cup_q, mk3(0, 0, tf[AVG_TF_CUP_TOP])); bottom = cup_p + qrot(cup_q, mk3(0, 0, tf[AVG_TF_CUP_BOTTOM]));
    }
    // get_food_rewards / get_water_rewards: lane l looks at particles l and l + 32
    float* prec = a.part + (size_t)e * AVG_P_STRIDE;
    uint32_t* prec_u = reinterpret_cast<uint32_t*>(prec);
    uint32_t alive[2] = {prec_u[AVG_P_ALIVE], prec_u[AVG_P_ALIVE + 1]}, hitm[2] = {prec_u[AVG_P_HIT], prec_u[AVG_P_HIT + 1]};
    const uint32_t th[2] = {prec_u[AVG_P_TOUCH_HUMAN], prec_u[AVG_P_TOUCH_HUMAN + 1]}, tsp[2] = {prec_u[AVG_P_TOUCH_SPILL], prec_u[AVG_P_TOUCH_SPILL + 1]};
    int n_eat = 0, n_spill = 0, n_hit = 0;
    float vel_sum = 0.0f;
    uint32_t ev_eat[2], ev_spill[2], ev_hit[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int p = lane + 32 * k;
        bool eat = false, spill = false, hit = false, hit_remove = false;
        float speed = 0.0f;
        if (p < h->n_particle && ((alive[k] >> lane) & 1u)) {
            const V3 x = mk3(prec[AVG_P_POS + p], prec[AVG_P_POS + 64 + p], prec[AVG_P_POS + 128 + p]);
            bool inside = false;
            if (drinking) {                                   // util.points_in_cylinder(top, bottom, 0.05, x), util.py:107-110
                const V3 vec = bottom - top;
                inside = dot(x - top, vec) >= 0.0f && dot(x - bottom, vec) <= 0.0f && norm(cross(x - top, vec)) <= tf[AVG_TF_CUP_RADIUS] * norm(vec);
            }
            if (!inside) {
                if (norm(mouth - x) < tf[AVG_TF_EAT_RADIUS]) {                                                    // feeding.py:102, drinking.py:114
                    eat = true;
                    if (!drinking) speed = norm(mk3(prec[AVG_P_VEL + p], prec[AVG_P_VEL + 64 + p], prec[AVG_P_VEL + 128 + p]));   // Drinking reads it after the teleport zeroed it
                } else if (x.z < tf[AVG_TF_Z_MIN] || (!drinking && ((tsp[k] >> lane) & 1u))) spill = true;      // feeding.py:111, drinking.py:124
                else if ((th[k] >> lane) & 1u) {
                    if (drinking) { hit = true; hit_remove = true; }                                             // drinking.py:131-134
                    else if (!((hitm[k] >> lane) & 1u)) hit = true;                                              // feeding.py:116-119
                }
            }
        }
        ev_eat[k] = __ballot_sync(AVG_FULL, eat); ev_spill[k] = __ballot_sync(AVG_FULL, spill); ev_hit[k] = __ballot_sync(AVG_FULL, hit);
        const uint32_t rem = ev_eat[k] | ev_spill[k] | __ballot_sync(AVG_FULL, hit_remove);
        alive[k] &= ~rem;
        if (!drinking) hitm[k] |= ev_hit[k];
        n_eat += __popc(ev_eat[k]); n_spill += __popc(ev_spill[k]); n_hit += __popc(ev_hit[k]);
        vel_sum += warp_sum(speed);
    }
    if (lane < 2) {
        prec_u[AVG_P_ALIVE + lane] = alive[lane]; prec_u[AVG_P_HIT + lane] = hitm[lane];
        prec_u[AVG_P_EV_EAT + lane] = ev_eat[lane]; prec_u[AVG_P_EV_SPILL + lane] = ev_spill[lane]; prec_u[AVG_P_EV_HIT + lane] = ev_hit[lane];
    }
    if (lane == 0) {
        fill_obs_fd(m, s, mouth, tool_force_on_human, robot_force_on_human);
        const int tb = m.frame[AVG_F_TOOL_TIP].body;
        const float* tv = s.env + AVG_E_QD + m.body[tb].dof;
        const float ee_vel = norm(ld3(tv));                                                   // getBaseVelocity(spoon)[0], feeding.py:59
        const float food_reward = tf[AVG_TF_EAT_REWARD] * (float)n_eat + tf[AVG_TF_SPILL_REWARD] * (float)n_spill;
        const float hit_reward = -(float)n_hit;
        // human_preferences, env.py:412-448: total_force_on_human = robot force, tool_force_at_target = tool force (feeding.py:63)
        const float pref = tf[AVG_TF_C_V] * (-ee_vel) + tf[AVG_TF_C_F] * (-robot_force_on_human)
                         + tf[AVG_TF_C_HF] * (tool_force_on_human < tf[AVG_TF_FORCE_CAP] ? 0.0f : -tool_force_on_human)
                         + tf[AVG_TF_C_FD] * hit_reward + tf[AVG_TF_C_FDV] * (-vel_sum);
        float reward_distance, reward_tilt = 0.0f;
        if (drinking) {
            reward_distance = -norm(mouth - top);                                             // drinking.py:68
            const float roll = quat_roll(cup_q);
            reward_tilt = tf[AVG_TF_TILT_SIGN] > 0.0f ? -fabsf(roll + 1.57079632679f) : -fabsf(roll - 1.57079632679f);   // :72
        } else reward_distance = -norm(mouth - tool);                                         // feeding.py:68
        const float reward_action = -raw_sq;
        const float task_success = s.env[AVG_E_TASK_SUCCESS] + (float)n_eat;
        grec[AVG_E_TASK_SUCCESS] = task_success;
        const float reward = tf[AVG_TF_DISTANCE_W] * reward_distance + tf[AVG_TF_ACTION_W] * reward_action + tf[AVG_TF_TILT_W] * reward_tilt
                           + tf[AVG_TF_FOOD_W] * food_reward + pref;                          // feeding.py:71, drinking.py:74
        int* grec_i = reinterpret_cast<int*>(grec);
        grec[AVG_E_EPISODE_RETURN] = s.env[AVG_E_EPISODE_RETURN] + reward;
        st3(grec + AVG_E_TARGET_POS, mouth);
        grec_i[AVG_E_ITERATION] = env_i[AVG_E_ITERATION] + 1;                                // env.py:351
        grec_i[AVG_E_OVERFLOW] = env_i[AVG_E_OVERFLOW] | scr_i[AVG_S_OVERFLOW] | reinterpret_cast<const int*>(prec)[AVG_P_NCONTACT + 1];
        grec_i[AVG_E_SOLVER_ITERS] = scr_i[AVG_S_ITERS];
        grec_i[AVG_E_NCAND] = scr_i[AVG_S_NCAND];
        a.reward[e] = reward;
        const float success = task_success >= tf[AVG_TF_SUCCESS_THR] ? 1.0f : 0.0f;           // feeding.py:76
        a.info[2 * e] = robot_force_on_human + tool_force_on_human;
        a.info[2 * e + 1] = success;
        if (a.done) a.done[e] = (a.time_limit > 0 && env_i[AVG_E_ITERATION] + 1 >= a.time_limit) ? 1 : 0;
        if (a.terms) {
            float* tr = a.terms + 8 * (size_t)e;
            tr[0] = robot_force_on_human + tool_force_on_human; tr[1] = success; tr[2] = robot_force_on_human; tr[3] = tool_force_on_human;
            tr[4] = reward_distance; tr[5] = reward_action; tr[6] = food_reward + tf[AVG_TF_TILT_W] * reward_tilt; tr[7] = pref;
        }
    }
    __syncwarp();
    const int nobs = h->n_obs_robot + h->n_obs_human;
    for (int i = lane; i < nobs; i += 32) a.obs[(size_t)e * nobs + i] = s.obs[i];
    if (a.contacts) dump_contacts(a, e, scr, ncontact, dt, lane);
}

// =================================================================================================================
// particle broadphase (Feeding / Drinking): one warp per environment, one lane per particle (two when there are 64).
// Particle-vs-shape candidates go to the narrowphase queue (the sphere template against the shape: GJK like any other pair);
// particle-vs-particle contacts are closed form and are written directly.  Canonical order of both lists: particle index,
// then shape-table index / second particle index.
// =================================================================================================================
namespace {
constexpr int kPCandPerParticle = 16;      // particle-vs-shape candidates kept per particle after the box and face-plane culls (more: flagged)
struct __align__(16) SmPCol {
    float bp[32][3]; float bq[32][4];
    float4 saabb[kMaxMS][2];
    float px[64][3];
    uint16_t cand[64][kPCandPerParticle];  // candidate shapes of each particle, ascending shape index (one enumeration, then the prefix sum)
    uint8_t near_idx[256];
    uint32_t cmask[16][AVG_MAX_COMPOUND][3];   // few particles (Feeding): children of compound c that pass the culls for particle p, found by the whole warp
};
}  // namespace

__global__ void __launch_bounds__(32 * AVG_K_WARPS_PER_BLOCK)
avg_pcollide_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmPCol)
    AVG_MASK_CHECK
    const int np = h->n_particle;
    if (np <= 0 || !a.part || !a.pscratch) return;
    const int nb = h->n_body, nms = h->n_mshape, nstat = h->n_shape - nms;
    float* ps = a.pscratch + (size_t)e * AVG_PS_STRIDE;
    int* ps_i = reinterpret_cast<int*>(ps);
    const float* prec = a.part + (size_t)e * AVG_P_STRIDE;
    const AvgShape* PS = &m.shape[h->pshape];
    const float r = PS->radius, pthr = PS->thr, infl = r + pthr;
    if (lane < nb + h->n_ebody) {
        const float4* gp = reinterpret_cast<const float4*>(scr + AVG_S_POSE) + 2 * lane;
        const float4 p4 = gp[0], q4 = gp[1];
        s.bp[lane][0] = p4.x; s.bp[lane][1] = p4.y; s.bp[lane][2] = p4.z;
        s.bq[lane][0] = q4.x; s.bq[lane][1] = q4.y; s.bq[lane][2] = q4.z; s.bq[lane][3] = q4.w;
    }
    const uint32_t alive0 = __float_as_uint(prec[AVG_P_ALIVE]), alive1 = __float_as_uint(prec[AVG_P_ALIVE + 1]);
    bool live[2]; V3 x[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int p = lane + 32 * k;
        live[k] = p < np && (((k ? alive1 : alive0) >> lane) & 1u);
        x[k] = live[k] ? mk3(prec[AVG_P_POS + p], prec[AVG_P_POS + 64 + p], prec[AVG_P_POS + 128 + p]) : mk3(0, 0, 0);
        s.px[p][0] = x[k].x; s.px[p][1] = x[k].y; s.px[p][2] = x[k].z;
    }
    __syncwarp();
    if (lane < nms) {                        // world AABBs of the top-level moving shapes (compounds: of the whole compound)
        const AvgShape* S = &m.shape[lane];
        const Q4 bq = ldq(s.bq[S->body]);
        const V3 p = ld3(s.bp[S->body]) + qrot(bq, ld3(S->pos));
        const M3 R = qmat(qnormalize(qmul(bq, ldq(S->quat))));
        const V3 lc = ld3(S->aabb_c), lh = ld3(S->aabb_h);
        const V3 c = p + mmul(R.m, lc);
        s.saabb[lane][0] = make_float4(c.x, c.y, c.z, fabsf(R.m[0]) * lh.x + fabsf(R.m[1]) * lh.y + fabsf(R.m[2]) * lh.z);
        s.saabb[lane][1] = make_float4(fabsf(R.m[3]) * lh.x + fabsf(R.m[4]) * lh.y + fabsf(R.m[5]) * lh.z,
                                       fabsf(R.m[6]) * lh.x + fabsf(R.m[7]) * lh.y + fabsf(R.m[8]) * lh.z, 0.0f, 0.0f);
    }
    // union box of the live particles -> static shapes near any of them
    float ulo[3], uhi[3];
    {
        auto key = [](float f) { const unsigned b = __float_as_uint(f); return b ^ ((unsigned)((int)b >> 31) | 0x80000000u); };
        auto unkey = [](unsigned q) { return __uint_as_float(q ^ ((q >> 31) ? 0x80000000u : 0xffffffffu)); };
        const float xs[2][3] = {{x[0].x, x[0].y, x[0].z}, {x[1].x, x[1].y, x[1].z}};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float lo = 3.0e38f, hi = -3.0e38f;
            if (live[0]) { lo = xs[0][k]; hi = xs[0][k]; }
            if (live[1]) { lo = fminf(lo, xs[1][k]); hi = fmaxf(hi, xs[1][k]); }
            ulo[k] = unkey(__reduce_min_sync(AVG_FULL, key(lo))) - infl; uhi[k] = unkey(__reduce_max_sync(AVG_FULL, key(hi))) + infl;
        }
    }
    int nnear = 0;
    for (int base = 0; base < nstat; base += 32) {
        const int si = base + lane;
        bool near = false;
        if (si < nstat) {
            const float4* rp = reinterpret_cast<const float4*>(&m.bps[si]);
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            near = r0.x + r0.w >= ulo[0] && r0.x - r0.w <= uhi[0] && r0.y + r1.x >= ulo[1] && r0.y - r1.x <= uhi[1] &&
                   r0.z + r1.y >= ulo[2] && r0.z - r1.y <= uhi[2];
        }
        const unsigned bal = __ballot_sync(AVG_FULL, near);
        if (near) s.near_idx[nnear + __popc(bal & ((1u << lane) - 1))] = (uint8_t)si;
        nnear += __popc(bal);
    }
    __syncwarp();
    // the culls of one (particle, compound child) pair: the child's body-frame box, then the face-plane bound -- a point outside
    // a hull is at least max_i (n_i . x - d_i) away from it, so children whose planes already put the particle beyond the contact
    // distance never reach GJK (thin VHACD pieces have loose boxes)
    auto child_passes = [&](V3 pl, int first, int k) -> bool {
        const float4 c4 = __ldg(&m.caabb[2 * (first - h->n_shape + k)]), h4 = __ldg(&m.caabb[2 * (first - h->n_shape + k) + 1]);
        if (!(fabsf(pl.x - c4.x) <= h4.x + infl && fabsf(pl.y - c4.y) <= h4.y + infl && fabsf(pl.z - c4.z) <= h4.z + infl)) return false;
        const AvgShape* C = &m.shape[first + k];
        const V3 xs = qrot_inv(ldq(C->quat), pl - ld3(C->pos));
        const float4* pln = reinterpret_cast<const float4*>(m.plane) + C->plane_off;
        float lb = -3.0e38f;
        for (int i = 0; i < C->plane_cnt; ++i) { const float4 q4 = __ldg(pln + i); lb = fmaxf(lb, fmaf(q4.x, xs.x, fmaf(q4.y, xs.y, q4.z * xs.z)) - q4.w); }
        return lb < infl + C->margin + 1e-6f;
    };
    // Few particles (Feeding: 8 lanes would walk 134 children each while 24 idle): the whole warp tests the children of a
    // compound for one particle at a time, one child per lane, and leaves the survivors as bit masks for the enumeration below.
    const bool coop = np <= 16;
    if (coop) {
        int ncomp = 0;
        for (int sa = 0; sa < nms && ncomp < AVG_MAX_COMPOUND; ++sa) {
            const AvgShape* S = &m.shape[sa];
            if (S->type != AVG_SHAPE_COMPOUND) continue;
            const float4 a0 = s.saabb[sa][0], a1 = s.saabb[sa][1];
            const Q4 cq = ldq(s.bq[S->body]); const V3 cp = ld3(s.bp[S->body]);
            const int first = S->vert_off, cnt = min(S->vert_cnt, 96);
            for (int p = 0; p < np; ++p) {
                const bool alive_p = (alive0 >> p) & 1u;
                const V3 xp = ld3(s.px[p]);
                const bool inbox = alive_p && fabsf(xp.x - a0.x) <= a0.w + infl && fabsf(xp.y - a0.y) <= a1.x + infl && fabsf(xp.z - a0.z) <= a1.y + infl;
                for (int r3 = 0; r3 < 3; ++r3) {
                    const int k = 32 * r3 + lane;
                    const bool pass = inbox && k < cnt && child_passes(qrot_inv(cq, xp - cp), first, k);
                    const unsigned w = __ballot_sync(AVG_FULL, pass);
                    if (lane == 0) s.cmask[p][ncomp][r3] = w;
                }
            }
            ++ncomp;
        }
        __syncwarp();
    }
    // candidate enumeration of one particle in ascending shape-table order; f(shape index) is called for every candidate
    auto enumerate = [&](V3 xp, int pidx, auto&& f) {
        for (int sa = 0; sa < nms; ++sa) {
            if (m.shape[sa].type == AVG_SHAPE_COMPOUND) continue;
            const float4 a0 = s.saabb[sa][0], a1 = s.saabb[sa][1];
            if (fabsf(xp.x - a0.x) <= a0.w + infl && fabsf(xp.y - a0.y) <= a1.x + infl && fabsf(xp.z - a0.z) <= a1.y + infl) f(sa);
        }
        for (int k = 0; k < nnear; ++k) {
            const int si = s.near_idx[k];
            const float4* rp = reinterpret_cast<const float4*>(&m.bps[si]);
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            if (fabsf(xp.x - r0.x) <= r0.w + infl && fabsf(xp.y - r0.y) <= r1.x + infl && fabsf(xp.z - r0.z) <= r1.y + infl) f(nms + si);
        }
        int ncomp = 0;
        for (int sa = 0; sa < nms; ++sa) {
            const AvgShape* S = &m.shape[sa];
            if (S->type != AVG_SHAPE_COMPOUND) continue;
            const int first = S->vert_off, cnt = S->vert_cnt;
            if (coop && ncomp < AVG_MAX_COMPOUND && cnt <= 96) {               // survivors found by the whole warp above
                for (int r3 = 0; r3 < 3; ++r3) {
                    unsigned w = s.cmask[pidx][ncomp][r3];
                    while (w) { const int k = __ffs(w) - 1; w &= w - 1; f(first + 32 * r3 + k); }
                }
                ++ncomp;
                continue;
            }
            ++ncomp;
            const float4 a0 = s.saabb[sa][0], a1 = s.saabb[sa][1];
            if (!(fabsf(xp.x - a0.x) <= a0.w + infl && fabsf(xp.y - a0.y) <= a1.x + infl && fabsf(xp.z - a0.z) <= a1.y + infl)) continue;
            const V3 pl = qrot_inv(ldq(s.bq[S->body]), xp - ld3(s.bp[S->body]));          // the particle in the compound's body frame
            for (int k = 0; k < cnt; ++k) if (child_passes(pl, first, k)) f(first + k);
        }
    };
    int cnt[2] = {0, 0};
    int cand_over = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) if (live[k]) {
        const int p = lane + 32 * k;
        enumerate(x[k], p, [&](int sb) { if (cnt[k] < kPCandPerParticle) s.cand[p][cnt[k]++] = (uint16_t)sb; else cand_over = 16; });
    }
    // particle-major offsets: particles 0..31 first, then 32..63
    int inc0 = cnt[0], inc1 = cnt[1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v0 = __shfl_up_sync(AVG_FULL, inc0, o), v1 = __shfl_up_sync(AVG_FULL, inc1, o);
        if (lane >= o) { inc0 += v0; inc1 += v1; }
    }
    const int tot0 = __shfl_sync(AVG_FULL, inc0, 31), tot1 = __shfl_sync(AVG_FULL, inc1, 31);
    int off[2] = {inc0 - cnt[0], tot0 + inc1 - cnt[1]};
    int total = tot0 + tot1, overflow = __any_sync(AVG_FULL, cand_over != 0) ? 16 : 0;
    if (total > AVG_MAX_PCAND) { overflow |= 16; total = AVG_MAX_PCAND; }
    int qbase = 0;
    if (lane == 0 && total > 0) qbase = atomicAdd(a.np_count, total);
    qbase = __shfl_sync(AVG_FULL, qbase, 0);
    if (qbase + total > a.np_capacity) { overflow |= 1; total = max(0, min(total, a.np_capacity - qbase)); }
#pragma unroll
    for (int k = 0; k < 2; ++k) if (live[k]) {
        const int p = lane + 32 * k;
        int o = off[k];
        for (int c = 0; c < cnt[k]; ++c, ++o)
            if (o < total) { AvgNpItem it; it.env = e; it.pair = (AVG_NP_PARTICLE | (uint32_t)p) | ((uint32_t)s.cand[p][c] << 16); it.slot = o; it.cert = -1; a.np_queue[qbase + o] = it; }
    }
    // particle-particle contacts: spheres of equal radius, closed form
    const float reach = 2.0f * r + pthr;
    int pc[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < 2; ++k) if (live[k]) {
        const int p = lane + 32 * k;
        for (int q = p + 1; q < np; ++q) {
            if (!(((q < 32 ? alive0 : alive1) >> (q & 31)) & 1u)) continue;
            const V3 d = x[k] - ld3(s.px[q]);
            pc[k] += dot(d, d) < reach * reach ? 1 : 0;
        }
    }
    int pi0 = pc[0], pi1 = pc[1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v0 = __shfl_up_sync(AVG_FULL, pi0, o), v1 = __shfl_up_sync(AVG_FULL, pi1, o);
        if (lane >= o) { pi0 += v0; pi1 += v1; }
    }
    const int pt0 = __shfl_sync(AVG_FULL, pi0, 31), pt1 = __shfl_sync(AVG_FULL, pi1, 31);
    int poff[2] = {pi0 - pc[0], pt0 + pi1 - pc[1]};
    int npp = pt0 + pt1;
    if (npp > AVG_PS_MAXPP) { overflow |= 8; npp = AVG_PS_MAXPP; }
    float4* pp = reinterpret_cast<float4*>(ps + AVG_PS_PP);
#pragma unroll
    for (int k = 0; k < 2; ++k) if (live[k]) {
        const int p = lane + 32 * k;
        int o = poff[k];
        for (int q = p + 1; q < np; ++q) {
            if (!(((q < 32 ? alive0 : alive1) >> (q & 31)) & 1u)) continue;
            const V3 d = x[k] - ld3(s.px[q]);
            const float dd = dot(d, d);
            if (!(dd < reach * reach)) continue;
            if (o < npp) {
                const float dn = sqrtf(dd);
                const V3 n = dn > 1e-9f ? d * (1.0f / dn) : mk3(0, 0, 1);
                pp[2 * o] = make_float4(n.x, n.y, n.z, dn - 2.0f * r);
                pp[2 * o + 1] = make_float4(__int_as_float(p), __int_as_float(q), 0.0f, 0.0f);
            }
            ++o;
        }
    }
    if (lane == 0) { ps_i[AVG_PS_NCAND] = total; ps_i[AVG_PS_NPP] = npp; ps_i[AVG_PS_OVERFLOW] = overflow; }
}

// initial observation after reset (scratch_itch.py:268): FK + target + _get_obs([0],[0,0])
__global__ void __launch_bounds__(32 * AVG_K_WARPS_PER_BLOCK)
avg_reset_obs_kernel(AvgStepArgs a) {
    AVG_KERNEL_PREAMBLE(SmEpi)
    if (a.mask && !a.mask[e]) return;
    for (int i = lane; i < AVG_ENV_STRIDE; i += 32) s.env[i] = grec[i];
    __syncwarp();
    const int* env_i = reinterpret_cast<const int*>(s.env);
    fk_warp(m, s, s.env + AVG_E_Q, lane, h->n_body, s.env + AVG_E_EBODY);
    frames_warp(m, s, lane);
    if (lane == 0) {
        if (h->task == AVG_TASK_BED_BATHING) fill_obs_bb(m, s, 0.0f, 0.0f, 0.0f);                        // bed_bathing.py:350
        else if (h->task == AVG_TASK_FEEDING || h->task == AVG_TASK_DRINKING) {                           // feeding.py:325
            const V3 mouth = fd_mouth(m, s);
            fill_obs_fd(m, s, mouth, 0.0f, 0.0f);
            st3(grec + AVG_E_TARGET_POS, mouth);
        }
        else {
            V3 lp; Q4 lq; frame_cached(s, env_i[AVG_E_LIMB_FRAME], lp, lq);
            const V3 tgt = lp + qrot(lq, ld3(s.env + AVG_E_TARGET_ON_ARM));
            fill_obs(m, s, tgt, 0.0f, 0.0f, 0.0f);
            st3(grec + AVG_E_TARGET_POS, tgt);
        }
    }
    __syncwarp();
    const int nobs = h->n_obs_robot + h->n_obs_human;
    for (int i = lane; i < nobs; i += 32) a.obs[(size_t)e * nobs + i] = s.obs[i];
}

// Episode reset on the device: the random draws of ScratchItchEnv.reset (scratch_itch.py:155-162,230-256,275-287),
// create_new_world (world_creation.py:66-72) and setup_human_joints (world_creation.py:136-141,172) restated as in
// compiler/reset.py sample_states(), with counter-based random numbers (AVG_RNG_MIX) so that any subset of
// environments can be reset without a host round trip.  One thread per environment.
namespace {
__device__ __forceinline__ uint32_t reset_u32(uint32_t seed, uint32_t env, uint32_t episode, uint32_t k) {
    uint32_t h = seed;
    AVG_RNG_MIX(h); h ^= env * 0x9e3779b9u;
    AVG_RNG_MIX(h); h ^= episode * 0x7f4a7c15u;
    AVG_RNG_MIX(h); h ^= k * 0x94d049bbu;
    AVG_RNG_MIX(h);
    return h;
}
__device__ __forceinline__ float reset_u01(uint32_t seed, uint32_t env, uint32_t episode, uint32_t k) {
    return (float)(reset_u32(seed, env, episode, k) >> 8) * (1.0f / 16777216.0f);
}
}  // namespace

__global__ void __launch_bounds__(128)
avg_reset_kernel(AvgResetArgs r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n_env) return;
    if (r.mask && !r.mask[e]) return;
    const uint32_t ep = (uint32_t)(r.episode[e] + 1);
    r.episode[e] = (int32_t)ep;
    const uint32_t sd = r.seed, ue = (uint32_t)e;
    // gender (scratch_itch.py:156, bed_bathing.py:191); BedBathing: x one variant per robot base pose (env.py:511-513)
    const uint32_t pick = reset_u32(sd, ue, ep, 17);
    const int npg = r.n_per_gender;
    const int v = (int)(reset_u32(sd, ue, ep, 0) % (uint32_t)min(r.n_variants, 2)) * npg + (int)(pick % (uint32_t)npg);
    const AvgResetTable* T = r.tables[v];
    const bool bb = T->task == AVG_TASK_BED_BATHING;                                              // human_impairment='none', bed_bathing.py:198
    const int impairment = (bb || T->new_mode) ? 0 : (int)(reset_u32(sd, ue, ep, 1) & 3u);        // none, limits, weakness, tremor; `New`: 'none' (scratch_itch.py:159)
    const float limit_scale = impairment == 1 ? 0.5f + 0.5f * reset_u01(sd, ue, ep, 2) : 1.0f;    // world_creation.py:70
    const float strength = impairment == 2 ? 0.25f + 0.75f * reset_u01(sd, ue, ep, 3) : 1.0f;     // world_creation.py:71
    float* rec = r.env + (size_t)e * AVG_ENV_STRIDE;
    int* rec_i = reinterpret_cast<int*>(rec);
    for (int i = 0; i < AVG_ENV_STRIDE; ++i) rec[i] = 0.0f;
    const int k = (int)((pick / (uint32_t)npg) % (uint32_t)T->n_pool);
    for (int j = 0; j < T->n_arm; ++j) { rec[AVG_E_Q + T->arm_qidx[j]] = T->pool_q[k][j]; rec[AVG_E_MTARGET + T->arm_dof[j]] = T->pool_q[k][j]; }
    for (int j = 0; j < T->n_fin; ++j) { rec[AVG_E_Q + T->fin_qidx[j]] = T->fin_open; rec[AVG_E_MTARGET + T->fin_dof[j]] = T->fin_open; }
    for (int j = 0; j < T->n_hum; ++j) {
        const float q = fminf(fmaxf(T->hum_reset[j], T->hum_lower[j] * limit_scale), T->hum_upper[j] * limit_scale);   // world_creation.py:172
        rec[AVG_E_Q + T->hum_qidx[j]] = q; rec[AVG_E_MTARGET + T->hum_dof[j]] = q;
        rec[AVG_E_TARGET_H + T->hum_joint[j] - 4] = q;                                            // scratch_itch.py:235
    }
    for (int j = 0; j < 7; ++j) rec[AVG_E_Q + T->tool_qidx + j] = T->pool_tool[k][j];
    const float deg10 = 0.17453292519943295f;
    for (int j = 0; j < 10; ++j) rec[AVG_E_TREMOR + j] = impairment == 3 ? (2.0f * reset_u01(sd, ue, ep, 4 + j) - 1.0f) * deg10 : 0.0f;   // world_creation.py:141
    if (bb) {                                                                                     // every wiping target alive, bed_bathing.py:369-379
        for (int w = 0; w < 5; ++w) {
            const int nbit = min(max(T->n_target - 32 * w, 0), 32);
            reinterpret_cast<uint32_t*>(rec)[AVG_E_TARGET_MASK + w] = nbit == 32 ? 0xffffffffu : ((1u << nbit) - 1u);
        }
        rec_i[AVG_E_LIMB_FRAME] = AVG_F_SHOULDER;
    } else {
        const int limb = (int)(reset_u32(sd, ue, ep, 14) & 1u);                                   // scratch_itch.py:278
        const float length = T->limb_dims[limb][0], radius = T->limb_dims[limb][1];
        const float rl = radius + reset_u01(sd, ue, ep, 15) * (length - radius);                  // util.py:118
        const float th = 6.283185307179586f * reset_u01(sd, ue, ep, 16);
        float sn, cs; sincosf(th, &sn, &cs);
        rec[AVG_E_TARGET_ON_ARM + 0] = -radius * sn; rec[AVG_E_TARGET_ON_ARM + 1] = -radius * cs; rec[AVG_E_TARGET_ON_ARM + 2] = -rl;
        rec_i[AVG_E_LIMB_FRAME] = limb == 0 ? AVG_F_SHOULDER : AVG_F_ELBOW;
    }
    rec[AVG_E_STRENGTH] = strength; rec[AVG_E_LIMIT_SCALE] = limit_scale;
    rec[AVG_E_TREMOR_ON] = impairment == 3 ? 1.0f : 0.0f;
    rec[AVG_E_HUMAN_KP] = (T->human_control || impairment == 3) ? 0.05f : 0.01f;                 // scratch_itch.py:45 / :231
    if (T->new_mode && T->head_mask) reinterpret_cast<uint32_t*>(rec)[AVG_E_FROZEN] = T->head_mask;   // BedBathing New: the arm is static during play (bed_bathing.py:271)
    r.variant[e] = v;
    int* scr_i = reinterpret_cast<int*>(r.scratch + (size_t)e * AVG_S_STRIDE);
    scr_i[AVG_S_NSEP] = 0; scr_i[AVG_S_NC] = 0; scr_i[AVG_S_NCS] = 0; scr_i[AVG_S_NQ] = 0;    // no certificates / contacts carried over
}

// Start pose by inverse kinematics on the device (reference scratch_itch.py:243-253 -> util.ik_random_restarts, util.py:34-105:
// PyBullet's calculateInverseKinematics with random rest poses, accepted when the end effector is within 0.03 of the target in
// position and in quaternion distance, up to 40 restarts).  Restated as damped least squares over the 7 arm joints, one
// thread per environment: dq = J^T (J J^T + lambda^2 I)^-1 e with the step clamped to 0.3 rad and the joints to their limits.
// Runs after avg_reset_kernel for the environments it has just reset; the pool pose written there stays as the fallback.
namespace {
struct IkChain {
    int n;
    int body[8], qidx[8], dof[8];
    float lo[8], hi[8];
};
__device__ void ik_fk(const KM& m, const IkChain& c, const float* q, const AvgResetTable* T, V3* jp, V3* ja, V3& pe, Q4& qe, V3& pb, Q4& qb) {
    V3 p = mk3(0, 0, 0); Q4 r = mkq(0, 0, 0, 1);
    for (int j = 0; j < c.n; ++j) {
        const AvgBody* B = &m.body[c.body[j]];
        const Q4 jq0 = qmul(r, ldq(B->ta_quat));
        jp[j] = p + qrot(r, ld3(B->ta_pos));
        const V3 ax = ld3(B->axis);
        ja[j] = qrot(jq0, ax);
        const Q4 jq = qmul(jq0, qaxis(ax, q[j]));
        p = jp[j] + qrot(jq, ld3(B->tb_pos));
        r = qnormalize(qmul(jq, ldq(B->tb_quat)));
    }
    pb = p; qb = r;                                              // pose of the last arm body (carries the end effector)
    pe = p + qrot(r, ld3(T->ik_ee_frame));
    qe = qnormalize(qmul(r, ldq(T->ik_ee_frame + 3)));
}
// start target: centre + U(-range, range)^3 (scratch_itch.py:243,251), fixed orientation
__device__ __forceinline__ V3 ik_start_target(const AvgResetTable* T, uint32_t sd, uint32_t ue, uint32_t ep) {
    V3 tp = ld3(T->ik_target);
    tp.x += (2.0f * reset_u01(sd, ue, ep, 20) - 1.0f) * T->ik_range;
    tp.y += (2.0f * reset_u01(sd, ue, ep, 21) - 1.0f) * T->ik_range;
    tp.z += (2.0f * reset_u01(sd, ue, ep, 22) - 1.0f) * T->ik_range;
    if (T->has_bowl) {                                                               // Feeding: the target sits above the bowl drawn for this episode (feeding.py:276)
        tp.x += (2.0f * reset_u01(sd, ue, ep, 11) - 1.0f) * 0.05f; tp.y += (2.0f * reset_u01(sd, ue, ep, 12) - 1.0f) * 0.05f;
    }
    return tp;
}
__device__ __forceinline__ V3 ik_rot_err(Q4 target, Q4 cur) {
    Q4 d = qmul(target, qconj(cur));
    if (d.w < 0) d = mkq(-d.x, -d.y, -d.z, -d.w);
    const float sn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    if (sn < 1e-9f) return mk3(0, 0, 0);
    const float ang = 2.0f * atan2f(sn, d.w) / sn;
    return mk3(d.x * ang, d.y * ang, d.z * ang);
}
}  // namespace

__global__ void __launch_bounds__(64)
avg_reset_ik_kernel(AvgResetArgs r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n_env) return;
    if (r.mask && !r.mask[e]) return;
    const int v = r.variant[e];
    const AvgResetTable* T = r.tables[v];
    if (!T->ik_enabled || T->n_arm <= 0 || T->n_arm > 8) return;
    const KM m = open_model(r.models[v]);
    const uint32_t sd = r.seed, ue = (uint32_t)e, ep = (uint32_t)r.episode[e];
    float* rec = r.env + (size_t)e * AVG_ENV_STRIDE;
    const bool resolve = r.round == 0 || (r.retry && r.retry[e]);                    // else: back onto the accepted pose after the 5 test steps
    IkChain c; c.n = T->n_arm;
    for (int j = 0; j < c.n; ++j) {
        c.dof[j] = T->arm_dof[j]; c.qidx[j] = T->arm_qidx[j];
        const AvgDof* D = &m.dof[c.dof[j]];
        c.body[j] = D->body;
        const bool unlimited = D->lower > D->upper;                                  // continuous joint: +-2 pi for the IK, util.py:86-88
        c.lo[j] = unlimited ? -6.283185307f : D->lower; c.hi[j] = unlimited ? 6.283185307f : D->upper;
    }
    const V3 tp = ik_start_target(T, sd, ue, ep);
    const float tol = T->ik_tol > 0.0f ? T->ik_tol : 0.03f;
    const Q4 tq = ldq(T->ik_target + 3);
    float q[8], qbest[8]; float best_err = 3.0e38f;
    V3 jp[8], ja[8], pe, pb; Q4 qe, qb;
    const float lambda2 = 0.05f * 0.05f;
    if (!resolve) for (int j = 0; j < c.n; ++j) qbest[j] = rec[AVG_E_MTARGET + c.dof[j]];
    for (int restart = 0; resolve && restart < 40; ++restart) {                       // max_ik_random_restarts
        for (int j = 0; j < c.n; ++j) {                                               // random rest pose, util.py:100
            const float a = fmaxf(c.lo[j], -3.14159265f), b = fminf(c.hi[j], 3.14159265f);
            q[j] = a + (b - a) * reset_u01(sd, ue, ep, 32 + 8 * (restart + 40 * r.round) + j);
        }
        for (int it = 0; it < 200; ++it) {
            ik_fk(m, c, q, T, jp, ja, pe, qe, pb, qb);
            const V3 ev = tp - pe, ew = ik_rot_err(tq, qe);
            if (dot(ev, ev) < 1e-8f && dot(ew, ew) < 1e-6f) break;
            float J[6][8], A[6][6], y[6];
            const float err[6] = {ev.x, ev.y, ev.z, ew.x, ew.y, ew.z};
            for (int j = 0; j < c.n; ++j) {
                const V3 jl = cross(ja[j], pe - jp[j]);
                J[0][j] = jl.x; J[1][j] = jl.y; J[2][j] = jl.z; J[3][j] = ja[j].x; J[4][j] = ja[j].y; J[5][j] = ja[j].z;
            }
            for (int a = 0; a < 6; ++a) for (int b = 0; b <= a; ++b) {
                float sacc = a == b ? lambda2 : 0.0f;
                for (int j = 0; j < c.n; ++j) sacc = fmaf(J[a][j], J[b][j], sacc);
                A[a][b] = sacc;
            }
            // Cholesky A = L L^T in place (lower), then forward / backward substitution
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b <= a; ++b) {
                    float sacc = A[a][b];
                    for (int k = 0; k < b; ++k) sacc -= A[a][k] * A[b][k];
                    A[a][b] = a == b ? sqrtf(fmaxf(sacc, 1e-12f)) : sacc / A[b][b];
                }
            }
            for (int a = 0; a < 6; ++a) { float sacc = err[a]; for (int k = 0; k < a; ++k) sacc -= A[a][k] * y[k]; y[a] = sacc / A[a][a]; }
            for (int a = 5; a >= 0; --a) { float sacc = y[a]; for (int k = a + 1; k < 6; ++k) sacc -= A[k][a] * y[k]; y[a] = sacc / A[a][a]; }
            float dq[8], mx = 0.0f;
            for (int j = 0; j < c.n; ++j) {
                float sacc = 0.0f;
                for (int a = 0; a < 6; ++a) sacc = fmaf(J[a][j], y[a], sacc);
                dq[j] = sacc; mx = fmaxf(mx, fabsf(sacc));
            }
            const float sc = mx > 0.3f ? 0.3f / mx : 1.0f;
            for (int j = 0; j < c.n; ++j) q[j] = fminf(fmaxf(q[j] + sc * dq[j], c.lo[j]), c.hi[j]);
        }
        ik_fk(m, c, q, T, jp, ja, pe, qe, pb, qb);
        const float epos = norm(tp - pe);
        const float dm = sqrtf((tq.x - qe.x) * (tq.x - qe.x) + (tq.y - qe.y) * (tq.y - qe.y) + (tq.z - qe.z) * (tq.z - qe.z) + (tq.w - qe.w) * (tq.w - qe.w));
        const float dp = sqrtf((tq.x + qe.x) * (tq.x + qe.x) + (tq.y + qe.y) * (tq.y + qe.y) + (tq.z + qe.z) * (tq.z + qe.z) + (tq.w + qe.w) * (tq.w + qe.w));
        const float eq = fminf(dm, dp);
        const bool ok = epos < tol && eq < tol;                                       // random_restart_threshold, util.py:51
        if (ok || epos < best_err) { best_err = epos; for (int j = 0; j < c.n; ++j) qbest[j] = q[j]; }   // util.py:53-55: else the closest attempt
        if (ok) break;
    }
    ik_fk(m, c, qbest, T, jp, ja, pe, qe, pb, qb);
    if (r.round > 0) {                                                                // after test steps: every joint back on its reset pose, at rest
        for (int d = 0; d < m.h->n_jdof; ++d) rec[AVG_E_Q + m.body[m.dof[d].body].qidx] = rec[AVG_E_MTARGET + d];
        for (int d = 0; d < 32; ++d) rec[AVG_E_QD + d] = 0.0f;
        for (int d = 0; d < 2 * AVG_WCACHE_N; ++d) rec[AVG_E_WCACHE + d] = 0.0f;      // no impulses carried over from the test steps
    }
    for (int j = 0; j < c.n; ++j) { rec[AVG_E_Q + c.qidx[j]] = qbest[j]; rec[AVG_E_MTARGET + c.dof[j]] = qbest[j]; }
    // the tool goes where init_tool puts it (world_creation.py:331-337): weld-parent frame o inverse of the tool's base frame
    {
        const AvgFrame* FW = &m.frame[AVG_F_WELD_PARENT]; const AvgFrame* FT = &m.frame[AVG_F_TOOL_BASE];
        const V3 wp = pb + qrot(qb, ld3(FW->pos)); const Q4 wq = qnormalize(qmul(qb, ldq(FW->quat)));
        const Q4 tbq = qnormalize(qmul(wq, qconj(ldq(FT->quat))));
        const V3 tbp = wp - qrot(tbq, ld3(FT->pos));
        float* tq7 = rec + AVG_E_Q + T->tool_qidx;
        tq7[0] = tbp.x; tq7[1] = tbp.y; tq7[2] = tbp.z; tq7[3] = tbq.x; tq7[4] = tbq.y; tq7[5] = tbq.z; tq7[6] = tbq.w;
    }
    // inspection slots (last env-static pose slot, unused by every task): drawn start target and the position error reached
    rec[AVG_E_EBODY + 21] = tp.x; rec[AVG_E_EBODY + 22] = tp.y; rec[AVG_E_EBODY + 23] = tp.z; rec[AVG_E_EBODY + 24] = norm(tp - pe);
}

// Episode reset of Feeding / Drinking on the device: the random draws of FeedingEnv.reset / DrinkingEnv.reset (feeding.py:171-185,
// 242-245, drinking.py:185-201, 241-243) restated as in compiler/reset_fd.py sample_states_fd(), one thread per environment.
// The start pose comes from avg_reset_ik_kernel (fresh target above the bowl, feeding.py:276-278), the particles are placed by
// avg_reset_particles_kernel once the tool sits in the gripper (feeding.py:291-307), and the caller then runs the reference's
// 100 settle steps (feeding.py:318-320) with avg_launch_settle.
__global__ void __launch_bounds__(128)
avg_reset_fd_kernel(AvgResetArgs r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n_env) return;
    if (r.mask && !r.mask[e]) return;
    const uint32_t ep = (uint32_t)(r.episode[e] + 1);
    r.episode[e] = (int32_t)ep;
    const uint32_t sd = r.seed, ue = (uint32_t)e;
    const uint32_t pick = reset_u32(sd, ue, ep, 17);
    const int npg = r.n_per_gender;
    const int v = (int)(reset_u32(sd, ue, ep, 0) % (uint32_t)min(r.n_variants, 2)) * npg + (int)(pick % (uint32_t)npg);   // gender, feeding.py:172
    const AvgResetTable* T = r.tables[v];
    const int impairment = T->new_mode ? 0 : (int)(reset_u32(sd, ue, ep, 1) & 3u);                 // none, limits, weakness, tremor (world_creation.py:67); `New`: 'none' (feeding.py:172)
    const float limit_scale = impairment == 1 ? 0.5f + 0.5f * reset_u01(sd, ue, ep, 2) : 1.0f;
    const float strength = impairment == 2 ? 0.25f + 0.75f * reset_u01(sd, ue, ep, 3) : 1.0f;
    float* rec = r.env + (size_t)e * AVG_ENV_STRIDE;
    for (int i = 0; i < AVG_ENV_STRIDE; ++i) rec[i] = 0.0f;
    const int k = (int)((pick / (uint32_t)npg) % (uint32_t)T->n_pool);
    for (int j = 0; j < T->n_arm; ++j) { rec[AVG_E_Q + T->arm_qidx[j]] = T->pool_q[k][j]; rec[AVG_E_MTARGET + T->arm_dof[j]] = T->pool_q[k][j]; }
    for (int j = 0; j < T->n_fin; ++j) { rec[AVG_E_Q + T->fin_qidx[j]] = T->fin_q[j]; rec[AVG_E_MTARGET + T->fin_dof[j]] = T->fin_q[j]; }
    const float deg30 = 0.5235987755982988f, deg20 = 0.3490658503988659f;
    for (int j = 0; j < T->n_hum; ++j) {
        float q = T->hum_reset[j];
        const int joint = T->hum_joint[j];
        if (joint >= 25 && joint <= 27) q = (2.0f * reset_u01(sd, ue, ep, 8 + joint - 25) - 1.0f) * deg30;              // feeding.py:243
        q = fminf(fmaxf(q, T->hum_lower[j] * limit_scale), T->hum_upper[j] * limit_scale);                              // world_creation.py:172
        rec[AVG_E_Q + T->hum_qidx[j]] = q; rec[AVG_E_MTARGET + T->hum_dof[j]] = q;
        rec[AVG_E_TARGET_H + T->hum_slot[j]] = q;                                                                       // feeding.py:248
    }
    for (int j = 0; j < 7; ++j) rec[AVG_E_Q + T->tool_qidx + j] = T->pool_tool[k][j];
    for (int j = 0; j < 4; ++j) rec[AVG_E_TREMOR + j] = impairment == 3 ? (2.0f * reset_u01(sd, ue, ep, 4 + j) - 1.0f) * deg20 : 0.0f;   // world_creation.py:138-139
    if (T->has_bowl) {                                                                                                  // feeding.py:184-185
        rec[AVG_E_EBODY + 0] = T->bowl_center[0] + (2.0f * reset_u01(sd, ue, ep, 11) - 1.0f) * 0.05f;
        rec[AVG_E_EBODY + 1] = T->bowl_center[1] + (2.0f * reset_u01(sd, ue, ep, 12) - 1.0f) * 0.05f;
        rec[AVG_E_EBODY + 2] = T->bowl_center[2];
        for (int j = 0; j < 4; ++j) rec[AVG_E_EBODY + 3 + j] = T->bowl_quat[j];
    }
    rec[AVG_E_STRENGTH] = strength; rec[AVG_E_LIMIT_SCALE] = limit_scale;
    rec[AVG_E_TREMOR_ON] = impairment == 3 ? 1.0f : 0.0f;
    rec[AVG_E_HUMAN_KP] = 0.005f;                                                                  // human_gains, feeding.py:48
    reinterpret_cast<uint32_t*>(rec)[AVG_E_FROZEN] = (T->human_control || impairment == 3) ? 0u : T->head_mask;   // feeding.py:244
    r.variant[e] = v;
    int* scr_i = reinterpret_cast<int*>(r.scratch + (size_t)e * AVG_S_STRIDE);
    scr_i[AVG_S_NSEP] = 0; scr_i[AVG_S_NC] = 0; scr_i[AVG_S_NCS] = 0; scr_i[AVG_S_NQ] = 0;
}

// feeding.py:291-307 / drinking.py:291-312: the particle grid above the tool, at rest; every particle alive
__global__ void __launch_bounds__(128)
avg_reset_particles_kernel(AvgResetArgs r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n_env || !r.part) return;
    if (r.mask && !r.mask[e]) return;
    const AvgResetTable* T = r.tables[r.variant[e]];
    const float* rec = r.env + (size_t)e * AVG_ENV_STRIDE;
    float* pr = r.part + (size_t)e * AVG_P_STRIDE;
    for (int i = 0; i < AVG_P_STRIDE; ++i) pr[i] = 0.0f;
    const float* tp = rec + AVG_E_Q + T->tool_qidx;
    const int np = T->n_particle;
    for (int p = 0; p < np; ++p) for (int c = 0; c < 3; ++c) pr[AVG_P_POS + 64 * c + p] = tp[c] + T->grid[p][c];
    uint32_t* pu = reinterpret_cast<uint32_t*>(pr);
    pu[AVG_P_ALIVE] = np >= 32 ? 0xffffffffu : ((1u << np) - 1u);
    pu[AVG_P_ALIVE + 1] = np > 32 ? (np >= 64 ? 0xffffffffu : ((1u << (np - 32)) - 1u)) : 0u;
}

// util.ik_random_restarts(step_sim=True), util.py:41-46,51: after 5 stepSimulation calls from the solved pose, the pose is kept only
// if the robot does not touch itself (getContactPoints(body, body)) and the arm is still where the IK put it (the reference measures
// the end effector after the steps against random_restart_threshold; a pose in contact with the table or the chair is pushed away).
__global__ void __launch_bounds__(128)
avg_reset_check_kernel(AvgResetArgs r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n_env || !r.retry) return;
    if (r.mask && !r.mask[e]) { r.retry[e] = 0; return; }
    const int v = r.variant[e];
    const AvgResetTable* T = r.tables[v];
    if (!T->ik_enabled) { r.retry[e] = 0; return; }
    const KM m = open_model(r.models[v]);
    const float* rec = r.env + (size_t)e * AVG_ENV_STRIDE;
    const float* scr = r.scratch + (size_t)e * AVG_S_STRIDE;
    const int nc = reinterpret_cast<const int*>(scr)[AVG_S_NCS];
    bool bad = false;
    for (int c = 0; c < nc && c < AVG_MAX_CONTACT; ++c) {
        const float* g = scr + AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * c;
        const int sa = __float_as_int(g[10]), sb = __float_as_int(g[11]);
        if (m.shape[sa].ref_body == AVG_REF_ROBOT && m.shape[sb].ref_body == AVG_REF_ROBOT) bad = true;
    }
    // the reference's acceptance test runs AFTER the 5 steps (util.py:50-52): end effector within random_restart_threshold of the
    // target in position and in quaternion distance -- a pose that something (the table, the chair) pushed away fails it
    if (T->n_arm > 0 && T->n_arm <= 8) {
        IkChain c; c.n = T->n_arm;
        float q[8];
        for (int j = 0; j < c.n; ++j) { c.dof[j] = T->arm_dof[j]; c.qidx[j] = T->arm_qidx[j]; c.body[j] = m.dof[c.dof[j]].body; q[j] = rec[AVG_E_Q + c.qidx[j]]; }
        V3 jp[8], ja[8], pe, pb; Q4 qe, qb;
        ik_fk(m, c, q, T, jp, ja, pe, qe, pb, qb);
        const V3 tp = ik_start_target(T, r.seed, (uint32_t)e, (uint32_t)r.episode[e]);
        const Q4 tq = ldq(T->ik_target + 3);
        const float tol = T->ik_tol > 0.0f ? T->ik_tol : 0.03f;
        const float dm = sqrtf((tq.x - qe.x) * (tq.x - qe.x) + (tq.y - qe.y) * (tq.y - qe.y) + (tq.z - qe.z) * (tq.z - qe.z) + (tq.w - qe.w) * (tq.w - qe.w));
        const float dp = sqrtf((tq.x + qe.x) * (tq.x + qe.x) + (tq.y + qe.y) * (tq.y + qe.y) + (tq.z + qe.z) * (tq.z + qe.z) + (tq.w + qe.w) * (tq.w + qe.w));
        if (!(norm(tp - pe) < tol && fminf(dm, dp) < tol)) bad = true;
    }
    r.retry[e] = bad ? 1 : 0;
}

// `New` ids of ScratchItch / BedBathing (scratch_itch.py:198-228, bed_bathing.py:256-279): the dynamic arm joints start at their
// preset plus U(-10, 10) degrees, redrawn until the arm keeps new_min_dist (0.01) from the rest of the person (links other than
// the arm itself, the chest 3 and the shoulder 6), from the robot and from the furniture.  The reference measures those
// distances with getClosestPoints; here every shape is represented by its bounding capsule (exact for the capsules and spheres
// the person is made of, conservative for hulls; static boxes use their face axes: a pose may be redrawn that the reference would
// have kept, never the reverse).
// One thread per environment, after the start pose of the robot is known (avg_reset_ik_kernel); at most 20 draws, the last kept.
namespace {
__device__ __forceinline__ float capsule_gap(V3 a0, V3 a1, float ra, V3 b0, V3 b1, float rb) {
    return seg_seg_distance(a0, a1 - a0, b0, b1 - b0) - ra - rb;
}
}  // namespace

__global__ void __launch_bounds__(64)
avg_reset_new_kernel(AvgResetArgs r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n_env) return;
    if (r.mask && !r.mask[e]) return;
    const int v = r.variant[e];
    const AvgResetTable* T = r.tables[v];
    if (!T->new_mode || !(T->hum_jitter > 0.0f) || T->n_hum <= 0 || T->n_hum > 8) return;
    const KM m = open_model(r.models[v]);
    const AvgModelHeader* h = m.h;
    const uint32_t sd = r.seed, ue = (uint32_t)e, ep = (uint32_t)r.episode[e];
    float* rec = r.env + (size_t)e * AVG_ENV_STRIDE;
    const int nb = h->n_body, nms = h->n_mshape, ns = h->n_shape;
    // world poses of the dynamic bodies (serial: parents precede children)
    V3 bp[AVG_MAX_BODY]; Q4 bq[AVG_MAX_BODY];
    auto fk_body = [&](int b) {
        const AvgBody* B = &m.body[b];
        if (B->jtype == AVG_JOINT_FREE) { const float* qq = rec + AVG_E_Q + B->qidx; bp[b] = ld3(qq); bq[b] = qnormalize(ldq(qq + 3)); return; }
        V3 pp = mk3(0, 0, 0); Q4 pq = mkq(0, 0, 0, 1);
        if (B->parent >= 0) { pp = bp[B->parent]; pq = bq[B->parent]; }
        const V3 ax = ld3(B->axis);
        const float qv = rec[AVG_E_Q + B->qidx];
        Q4 jq = ldq(B->ta_quat); V3 jp = ld3(B->ta_pos);
        if (B->jtype == AVG_JOINT_REVOLUTE) jq = qmul(jq, qaxis(ax, qv)); else jp = jp + qrot(jq, ax * qv);
        bp[b] = pp + qrot(pq, jp + qrot(jq, ld3(B->tb_pos)));
        bq[b] = qnormalize(qmul(pq, qmul(jq, ldq(B->tb_quat))));
    };
    for (int b = 0; b < nb; ++b) fk_body(b);
    auto world_capsule = [&](int si, V3& c0, V3& c1, float& rad) {
        const float4 k0 = __ldg(&m.bcap[2 * si]), k1 = __ldg(&m.bcap[2 * si + 1]);
        rad = k0.w;
        if (si < nms) {
            const AvgShape* S = &m.shape[si];
            const Q4 sq = qnormalize(qmul(bq[S->body], ldq(S->quat)));
            const V3 sp = bp[S->body] + qrot(bq[S->body], ld3(S->pos));
            c0 = sp + qrot(sq, mk3(k0.x, k0.y, k0.z)); c1 = sp + qrot(sq, mk3(k1.x, k1.y, k1.z));
        } else { c0 = mk3(k0.x, k0.y, k0.z); c1 = mk3(k1.x, k1.y, k1.z); }
    };
    // the arm's shapes: the moving shapes of the person
    int arm[4], narm = 0;
    for (int si = 0; si < nms && narm < 4; ++si) if (m.shape[si].ref_body == AVG_REF_HUMAN && m.shape[si].body >= 0 && m.shape[si].body < nb) arm[narm++] = si;
    float qbest[8]; float gap_best = -3.0e38f; int attempts = 0;
    for (int at = 0; at < 20; ++at) {
        attempts = at + 1;
        float q[8];
        for (int j = 0; j < T->n_hum; ++j) {
            const float u = 2.0f * reset_u01(sd, ue, ep, 2048u + 8u * (uint32_t)at + (uint32_t)j) - 1.0f;
            q[j] = fminf(fmaxf(T->hum_reset[j] + u * T->hum_jitter, T->hum_lower[j]), T->hum_upper[j]);      // add_joint_positions + enforce_joint_limits
            rec[AVG_E_Q + T->hum_qidx[j]] = q[j];
        }
        for (int j = 0; j < T->n_hum; ++j) fk_body(m.dof[T->hum_dof[j]].body);       // the arm chain is listed root first
        float gap = 3.0e38f;
        for (int k = 0; k < narm; ++k) {
            V3 a0, a1; float ra; world_capsule(arm[k], a0, a1, ra);
            for (int sj = 0; sj < ns; ++sj) {
                const AvgShape* S = &m.shape[sj];
                if (S->type == AVG_SHAPE_PLANE || (sj < nms && (S->body < 0 || S->body >= nb))) continue;      // env-static bodies: none in these tasks
                const int rb_ = S->ref_body;
                bool other = false;
                if (rb_ == AVG_REF_HUMAN) other = sj >= nms && S->ref_link != 3 && S->ref_link != 6;       // scratch_itch.py:219
                else if (rb_ == AVG_REF_ROBOT || rb_ == AVG_REF_FURNITURE) other = true;                    // :221-223
                if (!other) continue;
                if (S->type == AVG_SHAPE_BOX && sj >= nms) {
                    // a flat box (the mattress) has a hopeless bounding capsule: lower bound from its face axes instead -- the
                    // largest separation of the arm's segment from a pair of faces (exact when the closest point lies on a face)
                    const Q4 bq_ = ldq(S->quat); const V3 bc = ld3(S->pos);
                    const V3 l0 = qrot_inv(bq_, a0 - bc), l1 = qrot_inv(bq_, a1 - bc);
                    const float p0[3] = {l0.x, l0.y, l0.z}, p1[3] = {l1.x, l1.y, l1.z};
                    float sep = -3.0e38f;
                    for (int ax = 0; ax < 3; ++ax) {
                        const float sk = (p0[ax] > 0.0f) == (p1[ax] > 0.0f) ? fminf(fabsf(p0[ax]), fabsf(p1[ax])) - S->half[ax] : -S->half[ax];
                        sep = fmaxf(sep, sk);
                    }
                    gap = fminf(gap, sep - ra);
                    continue;
                }
                V3 b0, b1; float rbb; world_capsule(sj, b0, b1, rbb);
                gap = fminf(gap, capsule_gap(a0, a1, ra, b0, b1, rbb));
            }
        }
        if (gap > gap_best) { gap_best = gap; for (int j = 0; j < T->n_hum; ++j) qbest[j] = q[j]; }
        if (gap >= T->new_min_dist) break;
    }
    for (int j = 0; j < T->n_hum; ++j) {
        rec[AVG_E_Q + T->hum_qidx[j]] = qbest[j]; rec[AVG_E_MTARGET + T->hum_dof[j]] = qbest[j];
        rec[AVG_E_TARGET_H + T->hum_slot[j]] = qbest[j];                                                      // scratch_itch.py:235
    }
    // inspection slots (third env-static pose slot, unused by these tasks): bounding-capsule clearance of the kept pose, draws used
    rec[AVG_E_EBODY + 14] = gap_best; rec[AVG_E_EBODY + 15] = (float)attempts;
}

cudaError_t avg_launch_reset_check(const AvgResetArgs& r, cudaStream_t stream) {
    avg_reset_check_kernel<<<(r.n_env + 127) / 128, 128, 0, stream>>>(r);
    avg_reset_ik_kernel<<<(r.n_env + 63) / 64, 64, 0, stream>>>(r);
    if (r.part) avg_reset_particles_kernel<<<(r.n_env + 127) / 128, 128, 0, stream>>>(r);
    return cudaGetLastError();
}

cudaError_t avg_launch_reset(const AvgResetArgs& r, cudaStream_t stream) {
    if (r.part) avg_reset_fd_kernel<<<(r.n_env + 127) / 128, 128, 0, stream>>>(r);
    else avg_reset_kernel<<<(r.n_env + 127) / 128, 128, 0, stream>>>(r);
    if (r.any_ik) avg_reset_ik_kernel<<<(r.n_env + 63) / 64, 64, 0, stream>>>(r);
    if (r.any_new) avg_reset_new_kernel<<<(r.n_env + 63) / 64, 64, 0, stream>>>(r);
    if (r.part) avg_reset_particles_kernel<<<(r.n_env + 127) / 128, 128, 0, stream>>>(r);
    return cudaGetLastError();
}

// Policy inference (reference enjoy_vr.py:106-113 with the default a2c_ppo_acktr MLP actor): one warp per environment.
// Lane u owns hidden units u and u + 32; inputs and activations travel by shuffle; weight rows are read coalesced and
// stay L1/L2-resident (the whole policy is ~27 KB).
__global__ void __launch_bounds__(128)
avg_policy_kernel(const unsigned char* __restrict__ blob, const float* __restrict__ obs, float* __restrict__ actions, int n_env, int n_obs, int n_act) {
    const int lane = threadIdx.x & 31, e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n_env) return;
    const AvgPolicyHeader* ph = reinterpret_cast<const AvgPolicyHeader*>(blob);
    const int n_in = ph->n_in, n_out = ph->n_out;
    const float* mean = reinterpret_cast<const float*>(blob + sizeof(AvgPolicyHeader));
    const float* var = mean + n_in;
    const float* W1 = var + n_in; const float* b1 = W1 + n_in * 64;
    const float* W2 = b1 + 64; const float* b2 = W2 + 4096;
    const float* W3 = b2 + 64; const float* b3 = W3 + 64 * n_out;
    // VecNormalize: clip((obs - mean) / sqrt(var + eps), -clip, clip)
    float x0 = 0.0f, x1 = 0.0f;
    if (lane < n_in) x0 = fminf(fmaxf((obs[(size_t)e * n_obs + lane] - __ldg(mean + lane)) * rsqrtf(__ldg(var + lane) + ph->eps), -ph->clip_obs), ph->clip_obs);
    if (lane + 32 < n_in) x1 = fminf(fmaxf((obs[(size_t)e * n_obs + lane + 32] - __ldg(mean + lane + 32)) * rsqrtf(__ldg(var + lane + 32) + ph->eps), -ph->clip_obs), ph->clip_obs);
    float a0 = __ldg(b1 + lane), a1 = __ldg(b1 + lane + 32);
    for (int k = 0; k < n_in; ++k) {
        const float xk = __shfl_sync(AVG_FULL, k < 32 ? x0 : x1, k & 31);
        a0 = fmaf(xk, __ldg(W1 + k * 64 + lane), a0); a1 = fmaf(xk, __ldg(W1 + k * 64 + lane + 32), a1);
    }
    float h0 = tanhf(a0), h1 = tanhf(a1);
    a0 = __ldg(b2 + lane); a1 = __ldg(b2 + lane + 32);
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const float u = __shfl_sync(AVG_FULL, h0, j), v = __shfl_sync(AVG_FULL, h1, j);
        a0 = fmaf(u, __ldg(W2 + j * 64 + lane), a0); a1 = fmaf(u, __ldg(W2 + j * 64 + lane + 32), a1);
        a0 = fmaf(v, __ldg(W2 + (j + 32) * 64 + lane), a0); a1 = fmaf(v, __ldg(W2 + (j + 32) * 64 + lane + 32), a1);
    }
    h0 = tanhf(a0); h1 = tanhf(a1);
    float out = lane < n_out ? __ldg(b3 + lane) : 0.0f;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const float u = __shfl_sync(AVG_FULL, h0, j), v = __shfl_sync(AVG_FULL, h1, j);
        if (lane < n_out) out = fmaf(v, __ldg(W3 + (j + 32) * n_out + lane), fmaf(u, __ldg(W3 + j * n_out + lane), out));
    }
    if (lane < n_act) actions[(size_t)e * n_act + lane] = lane < n_out ? out : 0.0f;        // enjoy_vr.py:112-113: the human half is zero
}

cudaError_t avg_launch_policy(const unsigned char* policy_blob, const float* obs, float* actions, int n_env, int n_obs, int n_act, cudaStream_t stream) {
    avg_policy_kernel<<<(n_env + 3) / 4, 128, 0, stream>>>(policy_blob, obs, actions, n_env, n_obs, n_act);
    return cudaGetLastError();
}

// parity tap for the arm-limit classifier: raw joint angles (tz, tx, ty, qe) of joints 7..10 -> logit, one warp per pose
__global__ void __launch_bounds__(128)
avg_arm_limit_kernel(const unsigned char* blob, const float* __restrict__ q4, float* __restrict__ logits, int n) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const KM m = open_model(blob);
    if (!m.mlp) { if (lane == 0) logits[i] = 0.0f; return; }
    float x[4];
    arm_limit_inputs(q4[4 * i], q4[4 * i + 1], q4[4 * i + 2], q4[4 * i + 3], x);
    const float lg = arm_limit_logit_warp(m.mlp, x, lane);
    if (lane == 0) logits[i] = lg;
}

// =================================================================================================================
// host-side launch helpers
// =================================================================================================================
template <class K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// warps (= environments) per block, per kernel.  The solver uses one warp per block: its iteration count varies per
// environment (residual early exit), and a block holds its shared memory until its slowest warp is done.
constexpr int kWpbCollide = AVG_WPB_COLLIDE, kWpbDyn = AVG_WPB_DYN, kWpbSolve = 1, kWpbEpi = 4, kWpbPro = 4, kWpbPCol = 4;

// AVG_KERNEL_TIMES=1 in the environment: every launch of the step is bracketed by CUDA events and the per-kernel totals
// are printed to stderr every 8 steps (a development aid; it serialises the host with the stream, so never set it for
// a measurement of the step itself).
namespace {
struct KernelTimes {
    bool on = false, init = false;
    cudaEvent_t ev[128];
    double ms[7] = {0, 0, 0, 0, 0, 0, 0};
    int steps = 0;
};
KernelTimes g_kt;
bool g_configured[64];                        // per device ordinal: the shared-memory opt-ins are per device / context

cudaError_t configure_kernels() {
    int dev = 0;
    cudaError_t e1 = cudaGetDevice(&dev);
    if (e1 != cudaSuccess) return e1;
    if (dev >= 0 && dev < 64 && g_configured[dev]) return cudaSuccess;
    const size_t sm_col = sizeof(SmCollide) * kWpbCollide, sm_dyn = sizeof(SmDyn) * kWpbDyn;
    const size_t sm_sol = sizeof(SmSolve) * kWpbSolve, sm_epi = sizeof(SmEpi) * kWpbEpi;
    const size_t sm_sol1 = sizeof(SmSolve) + sizeof(SmPartSmall), sm_sol2 = sizeof(SmSolve) + sizeof(SmPartLarge);
    if ((e1 = set_smem(avg_collide_kernel, sm_col)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_collide_large_kernel, sm_col)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_pcollide_kernel, sizeof(SmPCol) * kWpbPCol)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_dynamics_kernel<8>, sm_dyn)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_dynamics_kernel<10>, sm_dyn)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_dynamics_kernel<12>, sm_dyn)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_dynamics_kernel<16>, sm_dyn)) != cudaSuccess) return e1;
    const size_t sm_fus = sizeof(SmDynSolve) * AVG_WPB_FUSED;
    if ((e1 = set_smem(avg_dynsolve_kernel<8>, sm_fus)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_dynsolve_kernel<10>, sm_fus)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_dynsolve_kernel<12>, sm_fus)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_dynsolve_kernel<16>, sm_fus)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<8, 0>, sm_sol)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<10, 0>, sm_sol)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<12, 0>, sm_sol)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<16, 0>, sm_sol)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<10, 1>, sm_sol1)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<12, 1>, sm_sol1)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<16, 1>, sm_sol1)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<10, 2>, sm_sol2)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<12, 2>, sm_sol2)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_solve_kernel<16, 2>, sm_sol2)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_epilogue_kernel, sm_epi)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_epilogue_bb_kernel, sizeof(SmEpiBB) * kWpbEpi)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_epilogue_fd_kernel, sm_epi)) != cudaSuccess) return e1;
    if ((e1 = set_smem(avg_reset_obs_kernel, sm_epi)) != cudaSuccess) return e1;
    if (dev >= 0 && dev < 64) g_configured[dev] = true;
    return cudaSuccess;
}

// the kernels of ONE internal physics step (collide [, particle broadphase], narrowphase, dynamics, solve)
template <class MARK>
void launch_internal_step(const AvgStepArgs& a, cudaStream_t stream, MARK&& mark) {
    const size_t sm_col = sizeof(SmCollide) * kWpbCollide, sm_dyn = sizeof(SmDyn) * kWpbDyn;
    const size_t sm_sol = sizeof(SmSolve) * kWpbSolve;
    const size_t sm_sol1 = sizeof(SmSolve) + sizeof(SmPartSmall), sm_sol2 = sizeof(SmSolve) + sizeof(SmPartLarge);
    const int n_range = a.env_end - a.env_begin;
    auto grid = [&](int wpb) { return (n_range + wpb - 1) / wpb; };
    const int np_grid = 148 * 16;                                       // grid-stride over the queue; a short queue is spread over all of these warps (ipw)
    const bool part = a.part != nullptr;
    if (a.n_env >= AVG_COLLIDE_LARGE) avg_collide_large_kernel<<<grid(kWpbCollide), 32 * kWpbCollide, sm_col, stream>>>(a);
    else avg_collide_kernel<<<grid(kWpbCollide), 32 * kWpbCollide, sm_col, stream>>>(a);
    mark(1);
    if (part) { avg_pcollide_kernel<<<grid(kWpbPCol), 32 * kWpbPCol, sizeof(SmPCol) * kWpbPCol, stream>>>(a); mark(6); }
    if (n_range > 16384) avg_narrow_kernel<AVG_OCC_NARROW><<<np_grid, 128, 0, stream>>>(a);
    else avg_narrow_kernel<3><<<np_grid, 128, 0, stream>>>(a);
    mark(5);
    static const int fuse_mode = [] { const char* c = getenv("AVG_FUSE"); return c ? atoi(c) : AVG_FUSE_DEFAULT; }();   // 0 never, 1 always, 2 small batches
    if (!part && (fuse_mode == 1 || (fuse_mode == 2 && n_range <= AVG_FUSE_MAX))) {
        const size_t sm_fus = sizeof(SmDynSolve) * AVG_WPB_FUSED;
        const int g = grid(AVG_WPB_FUSED);
        if (a.maxblk <= 8) avg_dynsolve_kernel<8><<<g, 32 * AVG_WPB_FUSED, sm_fus, stream>>>(a);
        else if (a.maxblk <= 10) avg_dynsolve_kernel<10><<<g, 32 * AVG_WPB_FUSED, sm_fus, stream>>>(a);
        else if (a.maxblk <= 12) avg_dynsolve_kernel<12><<<g, 32 * AVG_WPB_FUSED, sm_fus, stream>>>(a);
        else avg_dynsolve_kernel<16><<<g, 32 * AVG_WPB_FUSED, sm_fus, stream>>>(a);
        mark(3);
        return;
    }
    if (a.maxblk <= 8) avg_dynamics_kernel<8><<<grid(kWpbDyn), 32 * kWpbDyn, sm_dyn, stream>>>(a);
    else if (a.maxblk <= 10) avg_dynamics_kernel<10><<<grid(kWpbDyn), 32 * kWpbDyn, sm_dyn, stream>>>(a);
    else if (a.maxblk <= 12) avg_dynamics_kernel<12><<<grid(kWpbDyn), 32 * kWpbDyn, sm_dyn, stream>>>(a);
    else avg_dynamics_kernel<16><<<grid(kWpbDyn), 32 * kWpbDyn, sm_dyn, stream>>>(a);
    mark(2);
    if (part && a.task == AVG_TASK_FEEDING) {
        if (a.maxblk <= 10) avg_solve_kernel<10, 1><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol1, stream>>>(a);
        else if (a.maxblk <= 12) avg_solve_kernel<12, 1><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol1, stream>>>(a);
        else avg_solve_kernel<16, 1><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol1, stream>>>(a);
    } else if (part) {
        if (a.maxblk <= 10) avg_solve_kernel<10, 2><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol2, stream>>>(a);
        else if (a.maxblk <= 12) avg_solve_kernel<12, 2><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol2, stream>>>(a);
        else avg_solve_kernel<16, 2><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol2, stream>>>(a);
    }
    else if (a.maxblk <= 8) avg_solve_kernel<8, 0><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol, stream>>>(a);
    else if (a.maxblk <= 10) avg_solve_kernel<10, 0><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol, stream>>>(a);
    else if (a.maxblk <= 12) avg_solve_kernel<12, 0><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol, stream>>>(a);
    else avg_solve_kernel<16, 0><<<grid(kWpbSolve), 32 * kWpbSolve, sm_sol, stream>>>(a);
    mark(3);
}
}  // namespace

cudaError_t avg_launch_step(const AvgStepArgs& a_in, int substeps, cudaStream_t stream, int* launched) {
    cudaError_t ec = configure_kernels();
    if (ec != cudaSuccess) return ec;
    static bool env_read = false;
    const int n_internal = a_in.n_internal > 0 ? a_in.n_internal : 1;
    if (!env_read) {
        const char* kt = getenv("AVG_KERNEL_TIMES");
        g_kt.on = kt && kt[0] == '1' && 5 * substeps * n_internal + 3 <= 128;
        env_read = true;
    }
    AvgStepArgs a = a_in;
    const size_t sm_epi = sizeof(SmEpi) * kWpbEpi;
    const bool kt_on = g_kt.on && a.env_begin == 0 && a.env_end == a.n_env;     // whole-batch launches only
    int nev = 0; int kinds[128];
    if (kt_on && !g_kt.init) { for (int i = 0; i < 128; ++i) cudaEventCreate(&g_kt.ev[i]); g_kt.init = true; }
    static const bool sync_each = getenv("AVG_SYNC_EACH") != nullptr;        // development aid: name the kernel that faults
    int kseq = 0, nlaunch = 0;
    auto mark = [&](int kind) {
        if (kind >= 0) nlaunch++;
        if (kt_on && nev < 128) { kinds[nev] = kind; cudaEventRecord(g_kt.ev[nev++], stream); }
        if (sync_each) {
            const cudaError_t e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) fprintf(stderr, "[avg] launch #%d of the step (kind %d: 0 prologue, 1 collide, 6 particle broadphase, 5 narrowphase, 2 dynamics, 3 solve, 4 epilogue) failed: %s\n", kseq, kind, cudaGetErrorString(e));
            kseq++;
        }
    };
    const int n_range = a.env_end - a.env_begin;
    if (n_range <= 0) return cudaSuccess;
    auto grid = [&](int wpb) { return (n_range + wpb - 1) / wpb; };
    mark(-1);
    if (a.phase != 2) {                              // phase 2: the epilogue alone (avg_step_host overlaps result copies with it, range by range)
    avg_prologue_kernel<<<grid(kWpbPro), 32 * kWpbPro, 0, stream>>>(a);
    mark(0);
    for (int f = 0; f < substeps; ++f)
        for (int i = 0; i < n_internal; ++i) {
            a.post = (i == n_internal - 1) ? 1 : 0;
            launch_internal_step(a, stream, mark);
        }
    }
    if (a.phase == 1) { if (launched) *launched = nlaunch; return cudaGetLastError(); }     // phase 1: everything but the epilogue
    if (a.task == AVG_TASK_BED_BATHING) avg_epilogue_bb_kernel<<<grid(kWpbEpi), 32 * kWpbEpi, sizeof(SmEpiBB) * kWpbEpi, stream>>>(a);
    else if (a.task == AVG_TASK_FEEDING || a.task == AVG_TASK_DRINKING) avg_epilogue_fd_kernel<<<grid(kWpbEpi), 32 * kWpbEpi, sm_epi, stream>>>(a);
    else avg_epilogue_kernel<<<grid(kWpbEpi), 32 * kWpbEpi, sm_epi, stream>>>(a);
    mark(4);
    if (kt_on) {
        cudaEventSynchronize(g_kt.ev[nev - 1]);
        for (int i = 1; i < nev; ++i) { float t = 0; cudaEventElapsedTime(&t, g_kt.ev[i - 1], g_kt.ev[i]); g_kt.ms[kinds[i]] += (double)t; }
        if (++g_kt.steps % 8 == 0) {                 // mean over the last 8 steps
            double tot = 0; for (int i = 0; i < 7; ++i) tot += g_kt.ms[i];
            fprintf(stderr, "[avg kernel times, steps %d-%d, %d envs] prologue %.3f collide %.3f pcollide %.3f narrow %.3f dynamics %.3f solve %.3f epilogue %.3f ms/step (total %.3f)\n",
                    g_kt.steps - 8, g_kt.steps - 1, a.n_env, g_kt.ms[0] / 8, g_kt.ms[1] / 8, g_kt.ms[6] / 8, g_kt.ms[5] / 8, g_kt.ms[2] / 8, g_kt.ms[3] / 8, g_kt.ms[4] / 8, tot / 8);
            for (int i = 0; i < 7; ++i) g_kt.ms[i] = 0;
        }
    }
    if (launched) *launched = nlaunch;
    return cudaGetLastError();
}

cudaError_t avg_launch_settle(const AvgStepArgs& a_in, int n, cudaStream_t stream, int* launched) {
    cudaError_t ec = configure_kernels();
    if (ec != cudaSuccess) return ec;
    AvgStepArgs a = a_in;
    a.post = 0;                                      // reset() calls p.stepSimulation only: no per-frame hooks (feeding.py:318-320)
    if (a.env_end - a.env_begin <= 0) return cudaSuccess;
    const int n_internal = a.n_internal > 0 ? a.n_internal : 1;
    int nlaunch = 0;
    for (int i = 0; i < n * n_internal; ++i) launch_internal_step(a, stream, [&](int) { nlaunch++; });
    if (launched) *launched = nlaunch;
    return cudaGetLastError();
}

cudaError_t avg_register_model(int slot, int variant, const unsigned char* d_blob, const AvgModelHeader* hh) {
    if (slot < 0 || slot >= AVG_K_MAX_HANDLES || variant < 0 || variant >= AVG_K_MAX_VARIANTS) return cudaErrorInvalidValue;
    KM m;
    m.h = reinterpret_cast<const AvgModelHeader*>(d_blob);
    m.body = reinterpret_cast<const AvgBody*>(d_blob + hh->off_body);
    m.dof = reinterpret_cast<const AvgDof*>(d_blob + hh->off_dof);
    m.shape = reinterpret_cast<const AvgShape*>(d_blob + hh->off_shape);
    m.vert = reinterpret_cast<const float*>(d_blob + hh->off_vert);
    m.plane = reinterpret_cast<const float*>(d_blob + hh->off_plane);
    m.pair = reinterpret_cast<const uint32_t*>(d_blob + hh->off_pair);
    m.frame = reinterpret_cast<const AvgFrame*>(d_blob + hh->off_frame);
    m.bps = reinterpret_cast<const AvgBpStatic*>(d_blob + hh->off_bps);
    m.bpm = reinterpret_cast<const uint32_t*>(d_blob + hh->off_bpm);
    m.bcap = reinterpret_cast<const float4*>(d_blob + hh->off_bcap);
    m.mlp = hh->n_mlp > 0 ? reinterpret_cast<const float*>(d_blob + hh->off_mlp) : nullptr;
    m.target = hh->n_target > 0 ? reinterpret_cast<const float4*>(d_blob + hh->off_target) : nullptr;
    m.caabb = hh->n_cshape > 0 ? reinterpret_cast<const float4*>(d_blob + hh->off_caabb) : nullptr;
    return cudaMemcpyToSymbol(c_models, &m, sizeof(KM), sizeof(KM) * ((size_t)slot * AVG_K_MAX_VARIANTS + variant));
}

cudaError_t avg_launch_arm_limit(const unsigned char* blob, const float* q4, float* logits, int n, cudaStream_t stream) {
    avg_arm_limit_kernel<<<(n + 3) / 4, 128, 0, stream>>>(blob, q4, logits, n);
    return cudaGetLastError();
}

cudaError_t avg_launch_reset_obs(const AvgStepArgs& a, cudaStream_t stream) {
    const size_t sm_epi = sizeof(SmEpi) * kWpbEpi;
    cudaError_t e1 = configure_kernels();
    if (e1 != cudaSuccess) return e1;
    const int blocks = (a.env_end - a.env_begin + kWpbEpi - 1) / kWpbEpi;
    avg_reset_obs_kernel<<<blocks, 32 * kWpbEpi, sm_epi, stream>>>(a);
    return cudaGetLastError();
}
