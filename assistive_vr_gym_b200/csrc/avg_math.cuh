// avg_math.cuh — small float vector / quaternion helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#define AVG_FULL 0xffffffffu

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };
struct M3 { float m[9]; };            // row-major

__device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }

__device__ __forceinline__ Q4 mkq(float x, float y, float z, float w) { Q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
__device__ __forceinline__ Q4 ldq(const float* p) { return mkq(p[0], p[1], p[2], p[3]); }
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    return mkq(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
               a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
__device__ __forceinline__ Q4 qconj(Q4 q) { return mkq(-q.x, -q.y, -q.z, q.w); }
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
    float n = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return mkq(q.x * n, q.y * n, q.z * n, q.w * n);
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
    // v + 2 w (u x v) + 2 u x (u x v)
    V3 u = mk3(q.x, q.y, q.z);
    V3 t = cross(u, v) * 2.0f;
    return v + t * q.w + cross(u, t);
}
__device__ __forceinline__ V3 qrot_inv(Q4 q, V3 v) { return qrot(qconj(q), v); }
__device__ __forceinline__ Q4 qaxis(V3 a, float ang) {
    float s, c; sincosf(0.5f * ang, &s, &c);
    return mkq(a.x * s, a.y * s, a.z * s, c);
}
__device__ __forceinline__ M3 qmat(Q4 q) {
    M3 r; float x = q.x, y = q.y, z = q.z, w = q.w;
    r.m[0] = 1 - 2 * (y * y + z * z); r.m[1] = 2 * (x * y - z * w); r.m[2] = 2 * (x * z + y * w);
    r.m[3] = 2 * (x * y + z * w); r.m[4] = 1 - 2 * (x * x + z * z); r.m[5] = 2 * (y * z - x * w);
    r.m[6] = 2 * (x * z - y * w); r.m[7] = 2 * (y * z + x * w); r.m[8] = 1 - 2 * (x * x + y * y);
    return r;
}
__device__ __forceinline__ V3 mmul(const float* m, V3 v) {
    return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
__device__ __forceinline__ V3 mtmul(const float* m, V3 v) {
    return mk3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z, m[2] * v.x + m[5] * v.y + m[8] * v.z);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AVG_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(AVG_FULL, v, o));
    return v;
}
__device__ __forceinline__ V3 shfl3(V3 v, int src) {
    return mk3(__shfl_sync(AVG_FULL, v.x, src), __shfl_sync(AVG_FULL, v.y, src), __shfl_sync(AVG_FULL, v.z, src));
}
__device__ __forceinline__ Q4 shflq(Q4 v, int src) {
    return mkq(__shfl_sync(AVG_FULL, v.x, src), __shfl_sync(AVG_FULL, v.y, src), __shfl_sync(AVG_FULL, v.z, src), __shfl_sync(AVG_FULL, v.w, src));
}

// spatial vectors about a fixed reference point: motion [w; v], force [n; f]
struct Sv { V3 a, l; };
__device__ __forceinline__ Sv mksv(V3 a, V3 l) { Sv s; s.a = a; s.l = l; return s; }
__device__ __forceinline__ Sv operator+(Sv x, Sv y) { return mksv(x.a + y.a, x.l + y.l); }
__device__ __forceinline__ Sv operator-(Sv x, Sv y) { return mksv(x.a - y.a, x.l - y.l); }
__device__ __forceinline__ Sv operator*(Sv x, float s) { return mksv(x.a * s, x.l * s); }
__device__ __forceinline__ float dot(Sv x, Sv y) { return dot(x.a, y.a) + dot(x.l, y.l); }
__device__ __forceinline__ Sv shflsv(Sv v, int src) { return mksv(shfl3(v.a, src), shfl3(v.l, src)); }
// motion x motion
__device__ __forceinline__ Sv crm(Sv v, Sv x) { return mksv(cross(v.a, x.a), cross(v.a, x.l) + cross(v.l, x.a)); }
// motion x* force
__device__ __forceinline__ Sv crf(Sv v, Sv f) { return mksv(cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)); }

// rigid-body inertia about the reference point: mass, h = m c, symmetric I_O (xx, yy, zz, xy, xz, yz)
struct Inertia { float m; V3 h; float xx, yy, zz, xy, xz, yz; };
__device__ __forceinline__ V3 sym_mul(const Inertia& I, V3 w) {
    return mk3(I.xx * w.x + I.xy * w.y + I.xz * w.z, I.xy * w.x + I.yy * w.y + I.yz * w.z, I.xz * w.x + I.yz * w.y + I.zz * w.z);
}
__device__ __forceinline__ Sv inertia_mul(const Inertia& I, Sv v) {
    return mksv(sym_mul(I, v.a) + cross(I.h, v.l), v.l * I.m - cross(I.h, v.a));
}
