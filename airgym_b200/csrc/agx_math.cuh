// agx_math.cuh — per-env arithmetic of the fused quadrotor step, written once as
// __host__ __device__ inline functions: the sm_100a kernels in agx_step.cu instantiate it one env
// per thread, and tests/hostsim compiles the very same text with g++ to debug it against the oracle
// on a box without a GPU (test infrastructure only; the product never runs the host build).
//
// Numerics: IEEE fp32, no fast-math; clamps are comparison based so that NaN propagates exactly as
// torch.clamp / torch.max(torch.min()) do in the reference (quirk Q5, hovering.py:393-397).
#pragma once
#include <math.h>
#include <stdint.h>
#include "agx.h"

#if defined(__CUDACC__)
#define AGX_HD __host__ __device__ __forceinline__
#else
#define AGX_HD inline
#endif

namespace agx {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };  // xyzw, like the reference state row (hovering.py:75)

AGX_HD float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// Division / square root of the step.  Device: MUFU.RCP + FMUL (div.approx.ftz, <= 2 ulp for |b| in [2^-126, 2^126]; IEEE
// results for b = 0 / inf / NaN operands, so quirk Q5's 0/0 = NaN survives) and MUFU.SQRT / MUFU.RSQ (sqrt.approx.ftz:
// a subnormal argument counts as 0).  `x / y` itself would compile to div.full's range-scaling sequence (~8 SASS
// instructions per divide, 54 of the 974 main-path instructions per warp in the round-1 capture).  Host build: exact.
AGX_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
#else
    return a / b;
#endif
}
AGX_HD float fsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
AGX_HD float frsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
AGX_HD float sq(float x) { return x * x; }
AGX_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
AGX_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
AGX_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
AGX_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
AGX_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
AGX_HD V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
AGX_HD float norm(V3 a) { return fsqrt(dot(a, a)); }

// ---- rotation conventions of pytorch3d.transforms [EXT] restated ------------------

// quaternion_to_matrix on (w,x,y,z) = (q.w,q.x,q.y,q.z); row-major 3x3; 2/|q|^2 scaling.
AGX_HD void quat_to_matrix(Q4 q, float* m) {
    const float two_s = fdiv(2.0f, q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    m[0] = 1.0f - two_s * (q.y * q.y + q.z * q.z);
    m[1] = two_s * (q.x * q.y - q.z * q.w);
    m[2] = two_s * (q.x * q.z + q.y * q.w);
    m[3] = two_s * (q.x * q.y + q.z * q.w);
    m[4] = 1.0f - two_s * (q.x * q.x + q.z * q.z);
    m[5] = two_s * (q.y * q.z - q.x * q.w);
    m[6] = two_s * (q.x * q.z - q.y * q.w);
    m[7] = two_s * (q.y * q.z + q.x * q.w);
    m[8] = 1.0f - two_s * (q.x * q.x + q.y * q.y);
}

// euler_angles_to_matrix(a,'XYZ') = Rx(a0) Ry(a1) Rz(a2)
AGX_HD void euler_xyz_to_matrix(float a0, float a1, float a2, float* m) {
    const float c0 = cosf(a0), s0 = sinf(a0), c1 = cosf(a1), s1 = sinf(a1), c2 = cosf(a2), s2 = sinf(a2);
    // A = Rx Ry
    const float a00 = c1, a01 = 0.0f, a02 = s1;
    const float a10 = s0 * s1, a11 = c0, a12 = -s0 * c1;
    const float a20 = -c0 * s1, a21 = s0, a22 = c0 * c1;
    m[0] = a00 * c2 + a01 * s2;  m[1] = a01 * c2 - a00 * s2;  m[2] = a02;
    m[3] = a10 * c2 + a11 * s2;  m[4] = a11 * c2 - a10 * s2;  m[5] = a12;
    m[6] = a20 * c2 + a21 * s2;  m[7] = a21 * c2 - a20 * s2;  m[8] = a22;
}

AGX_HD float sqrt_pos(float x) { return x > 0.0f ? fsqrt(x) : 0.0f; }

// matrix_to_quaternion: max-of-four-candidates, returns xyzw with w >= 0 (standardised).
AGX_HD Q4 matrix_to_quat(const float* m) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6],
                m21 = m[7], m22 = m[8];
    const float qa0 = sqrt_pos(1.0f + m00 + m11 + m22);
    const float qa1 = sqrt_pos(1.0f + m00 - m11 - m22);
    const float qa2 = sqrt_pos(1.0f - m00 + m11 - m22);
    const float qa3 = sqrt_pos(1.0f - m00 - m11 + m22);
    float w, x, y, z, d;
    if (qa0 >= qa1 && qa0 >= qa2 && qa0 >= qa3) {
        d = 2.0f * (qa0 > 0.1f ? qa0 : 0.1f);
        const float id = fdiv(1.0f, d); w = qa0 * qa0 * id; x = (m21 - m12) * id; y = (m02 - m20) * id; z = (m10 - m01) * id;
    } else if (qa1 >= qa2 && qa1 >= qa3) {
        d = 2.0f * (qa1 > 0.1f ? qa1 : 0.1f);
        const float id = fdiv(1.0f, d); w = (m21 - m12) * id; x = qa1 * qa1 * id; y = (m10 + m01) * id; z = (m02 + m20) * id;
    } else if (qa2 >= qa3) {
        d = 2.0f * (qa2 > 0.1f ? qa2 : 0.1f);
        const float id = fdiv(1.0f, d); w = (m02 - m20) * id; x = (m10 + m01) * id; y = qa2 * qa2 * id; z = (m12 + m21) * id;
    } else {
        d = 2.0f * (qa3 > 0.1f ? qa3 : 0.1f);
        const float id = fdiv(1.0f, d); w = (m10 - m01) * id; x = (m20 + m02) * id; y = (m21 + m12) * id; z = qa3 * qa3 * id;
    }
    Q4 q;
    if (w < 0.0f) { q.x = -x; q.y = -y; q.z = -z; q.w = -w; }
    else { q.x = x; q.y = y; q.z = z; q.w = w; }
    return q;
}

// Hamilton product, xyzw
AGX_HD Q4 qmul(Q4 a, Q4 b) {
    Q4 r;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return r;
}
AGX_HD Q4 qconj(Q4 a) { Q4 r; r.x = -a.x; r.y = -a.y; r.z = -a.z; r.w = a.w; return r; }
AGX_HD Q4 qnormalize(Q4 a) {
    const float in = frsqrt(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
    Q4 r; r.x = a.x * in; r.y = a.y * in; r.z = a.z * in; r.w = a.w * in; return r;
}
// third column of the rotation matrix of a unit quaternion (body z in world)
AGX_HD V3 quat_body_z(Q4 q) {
    return v3(2.0f * (q.x * q.z + q.y * q.w), 2.0f * (q.y * q.z - q.x * q.w),
              1.0f - 2.0f * (q.x * q.x + q.y * q.y));
}
AGX_HD V3 mat_mul_v(const float* m, V3 v) {
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z,
              m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
AGX_HD V3 mat_tmul_v(const float* m, V3 v) {
    return v3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
              m[2] * v.x + m[5] * v.y + m[8] * v.z);
}

// compute_yaw_diff(a, b) (hovering.py:33-38)
AGX_HD float yaw_diff(float a, float b) {
    float d = b - a;
    if (d < -kPi) d = d + kTwoPi;
    if (d > kPi) d = d - kTwoPi;
    return d;
}

// ---- Philox4x32-10 counter RNG (perf mode; explicit-randomness mode bypasses it) ----------------------
struct U4 { uint32_t x, y, z, w; };

AGX_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

AGX_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one 32x32->64 multiply per lane pair (IMAD.WIDE.U32) yields both halves
        const uint64_t p0 = (uint64_t)0xD2511F53u * (uint64_t)c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * (uint64_t)c.z;
        U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0; n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1; n.w = (uint32_t)p0;
        c = n;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

// counter = (global env id lo, step lo, stream_id | step_hi<<8 ... , block)
struct PhiloxCtx { uint32_t k0, k1, env_lo, env_hi, step_lo, step_hi; };

AGX_HD U4 philox_block(const PhiloxCtx& c, uint32_t stream_id, uint32_t blk) {
    U4 ctr;
    ctr.x = c.env_lo;
    ctr.y = c.step_lo;
    ctr.z = (c.env_hi << 16) ^ (c.step_hi << 4) ^ stream_id;  // env_hi/step_hi are 0 in practice
    ctr.w = blk;
    return philox4x32_10(ctr, c.k0, c.k1);
}
AGX_HD float u32_to_unit(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }  // [0,1)

// Box-Muller on the two 16-bit halves of one Philox word: radius from the high half (u1 in (0,1], |z| <= 4.71),
// angle from the low half.  One 128-bit Philox block therefore yields 8 normals; sensor-noise quality, documented in
// DESIGN.md §4 (the reference draws torch.randn; the stream is builder-defined either way).
AGX_HD void box_muller16(uint32_t w, float* z0, float* z1) {
    const uint32_t hi = (w >> 16) + 1u, lo = w & 0xFFFFu;  // u1 = hi / 65536 in (0,1], u2 = lo / 65536 in [0,1)
    float s, c, r;
#if defined(__CUDA_ARCH__)
    // MUFU.LG2 / MUFU.SQRT / MUFU.SIN / MUFU.COS (u1 >= 2^-16 is never subnormal, so the .ftz forms are exact
    // equivalents): abs error ~2^-21 on N(0,1) draws, far below the noise itself
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)hi * 1.52587890625e-5f));
    r = fsqrt(lg * -1.3862943611198906f);  // -2 ln u1 = -2 ln2 * log2 u1 >= 0
    __sincosf((float)lo * 9.5873799242852576e-5f, &s, &c);
#else
    r = sqrtf(-2.0f * logf((float)hi * 1.52587890625e-5f));
    s = sinf((float)lo * 9.5873799242852576e-5f); c = cosf((float)lo * 9.5873799242852576e-5f);
#endif
    *z0 = r * c; *z1 = r * s;
}

// `count` uniforms of stream `sid` (0: pre-step reset, 1: post-step reset)
AGX_HD void philox_uniforms(const PhiloxCtx& c, uint32_t sid, int count, float* u) {
#pragma unroll
    for (int b = 0; b < AGX_RESET_DRAWS_MAX / 4; ++b) {
        if (b * 4 >= count) break;
        const U4 r = philox_block(c, sid, (uint32_t)b);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (b * 4 + j < count) u[b * 4 + j] = u32_to_unit(w[j]);
    }
}
// `count` (<= 24) standard normals of stream 2: block b → normals [8b, 8b+8)
AGX_HD void philox_normals(const PhiloxCtx& c, int count, float* z) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        if (b * 8 >= count) break;
        const U4 r = philox_block(c, 2u, (uint32_t)b);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float n0, n1;
            box_muller16(w[j], &n0, &n1);
            if (b * 8 + 2 * j < count) z[b * 8 + 2 * j] = n0;
            if (b * 8 + 2 * j + 1 < count) z[b * 8 + 2 * j + 1] = n1;
        }
    }
}

// matrix_to_quaternion(euler_angles_to_matrix(a,'XYZ')) in closed form: q = qx(a0) * qy(a1) * qz(a2) (Hamilton, half
// angles), standardised to w >= 0 like pytorch3d.  Same rotation as matrix_to_quat(euler_xyz_to_matrix()) to fp32
// rounding, at 3 sincos instead of 6 sin/cos + 4 sqrt + 4 divisions (reset paths only; hovering.py:323-324).
AGX_HD Q4 euler_xyz_to_quat(float a0, float a1, float a2) {
    float s0, c0, s1, c1, s2, c2;
#if defined(__CUDA_ARCH__)
    // MUFU.SIN/COS: abs error 2^-21.4 on [-pi, pi] — the reset half-angles are at most 0.32 rad, the quaternion stays within
    // 5e-7 of the libm result (parity floor 2e-5), and the reset path loses three range-reduction sequences per sample
    __sincosf(0.5f * a0, &s0, &c0); __sincosf(0.5f * a1, &s1, &c1); __sincosf(0.5f * a2, &s2, &c2);
#else
    s0 = sinf(0.5f * a0); c0 = cosf(0.5f * a0); s1 = sinf(0.5f * a1); c1 = cosf(0.5f * a1);
    s2 = sinf(0.5f * a2); c2 = cosf(0.5f * a2);
#endif
    Q4 q;
    q.w = c0 * c1 * c2 - s0 * s1 * s2;
    q.x = s0 * c1 * c2 + c0 * s1 * s2;
    q.y = c0 * s1 * c2 - s0 * c1 * s2;
    q.z = c0 * c1 * s2 + s0 * s1 * c2;
    if (q.w < 0.0f) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    return q;
}

// ---- reset samplers ---------------------------------------------------------------------------------
// U(lo,hi) exactly as torch_rand_float: (hi - lo) * u + lo (airgym/utils/torch_utils.py:192-193)
AGX_HD float urange(float u, float lo, float hi) { return (hi - lo) * u + lo; }

// Hovering.reset_idx (hovering.py:310-335) / Tracking.reset_idx (tracking.py:159-192): 12 uniforms in
// call order xy(2) z(1) roll,pitch(2) yaw(1) linvel(3) angvel(3); writes the 13-float root state.
// Balloon.reset_idx (balloon.py:57-93): 15 uniforms — ball x,y,z (3) then drone xy(2) z(1) roll(1) pitch(1) yaw(1)
// linvel(3) angvel(3).  aux = [ball xyz | previous drone xyz | collision flag | pad].
AGX_HD void reset_sample_balloon(const float* u, float* s, float* aux) {
    aux[0] = 0.5f * urange(u[0], -1.0f, 1.0f) + 2.5f;
    aux[1] = 2.0f * urange(u[1], -1.0f, 1.0f) + 0.0f;
    aux[2] = 0.3f * urange(u[2], -1.0f, 1.0f) + 1.0f;
    s[0] = 0.1f * urange(u[3], -1.0f, 1.0f) + 0.0f;
    s[1] = 0.1f * urange(u[4], -1.0f, 1.0f) + 0.0f;
    s[2] = 0.2f * urange(u[5], -1.0f, 1.0f) + 1.0f;
    const float a0 = 0.1f * urange(u[6], -kPi, kPi);
    const float a1 = 0.1f * urange(u[7], 0.0f, kPi);
    const float a2 = 0.2f * urange(u[8], -kPi, kPi);
    const Q4 q = euler_xyz_to_quat(a0, a1, a2);
    s[3] = q.x; s[4] = q.y; s[5] = q.z; s[6] = q.w;
    s[7] = 0.5f * urange(u[9], -1.0f, 1.0f);
    s[8] = 0.5f * urange(u[10], -1.0f, 1.0f);
    s[9] = 0.5f * urange(u[11], -1.0f, 1.0f);
    s[10] = 0.2f * urange(u[12], -1.0f, 1.0f);
    s[11] = 0.2f * urange(u[13], -1.0f, 1.0f);
    s[12] = 0.2f * urange(u[14], -1.0f, 1.0f);
    aux[3] = 0.0f; aux[4] = 0.0f; aux[5] = 0.0f;  // pre_root_positions[env_ids] = 0 (balloon.py:90)
}

// ---- scene stand-ins for IsaacGym cameras / PhysX contacts (builder-defined; oracle/scene.py is the restatement) ----
constexpr float kDroneRadius = 0.2f;   // robots/X152b/model.urdf:13-18 collision sphere
constexpr float kBallRadius = 0.2f;    // balls/ball/model.urdf
constexpr float kCubeHalf = 0.15f;     // cubes/1x1: +-1 mesh scaled 0.15
constexpr float kCamFar = 5.0f;        // avoid_config.py:60 far_plane
constexpr float kCamF = 111.70069f;    // (212 / 2) / tan(87 deg / 2): focal length in pixels
constexpr float kTMin = 1e-3f;
constexpr float kInf = __builtin_huge_valf();  // +inf

struct Capsule { V3 c, a; float r, h; };  // capped cylinder: centre, unit axis, radius, half length (world frame)

// tree `j` of an env: asset-frame table row (centre xyz, axis xyz, radius, half length) placed at (x, y, 0), yawed by (cy, sy)
AGX_HD Capsule place_tree(const float* t, float x, float y, float cy, float sy) {
    Capsule k;
    k.c = v3(cy * t[0] - sy * t[1] + x, sy * t[0] + cy * t[1] + y, t[2]);
    k.a = v3(cy * t[3] - sy * t[4], sy * t[3] + cy * t[4], t[5]);
    k.r = t[6];
    k.h = t[7];
    return k;
}

AGX_HD float hit_ground(V3 o, V3 d) {
    const float t = fdiv(-o.z, d.z);
    return (d.z < 0.0f && t > kTMin) ? t : kInf;
}
AGX_HD float hit_sphere(V3 o, V3 d, V3 c, float r) {
    const V3 oc = o - c;
    const float a = dot(d, d), b = dot(d, oc), cc = dot(oc, oc) - r * r;
    const float disc = b * b - a * cc;
    const float t = fdiv(-b - fsqrt(disc > 0.0f ? disc : 0.0f), a);
    return (disc > 0.0f && t > kTMin) ? t : kInf;
}
AGX_HD float hit_box(V3 o, V3 d, V3 c, float half) {
    const float ix = fdiv(1.0f, d.x), iy = fdiv(1.0f, d.y), iz = fdiv(1.0f, d.z);
    const float ax = (c.x - half - o.x) * ix, bx = (c.x + half - o.x) * ix;
    const float ay = (c.y - half - o.y) * iy, by = (c.y + half - o.y) * iy;
    const float az = (c.z - half - o.z) * iz, bz = (c.z + half - o.z) * iz;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    return (tn <= tf && tn > kTMin) ? tn : kInf;
}
AGX_HD float hit_capsule(V3 o, V3 d, const Capsule& k) {  // side wall, then the end cap facing the ray
    const V3 oc = o - k.c;
    const float card = dot(k.a, d), caoc = dot(k.a, oc);
    const float A = dot(d, d) - card * card;
    const float B = dot(oc, d) - caoc * card;
    const float C = dot(oc, oc) - caoc * caoc - k.r * k.r;
    const float disc = B * B - A * C;
    if (!(disc > 0.0f)) return kInf;
    const float sq = fsqrt(disc);
    const float t = fdiv(-B - sq, A);
    const float y = caoc + t * card;
    const float side = (fabsf(y) < k.h && t > kTMin) ? t : kInf;
    const float tc = fdiv((y < 0.0f ? -k.h : k.h) - caoc, card);
    const float cap = (fabsf(B + A * tc) < sq && tc > kTMin) ? tc : kInf;
    return fminf(side, cap);
}
// drone collision sphere vs a tree, the tree taken as a capsule (segment distance)
AGX_HD bool touch_capsule(V3 p, const Capsule& k) {
    const V3 rel = p - k.c;
    float s = dot(rel, k.a);
    s = s < -k.h ? -k.h : (s > k.h ? k.h : s);
    const V3 q = rel - s * k.a;
    return norm(q) < k.r + kDroneRadius;
}
AGX_HD bool touch_box(V3 p, V3 c, float half) {
    const float qx = fmaxf(fabsf(p.x - c.x) - half, 0.0f), qy = fmaxf(fabsf(p.y - c.y) - half, 0.0f),
                qz = fmaxf(fabsf(p.z - c.z) - half, 0.0f);
    return fsqrt(qx * qx + qy * qy + qz * qz) < kDroneRadius;
}
// one dt of the thrown cube (avoid.py leaves its flight to PhysX): semi-implicit Euler, rests where it lands; cubes parked at
// x = -999 (avoid.py:124-129) stay.  c = [px py pz vx vy vz]
AGX_HD void cube_step(float* c, float dt, float g) {
    if (c[0] == -999.0f) return;
    const float vz = c[5] - g * dt;
    float px = c[0] + c[3] * dt, py = c[1] + c[4] * dt, pz = c[2] + vz * dt;
    float vx = c[3], vy = c[4], vzz = vz;
    if (pz < kCubeHalf) { pz = kCubeHalf; vx = 0.0f; vy = 0.0f; vzz = 0.0f; }
    c[0] = px; c[1] = py; c[2] = pz; c[3] = vx; c[4] = vy; c[5] = vzz;
}

// ---- depth camera (builder-defined stand-in for IsaacGym's camera sensor; oracle/scene.py camera_rays) ----
struct Camera { V3 o; float R[9]; };
AGX_HD Camera make_camera(const float* s) {  // s = root-state row; unit quaternion assumed (1 - 2(yy+zz) form, like the oracle)
    Camera c;
    const float x = s[3], y = s[4], z = s[5], w = s[6];
    c.R[0] = 1.0f - 2.0f * (y * y + z * z); c.R[1] = 2.0f * (x * y - z * w); c.R[2] = 2.0f * (x * z + y * w);
    c.R[3] = 2.0f * (x * y + z * w); c.R[4] = 1.0f - 2.0f * (x * x + z * z); c.R[5] = 2.0f * (y * z - x * w);
    c.R[6] = 2.0f * (x * z - y * w); c.R[7] = 2.0f * (y * z + x * w); c.R[8] = 1.0f - 2.0f * (x * x + y * y);
    const V3 off = mat_mul_v(c.R, v3(0.15f, 0.0f, 0.1f));  // local_transform.p (avoid_config.py:66)
    c.o = v3(s[0] + off.x, s[1] + off.y, s[2] + off.z);
    return c;
}
// world direction of pixel (u, v), u along the width (to the left is +y body), v down; x_cam component = 1 so the ray
// parameter IS the planar depth
AGX_HD V3 pixel_dir(const Camera& c, int u, int v) {
    const float dy = fdiv((float)(AGX_CAM_W / 2) - (float)u - 0.5f, kCamF);
    const float dz = fdiv((float)(AGX_CAM_H / 2) - (float)v - 0.5f, kCamF);
    // R (1, dy, dz) as (R[:,0] + dy R[:,1]) + dz R[:,2]: the render kernel hoists the first half out of its 4-pixel groups
    return v3(fmaf(c.R[2], dz, fmaf(c.R[1], dy, c.R[0])), fmaf(c.R[5], dz, fmaf(c.R[4], dy, c.R[3])),
              fmaf(c.R[8], dz, fmaf(c.R[7], dy, c.R[6])));
}
// customized.py:402-404: beyond the far plane = no hit = +inf; clip at 4.5 m; / 4.5
AGX_HD float normalize_depth(float t) {
    if (t > kCamFar) t = kInf;
    t = t > 4.5f ? 4.5f : t;
    t = t < 0.0f ? 0.0f : t;
    return t / 4.5f;
}
// Image columns [u0, u1] a capsule can touch (conservative), or u0 > u1 when it cannot show up at all.  The segment is
// clipped to the half space 5 cm in front of the camera plane, its end points are projected to pixel columns and the band is
// padded by the projected radius (at the clipped segment's nearest depth, x2 for the perspective stretch off-axis) + 2 px.
AGX_HD void capsule_columns(const Camera& c, const Capsule& k, int* u0, int* u1) {
    *u0 = 1; *u1 = 0;
    const V3 rel = k.c - c.o;
    if (norm(rel) - (k.h + k.r) > 7.5f) return;  // 5 m far plane x |d|max = 1.475
    const V3 fwd = v3(c.R[0], c.R[3], c.R[6]), left = v3(c.R[1], c.R[4], c.R[7]);
    const float xc = dot(rel, fwd), yc = dot(rel, left), xa = dot(k.a, fwd), ya = dot(k.a, left);
    float s0 = -k.h, s1 = k.h;                    // segment parameter range with x_cam(s) = xc + s xa >= near
    const float near = 0.05f;
    if (xc + s0 * xa < near) {
        if (xc + s1 * xa < near) { if (xc + k.h * fabsf(xa) + k.r <= 0.0f) return; *u0 = 0; *u1 = AGX_CAM_W - 1; return; }
        s0 = (near - xc) / xa;
    } else if (xc + s1 * xa < near) {
        s1 = (near - xc) / xa;
    }
    const float xa0 = xc + s0 * xa, xa1 = xc + s1 * xa;
    const float ua = (float)(AGX_CAM_W / 2) - 0.5f - kCamF * (yc + s0 * ya) / xa0;
    const float ub = (float)(AGX_CAM_W / 2) - 0.5f - kCamF * (yc + s1 * ya) / xa1;
    const float xmin = fmaxf(fminf(xa0, xa1) - k.r, near);
    const float pad = 2.0f * kCamF * k.r / xmin + 2.0f;
    const float lo = fminf(ua, ub) - pad, hi = fmaxf(ua, ub) + pad;
    if (hi < 0.0f || lo > (float)(AGX_CAM_W - 1)) return;
    *u0 = lo < 0.0f ? 0 : (int)lo;
    *u1 = hi > (float)(AGX_CAM_W - 1) ? AGX_CAM_W - 1 : (int)hi;
}

// Avoid.reset_idx (avoid.py:91-158) on the 11 compact draws [u_mode, theta, aim xyz, x, y, z, roll, pitch, yaw]
// (zero-weighted reference draws dropped).  aux = [cube xyz | cube linvel xyz | collisions | pad].
AGX_HD void reset_sample_avoid(const float* u, float* s, float* aux) {
    if (u[0] < 0.8f) {  // thrown at the hover point (avoid.py:104-117, calculate_object_velocity :58-89)
        const float theta = (kPi / 6.0f) * urange(u[1], -1.0f, 1.0f);
        const float cx = 4.2f * cosf(theta), cy = 4.2f * sinf(theta), cz = 1.4f;
        const float tx = 0.3f * urange(u[2], -1.0f, 1.0f) + 0.0f, ty = 0.3f * urange(u[3], -1.0f, 1.0f) + 0.0f,
                    tz = 0.3f * urange(u[4], -1.0f, 1.0f) + 1.0f;
        const float dx = tx - cx, dy = ty - cy;
        const float dxy = sqrtf(dx * dx + dy * dy);
        const float t = dxy / 4.5f;
        aux[0] = cx; aux[1] = cy; aux[2] = cz;
        aux[3] = (dx / dxy) * 4.5f;
        aux[4] = (dy / dxy) * 4.5f;
        aux[5] = (tz - cz + 4.905f * (t * t)) / t;
    } else {
        aux[0] = -999.0f; aux[1] = -999.0f; aux[2] = 0.0f;
        aux[3] = 0.0f; aux[4] = 0.0f; aux[5] = 0.0f;
    }
    s[0] = 0.2f * urange(u[5], -1.0f, 1.0f) + 0.0f;
    s[1] = 0.2f * urange(u[6], -1.0f, 1.0f) + 0.0f;
    s[2] = 0.2f * urange(u[7], -1.0f, 1.0f) + 1.0f;
    const Q4 q = euler_xyz_to_quat(0.01f * urange(u[8], -kPi, kPi), 0.01f * urange(u[9], -kPi, kPi), 0.05f * urange(u[10], -kPi, kPi));
    s[3] = q.x; s[4] = q.y; s[5] = q.z; s[6] = q.w;
#pragma unroll
    for (int i = 7; i < 13; ++i) s[i] = 0.0f;
}

// Planning.reset_idx (planning.py:63-136): draw d of the 124 compact draws [asset x(41) | y(41) | yaw(41) | goal y] → its slot(s)
// in the env's asset row [x(41) | y(41) | cos yaw(41) | sin yaw(41)].  The goal fix-up (asset 0 = the ball: x = 8.5,
// y = 1.5 U(-1,1)) must be applied AFTER all 124 draws are placed: planning_goal_fixup.
AGX_HD void planning_place_draw(int d, float u, float* row) {
    if (d < AGX_NUM_ASSETS) row[d] = 8.0f * urange(u, -1.0f, 1.0f) + 0.0f;             // LENGTH
    else if (d < 2 * AGX_NUM_ASSETS) row[d] = 4.0f * urange(u, -1.0f, 1.0f) + 0.0f;    // WIDTH
    else if (d < 3 * AGX_NUM_ASSETS) {
        const float yaw = urange(u, -kPi, kPi);
        row[d] = cosf(yaw);
        row[d + AGX_NUM_ASSETS] = sinf(yaw);
    }
}
// goal + drone pose from the last draw; aux = [goal xyz | pre_root_positions xyz | collisions | esdf_dist]
AGX_HD void planning_goal_fixup(float u_goal, float* row, float* s, float* aux) {
    const float gy = 1.5f * urange(u_goal, -1.0f, 1.0f) + 0.0f;
    row[0] = 8.5f;
    row[AGX_NUM_ASSETS] = gy;
    aux[0] = 8.5f; aux[1] = gy; aux[2] = 1.5f;
    aux[3] = 0.0f; aux[4] = 0.0f; aux[5] = 0.0f;  // pre_root_positions[env_ids] = 0 (planning.py:119)
    s[0] = -8.5f; s[1] = 0.0f; s[2] = 1.5f;
    const float yaw0 = atan2f(gy - 0.0f, 8.5f - (-8.5f));  // compute_direction_angle (planning.py:86-99)
    const Q4 q = euler_xyz_to_quat(0.0f, 0.0f, yaw0);
    s[3] = q.x; s[4] = q.y; s[5] = q.z; s[6] = q.w;
#pragma unroll
    for (int i = 7; i < 13; ++i) s[i] = 0.0f;
}

template <int TASK>
AGX_HD void reset_sample(const float* u, float* s, float* aux) {
    if (TASK == AGX_TASK_BALLOON) { reset_sample_balloon(u, s, aux); return; }
    if (TASK == AGX_TASK_AVOID) { reset_sample_avoid(u, s, aux); return; }
    float a0, a1, a2;
    if (TASK == AGX_TASK_TRACKING) {
        s[0] = 0.1f * urange(u[0], -1.0f, 1.0f);
        s[1] = 0.1f * urange(u[1], -1.0f, 1.0f);
        s[2] = 0.1f * urange(u[2], -1.0f, 1.0f) + 1.0f;
        a0 = 0.1f * urange(u[3], -kPi, kPi);
        a1 = 0.1f * urange(u[4], -kPi, kPi);
        a2 = 0.2f * urange(u[5], -kPi, kPi);
    } else {
        s[0] = urange(u[0], -1.0f, 1.0f);
        s[1] = urange(u[1], -1.0f, 1.0f);
        s[2] = urange(u[2], -1.0f, 1.0f);
        a0 = 0.01f * urange(u[3], -kPi, kPi);
        a1 = 0.01f * urange(u[4], -kPi, kPi);
        a2 = 0.05f * urange(u[5], -kPi, kPi);
    }
    const Q4 q = euler_xyz_to_quat(a0, a1, a2);
    s[3] = q.x; s[4] = q.y; s[5] = q.z; s[6] = q.w;
    s[7] = 0.5f * urange(u[6], -1.0f, 1.0f);
    s[8] = 0.5f * urange(u[7], -1.0f, 1.0f);
    s[9] = 0.5f * urange(u[8], -1.0f, 1.0f);
    s[10] = 0.2f * urange(u[9], -1.0f, 1.0f);
    s[11] = 0.2f * urange(u[10], -1.0f, 1.0f);
    s[12] = 0.2f * urange(u[11], -1.0f, 1.0f);
}

// ---- PX4-aligned controller cascade (builder-defined; replaces the rlPx4Controller FFI, -------------
// hovering.py:217-250).  Everything is expressed in the sim's FLU body / ENU world frames; with
// PX4's diagonal gains this is identical to converting to FRD/NED, running PX4 and converting back.
// Controller state cs[]: [0:3] rate integrator, [3:6] previous body rate, [6:9] velocity integrator,
// [9:12] previous world velocity.

// quad-X mixer for the URDF rotor order prop_1(+x,-y) prop_2(-x,+y) prop_3(+x,+y) prop_4(-x,-y) with
// reaction torques (-,-,+,+) (hovering.py:272-275): c_i = T + s_r tau_x + s_p tau_y + s_y tau_z.
AGX_HD void mixer(float T, V3 tau, float* cmd) {
    cmd[0] = clampf(T - tau.x - tau.y - tau.z, 0.0f, 1.0f);
    cmd[1] = clampf(T + tau.x + tau.y - tau.z, 0.0f, 1.0f);
    cmd[2] = clampf(T + tau.x - tau.y + tau.z, 0.0f, 1.0f);
    cmd[3] = clampf(T - tau.x + tau.y + tau.z, 0.0f, 1.0f);
}

// PX4 RateControl: tau = P e + int - D d(w)/dt; integrator faded out for large errors, clamped.
AGX_HD V3 rate_loop(const AgxParams& P, V3 w_sp, V3 w_b, float* cs) {
    const float wsp[3] = {w_sp.x, w_sp.y, w_sp.z};
    const float wb[3] = {w_b.x, w_b.y, w_b.z};
    float tau[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float e = wsp[i] - wb[i];
        const float wdot = fdiv(wb[i] - cs[3 + i], P.dt);
        tau[i] = P.rate_p[i] * e + cs[i] - P.rate_d[i] * wdot;
        const float ef = fdiv(e, P.rate_i_fade);
        float fade = 1.0f - ef * ef;
        fade = fade < 0.0f ? 0.0f : fade;
        cs[i] = clampf(cs[i] + fade * P.rate_i[i] * e * P.dt, -P.rate_int_lim, P.rate_int_lim);
        cs[3 + i] = wb[i];
    }
    return v3(tau[0], tau[1], tau[2]);
}

// PX4 AttitudeControl::update: reduced-attitude (tilt first) law with yaw weight → body-rate set-point.
AGX_HD V3 attitude_loop(const AgxParams& P, Q4 q, Q4 qd) {
    const V3 ez = quat_body_z(q);
    const V3 ezd = quat_body_z(qd);
    const float d = dot(ez, ezd);
    Q4 qd_red;
    if (d < -1.0f + 1e-5f) {
        qd_red = qd;  // opposite thrust directions: no unique tilt rotation, use the full attitude
    } else {
        const V3 c = cross(ez, ezd);
        Q4 t; t.x = c.x; t.y = c.y; t.z = c.z; t.w = d + 1.0f;  // shortest rotation ez → ezd
        qd_red = qmul(qnormalize(t), q);
    }
    Q4 qmix = qmul(qconj(qd_red), qd);
    if (qmix.w < 0.0f) { qmix.x = -qmix.x; qmix.y = -qmix.y; qmix.z = -qmix.z; qmix.w = -qmix.w; }
    const float mw = clampf(qmix.w, -1.0f, 1.0f);
    const float mz = clampf(qmix.z, -1.0f, 1.0f);
    Q4 yawq; yawq.x = 0.0f; yawq.y = 0.0f;
    yawq.w = cosf(P.att_yaw_w * acosf(mw));
    yawq.z = sinf(P.att_yaw_w * asinf(mz));
    const Q4 qdd = qmul(qd_red, yawq);
    Q4 qe = qmul(qconj(q), qdd);
    const float sgn = qe.w < 0.0f ? -2.0f : 2.0f;
    V3 r;
    r.x = clampf(sgn * qe.x * P.att_p[0], -P.att_rate_lim[0], P.att_rate_lim[0]);
    r.y = clampf(sgn * qe.y * P.att_p[1], -P.att_rate_lim[1], P.att_rate_lim[1]);
    r.z = clampf(sgn * qe.z * P.att_p[2], -P.att_rate_lim[2], P.att_rate_lim[2]);
    return r;
}

// PX4 PositionControl::_velocityControl + _accelerationControl + bodyzToAttitude (ENU form):
// PID on velocity → specific-force vector → tilt limit → collective thrust + attitude set-point.
AGX_HD void velocity_loop(const AgxParams& P, V3 v_sp, float yaw_sp, V3 v, float* cs, Q4* q_sp,
                          float* thrust) {
    const float vsp[3] = {v_sp.x, v_sp.y, v_sp.z};
    const float vv[3] = {v.x, v.y, v.z};
    float acc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float e = vsp[i] - vv[i];
        const float vdot = fdiv(vv[i] - cs[9 + i], P.dt);
        acc[i] = P.vel_p[i] * e + cs[6 + i] - P.vel_d[i] * vdot;
        cs[6 + i] = clampf(cs[6 + i] + P.vel_i[i] * e * P.dt, -P.vel_int_lim[i], P.vel_int_lim[i]);
        cs[9 + i] = vv[i];
    }
    float fx = acc[0], fy = acc[1], fz = acc[2] + P.gravity;
    const float fz_min = 0.1f * P.gravity;
    fz = fz < fz_min ? fz_min : fz;
    const float h = fsqrt(fx * fx + fy * fy);
    const float hmax = fz * P.tilt_max_tan;
    if (h > hmax) { const float k = fdiv(hmax, h); fx = fx * k; fy = fy * k; }
    const float fn = fsqrt(fx * fx + fy * fy + fz * fz);
    const float ifn = fdiv(1.0f, fn);
    const V3 bz = v3(fx * ifn, fy * ifn, fz * ifn);
    *thrust = clampf(fdiv(P.hover_thrust * fn, P.gravity), P.thr_min, P.thr_max);
    // bodyzToAttitude: x axis from the yaw heading, orthogonalised against body z
    const V3 yc = v3(-sinf(yaw_sp), cosf(yaw_sp), 0.0f);
    V3 bx = cross(yc, bz);
    const float bxn = norm(bx);
    { const float ib = fdiv(1.0f, bxn); bx = v3(bx.x * ib, bx.y * ib, bx.z * ib); }
    const V3 by = cross(bz, bx);
    const float m[9] = {bx.x, by.x, bz.x, bx.y, by.y, bz.y, bx.z, by.z, bz.z};
    *q_sp = matrix_to_quat(m);
}

// Full cascade dispatch.  a[] = shaped+clamped action, s[] = root state row (canonical quaternion).
template <int MODE>
AGX_HD void controller(const AgxParams& P, const float* s, V3 w_b, const float* a, float* cs, float* cmd) {
    if (MODE == AGX_CTL_PROP) {  // hovering.py:251-252
        cmd[0] = a[0]; cmd[1] = a[1]; cmd[2] = a[2]; cmd[3] = a[3];
        return;
    }
    Q4 q; q.x = s[3]; q.y = s[4]; q.z = s[5]; q.w = s[6];
    V3 w_sp;  // w_b = R(q)^T w_world: world → body rates (what set_q_world is for, hovering.py:249)
    float thrust;
    if (MODE == AGX_CTL_RATE) {
        w_sp = v3(a[0], a[1], a[2]);
        thrust = a[3];
    } else {
        Q4 q_sp;
        if (MODE == AGX_CTL_ATTI) {  // action = (w,x,y,z,thrust)
            Q4 t; t.w = a[0]; t.x = a[1]; t.y = a[2]; t.z = a[3];
            const float n2 = t.w * t.w + t.x * t.x + t.y * t.y + t.z * t.z;
            if (n2 < 1e-12f) { t.w = 1.0f; t.x = 0.0f; t.y = 0.0f; t.z = 0.0f; }
            q_sp = qnormalize(t);
            thrust = a[4];
        } else {
            V3 v_sp;
            if (MODE == AGX_CTL_POS) {
                v_sp.x = clampf(P.pos_p[0] * (a[0] - s[0]), -P.vel_sp_lim[0], P.vel_sp_lim[0]);
                v_sp.y = clampf(P.pos_p[1] * (a[1] - s[1]), -P.vel_sp_lim[1], P.vel_sp_lim[1]);
                v_sp.z = clampf(P.pos_p[2] * (a[2] - s[2]), -P.vel_sp_lim[2], P.vel_sp_lim[2]);
            } else {
                v_sp = v3(a[0], a[1], a[2]);
            }
            velocity_loop(P, v_sp, a[3], v3(s[7], s[8], s[9]), cs, &q_sp, &thrust);
        }
        w_sp = attitude_loop(P, qnormalize(q), q_sp);
    }
    const V3 tau = rate_loop(P, w_sp, w_b, cs);
    mixer(thrust, tau, cmd);
}

// ---- rigid body (builder-defined; replaces gym.simulate, hovering.py:290) -----------
struct Deriv { V3 dp, dv, dw; Q4 dq; };

AGX_HD Deriv body_deriv(const AgxParams& P, V3 v, Q4 q, V3 w, float fz_over_m, V3 tau) {
    Deriv d;
    d.dp = v;
    const V3 bz = quat_body_z(q);
    d.dv = v3(bz.x * fz_over_m, bz.y * fz_over_m, bz.z * fz_over_m - P.gravity);
    d.dq.x = 0.5f * (q.w * w.x + q.y * w.z - q.z * w.y);
    d.dq.y = 0.5f * (q.w * w.y + q.z * w.x - q.x * w.z);
    d.dq.z = 0.5f * (q.w * w.z + q.x * w.y - q.y * w.x);
    d.dq.w = 0.5f * (-q.x * w.x - q.y * w.y - q.z * w.z);
    const V3 Iw = v3(P.inertia[0] * w.x, P.inertia[1] * w.y, P.inertia[2] * w.z);
    const V3 g = cross(w, Iw);
    d.dw = v3(fdiv(tau.x - g.x, P.inertia[0]), fdiv(tau.y - g.y, P.inertia[1]), fdiv(tau.z - g.z, P.inertia[2]));
    return d;
}

AGX_HD Q4 qaxpy(Q4 q, float h, Q4 d) {
    Q4 r; r.x = q.x + h * d.x; r.y = q.y + h * d.y; r.z = q.z + h * d.z; r.w = q.w + h * d.w; return r;
}

// One dt of the composite rigid body under a zero-order-hold body-frame wrench.
// s: in/out root state row; thrust[4]: rotor forces [N] (already zeroed for envs reset this step,
// hovering.py:268); tau_z: rotor reaction torque (hovering.py:270-275, NOT zeroed).
AGX_HD void integrate(const AgxParams& P, float* s, V3 w, const float* thrust, float tau_z, float* Rnew) {
    V3 p = v3(s[0], s[1], s[2]);
    Q4 q; q.x = s[3]; q.y = s[4]; q.z = s[5]; q.w = s[6];
    V3 v = v3(s[7], s[8], s[9]);  // w: body rates R(q)^T w_world, computed once by the caller
    const float fz_over_m = fdiv(thrust[0] + thrust[1] + thrust[2] + thrust[3], P.mass);
    const V3 tau = v3(P.arm * (-thrust[0] + thrust[1] + thrust[2] - thrust[3]),
                      P.arm * (-thrust[0] + thrust[1] - thrust[2] + thrust[3]), tau_z);
    const float h = P.dt;
    if (P.integrator == AGX_INT_EULER) {  // semi-implicit Euler (PhysX-like A/B switch)
        const Deriv k = body_deriv(P, v, q, w, fz_over_m, tau);
        v = v + h * k.dv;
        p = p + h * v;
        w = w + h * k.dw;
        const Deriv k2 = body_deriv(P, v, q, w, fz_over_m, tau);
        q = qaxpy(q, h, k2.dq);
    } else {  // classic RK4
        const Deriv k1 = body_deriv(P, v, q, w, fz_over_m, tau);
        const Deriv k2 = body_deriv(P, v + (0.5f * h) * k1.dv, qaxpy(q, 0.5f * h, k1.dq),
                                    w + (0.5f * h) * k1.dw, fz_over_m, tau);
        const Deriv k3 = body_deriv(P, v + (0.5f * h) * k2.dv, qaxpy(q, 0.5f * h, k2.dq),
                                    w + (0.5f * h) * k2.dw, fz_over_m, tau);
        const Deriv k4 = body_deriv(P, v + h * k3.dv, qaxpy(q, h, k3.dq), w + h * k3.dw, fz_over_m, tau);
        const float h6 = h * (1.0f / 6.0f);
        p = p + h6 * (k1.dp + 2.0f * k2.dp + 2.0f * k3.dp + k4.dp);
        v = v + h6 * (k1.dv + 2.0f * k2.dv + 2.0f * k3.dv + k4.dv);
        w = w + h6 * (k1.dw + 2.0f * k2.dw + 2.0f * k3.dw + k4.dw);
        q.x = q.x + h6 * (k1.dq.x + 2.0f * k2.dq.x + 2.0f * k3.dq.x + k4.dq.x);
        q.y = q.y + h6 * (k1.dq.y + 2.0f * k2.dq.y + 2.0f * k3.dq.y + k4.dq.y);
        q.z = q.z + h6 * (k1.dq.z + 2.0f * k2.dq.z + 2.0f * k3.dq.z + k4.dq.z);
        q.w = q.w + h6 * (k1.dq.w + 2.0f * k2.dq.w + 2.0f * k3.dq.w + k4.dq.w);
    }
    q = qnormalize(q);
    quat_to_matrix(q, Rnew);
    V3 ww = mat_mul_v(Rnew, w);  // back to world-frame angular velocity (IsaacGym root-state convention)
    const float vn = norm(v);
    if (vn > P.max_lin_vel) v = fdiv(P.max_lin_vel, vn) * v;
    const float wn = norm(ww);
    if (wn > P.max_ang_vel) ww = fdiv(P.max_ang_vel, wn) * ww;
    s[0] = p.x; s[1] = p.y; s[2] = p.z;
    s[3] = q.x; s[4] = q.y; s[5] = q.z; s[6] = q.w;
    s[7] = v.x; s[8] = v.y; s[9] = v.z;
    s[10] = ww.x; s[11] = ww.y; s[12] = ww.z;
}

// uniforms one reset_idx consumes (== AgxParams.reset_draws, checked at the C ABI)
template <int TASK>
struct ResetDraws {  // planning (124 draws) has its own sampler: planning_place_draw / planning_goal_fixup
    static constexpr int kD = (TASK == AGX_TASK_BALLOON) ? 15 : (TASK == AGX_TASK_AVOID ? 11 : (TASK == AGX_TASK_PLANNING ? AGX_PLANNING_DRAWS : 12));
};

// ---- random source: explicit rows (parity mode) or Philox (perf mode) ---------------------------------
struct RandSrc {
    const float* reset_row;  // [2,D] or nullptr
    const float* noise_row;  // [18] or nullptr
    PhiloxCtx ph;
};

AGX_HD void draw_reset(const RandSrc& r, int which, int D, float* u) {
    if (r.reset_row) {
        for (int i = 0; i < D; ++i) u[i] = r.reset_row[which * D + i];
    } else {
        philox_uniforms(r.ph, (uint32_t)which, D, u);
    }
}
AGX_HD void draw_noise(const RandSrc& r, float* z) {
    if (r.noise_row) {
#pragma unroll
        for (int i = 0; i < AGX_NOISE_DRAWS; ++i) z[i] = r.noise_row[i];
    } else {
        philox_normals(r.ph, AGX_NOISE_DRAWS, z);
    }
}

// ---- the fused per-env step ----------------------------------------------------------------------------
// Per-env working set held in registers by the kernel (and in a plain struct by the host build).
struct EnvRegs {
    float s[13];
    float a[AGX_MAX_ACTIONS];    // in: raw action; out: shaped + clamped (reference self.actions)
    float pa[AGX_MAX_ACTIONS];   // reference self.pre_actions
    float cs[AGX_CTRL_STATE_MAX];
    int64_t progress;
    int pending;                 // in: reference reset_buf != 0
    int reset;                   // out: new reset_buf
    int timeout;                 // out: time_out_buf
    float a_last_remap;          // out: 0.5+0.5a of the last action column (Q4 write-back)
    float rew;
    float cmd[4];
    float terms[AGX_REWARD_TERMS];
    float aux[AGX_AUX_MAX];      // task state beyond the drone — balloon: ball xyz, previous drone xyz, collision flag; avoid: cube xyz,
                                 // cube linvel xyz, collision flag; planning: goal xyz, previous drone xyz, collision flag, esdf_dist
};

// Bookkeeping half of reset_idx (hovering.py:332-335): the sampled state/aux are supplied by the caller (the kernel
// computes them warp-cooperatively, the host build serially).
AGX_HD void reset_apply(const AgxParams& P, EnvRegs& e) {
    e.progress = 0;
#pragma unroll
    for (int i = 0; i < AGX_MAX_ACTIONS; ++i) e.pa[i] = 0.0f;
    if (P.flags & AGX_FLAG_CTRL_RESET) {
#pragma unroll
        for (int i = 0; i < AGX_CTRL_STATE_MAX; ++i) e.cs[i] = 0.0f;
    }
}

// what the planning task needs beyond EnvRegs: this env's asset row (in/out) and the tree table
struct SceneRef { float* assets_row; const float* trees; };

// Planning.reset_idx, serially (host build, standalone reset_idx kernel): 124 draws = 31 Philox blocks of stream `which`
AGX_HD void reset_planning_serial(const RandSrc& r, int which, float* s, float* aux, float* row) {
    float u_goal = 0.0f;
    for (int b = 0; b < AGX_PLANNING_DRAWS / 4; ++b) {
        float u[4];
        if (r.reset_row) {
            for (int j = 0; j < 4; ++j) u[j] = r.reset_row[which * AGX_PLANNING_DRAWS + b * 4 + j];
        } else {
            const U4 w = philox_block(r.ph, (uint32_t)which, (uint32_t)b);
            u[0] = u32_to_unit(w.x); u[1] = u32_to_unit(w.y); u[2] = u32_to_unit(w.z); u[3] = u32_to_unit(w.w);
        }
        for (int j = 0; j < 4; ++j) {
            const int d = b * 4 + j;
            if (d == AGX_PLANNING_DRAWS - 1) u_goal = u[j];
            else planning_place_draw(d, u[j], row);
        }
    }
    planning_goal_fixup(u_goal, row, s, aux);
}

template <int TASK>
AGX_HD void do_reset(const AgxParams& P, const RandSrc& rnd, int which, EnvRegs& e, const SceneRef& sc) {
    if (TASK == AGX_TASK_PLANNING) {
        reset_planning_serial(rnd, which, e.s, e.aux, sc.assets_row);
    } else {
        float u[AGX_RESET_DRAWS_MAX];
        draw_reset(rnd, which, ResetDraws<TASK>::kD, u);
        reset_sample<TASK>(u, e.s, e.aux);  // root_states[ids] = initial (zeros + identity quat) then overwritten
    }
    reset_apply(P, e);
}

// Tracking.compute_traj_lemniscate (tracking.py:194-200), point k of 10
AGX_HD V3 lemniscate(int64_t progress, int k, float dt) {
    const float t = (float)(progress + 5 * k) * dt * 0.25f;
    const float st = sinf(t), ct = cosf(t);
    const float den = 1.0f + ct * ct;
    { const float id = fdiv(1.0f, den); return v3(3.0f * st * id, 3.0f * st * ct * id, 1.0f); }
}

// Observation noise draws, already scaled by sigma (hovering.py:350-353).  Independent of the env state, so the
// kernel evaluates it while the state/action loads are still in flight.
AGX_HD void scaled_noise(const AgxParams& P, const RandSrc& rnd, float* z) {
    if (P.flags & AGX_FLAG_NO_NOISE) {
#pragma unroll
        for (int i = 0; i < AGX_NOISE_DRAWS; ++i) z[i] = 0.0f;
        return;
    }
    draw_noise(rnd, z);
#pragma unroll
    for (int i = 0; i < 9; ++i) z[i] = P.noise_sigma[0] * z[i];
#pragma unroll
    for (int i = 9; i < 12; ++i) z[i] = P.noise_sigma[1] * z[i];
#pragma unroll
    for (int i = 12; i < 15; ++i) z[i] = P.noise_sigma[2] * z[i];
#pragma unroll
    for (int i = 15; i < 18; ++i) z[i] = P.noise_sigma[3] * z[i];
}

// The step between its two reset_idx passes, in two halves (the depth-camera tasks render between them on a render step,
// customized.py:318-325).  The caller has already applied the pre-step reset (hovering.py:209-211, quirk Q1) to `e` when
// e.pending, and applies the end-of-step reset when e.reset comes back set.
//
// env_phys: action shaping, controller cascade, rigid body, progress += 1, object flight and contacts.  R leaves as R(q_new).
template <int TASK, int MODE>
AGX_HD void env_phys(const AgxParams& P, EnvRegs& e, const SceneRef& sc, float* R) {
    constexpr int A = (MODE == AGX_CTL_ATTI) ? 5 : 4;
    constexpr bool kThrustMode = (MODE == AGX_CTL_RATE || MODE == AGX_CTL_ATTI);

    // -- action shaping (hovering.py:212-216).  Customized family (customized.py:226-232): the remap mutates
    //    self.actions in place but the clamp result only feeds the controller, so reward / pre_actions / the
    //    `actions` attribute see the remapped-but-UNCLAMPED values `ar`.
    constexpr bool kCustom = (TASK == AGX_TASK_BALLOON || TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING);
    if (kThrustMode) {
        e.a[A - 1] = 0.5f + 0.5f * e.a[A - 1];
        e.a_last_remap = e.a[A - 1];
    }
    float ar[AGX_MAX_ACTIONS];
#pragma unroll
    for (int i = 0; i < AGX_MAX_ACTIONS; ++i) ar[i] = e.a[i];
#pragma unroll
    for (int i = 0; i < A; ++i) {  // tensor_clamp = max(min(t, hi), lo)
        float t = e.a[i];
        t = (t > P.act_hi[i]) ? P.act_hi[i] : t;
        t = (t < P.act_lo[i]) ? P.act_lo[i] : t;
        e.a[i] = t;
    }

    // -- quaternion sign canonicalisation, written back into the state (hovering.py:224-226)
    if (e.s[6] < 0.0f) { e.s[3] = -e.s[3]; e.s[4] = -e.s[4]; e.s[5] = -e.s[5]; e.s[6] = -e.s[6]; }

    // -- body rates, shared by the controller and the integrator
    {
        Q4 q0; q0.x = e.s[3]; q0.y = e.s[4]; q0.z = e.s[5]; q0.w = e.s[6];
        quat_to_matrix(q0, R);
    }
    const V3 w_b = mat_tmul_v(R, v3(e.s[10], e.s[11], e.s[12]));

    // -- controller cascade → normalised rotor commands (hovering.py:235-252)
    controller<MODE>(P, e.s, w_b, e.a, e.cs, e.cmd);

    // -- wrench (hovering.py:256-281): thrust zeroed for envs reset this step, torque not
    float thrust[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) thrust[i] = e.pending ? 0.0f : e.cmd[i] * P.k_thrust;
    const float tau_z = P.k_torque * (-e.cmd[0] - e.cmd[1] + e.cmd[2] + e.cmd[3]);

    // -- gym.simulate (hovering.py:290); R becomes R(q_new)
    integrate(P, e.s, w_b, thrust, tau_z, R);

    e.progress += 1;  // hovering.py:297

    // Customized family from here on sees the remapped-but-unclamped actions only
    if (kCustom) {
#pragma unroll
        for (int i = 0; i < A; ++i) e.a[i] = ar[i];
        // -- object flight + check_collisions (customized.py:393-397), builder-defined contact model: the drone's r = 0.2
        //    collision sphere (model.urdf:13-18) against the ground plane, the thrown cube (avoid), the tree capsules (planning);
        //    the balloon / goal ball never collide.
        const V3 pn = v3(e.s[0], e.s[1], e.s[2]);
        bool hit = pn.z < P.collision_radius;
        if (TASK == AGX_TASK_AVOID) {
            cube_step(e.aux, P.dt, P.gravity);
            hit = hit || touch_box(pn, v3(e.aux[0], e.aux[1], e.aux[2]), kCubeHalf);
        }
        if (TASK == AGX_TASK_PLANNING) {
            const float* row = sc.assets_row;
            for (int j = 1; j < AGX_NUM_ASSETS; ++j) {
                const float dx = pn.x - row[j], dy = pn.y - row[AGX_NUM_ASSETS + j];
                if (dx * dx + dy * dy > 9.0f) continue;  // tree bounding cylinder: r + lean < 1.9 m from its root
                const Capsule k = place_tree(sc.trees + (j - 1) * 8, row[j], row[AGX_NUM_ASSETS + j], row[2 * AGX_NUM_ASSETS + j],
                                             row[3 * AGX_NUM_ASSETS + j]);
                hit = hit || touch_capsule(pn, k);
            }
        }
        e.aux[6] = hit ? 1.0f : 0.0f;
    }
}

// yaw-aligned local frame of the depth-camera tasks (avoid.py:203-226): W = Rz(yaw)^T, yaw = atan2(R10, R00)
struct LocalFrame { float c, s; float eul[3]; V3 v, w; };
AGX_HD LocalFrame local_frame(const float* R, V3 v, V3 w) {
    LocalFrame f;
    const float yaw = atan2f(R[3], R[0]);
    f.c = cosf(yaw); f.s = sinf(yaw);
    const float m00 = f.c * R[0] + f.s * R[3], m01 = f.c * R[1] + f.s * R[4], m02 = f.c * R[2] + f.s * R[5];
    const float m12 = -f.s * R[2] + f.c * R[5], m22 = R[8];
    f.eul[0] = atan2f(-m12, m22); f.eul[1] = asinf(m02); f.eul[2] = atan2f(-m01, m00);  // matrix_to_euler_angles(W R, 'XYZ')
    f.v = v3(f.c * v.x + f.s * v.y, -f.s * v.x + f.c * v.y, v.z);
    f.w = v3(f.c * w.x + f.s * w.y, -f.s * w.x + f.c * w.y, w.z);
    return f;
}

// env_task: observations, reward, termination flags, pre_actions.  R = R(q) of the current state.
template <int TASK, int MODE>
AGX_HD void env_task(const AgxParams& P, const float* z, EnvRegs& e, const float* R, float* obs) {
    constexpr int A = (MODE == AGX_CTL_ATTI) ? 5 : 4;
    constexpr bool kThrustMode = (MODE == AGX_CTL_RATE || MODE == AGX_CTL_ATTI);
    constexpr bool kImageTask = (TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING);
    const float* ar = e.a;  // Customized family: remapped, unclamped (env_phys left them in e.a)

    // -- compute_observations + add_noise (hovering.py:337-358; tracking.py:202-214); z = sigma * N(0,1)
    const V3 p = v3(e.s[0], e.s[1], e.s[2]);
    const V3 v = v3(e.s[7], e.s[8], e.s[9]);
    const V3 w = v3(e.s[10], e.s[11], e.s[12]);
    int reset;
    float reward;
    Q4 q; q.x = e.s[3]; q.y = e.s[4]; q.z = e.s[5]; q.w = e.s[6];
    const float up_z = (2.0f * q.w * q.w - 1.0f) + q.z * q.z * 2.0f;  // quat_axis(q,2)[2] (hovering.py:464-481)
    const float yaw = atan2f(-R[1], R[0]);  // pytorch3d matrix_to_euler_angles(.,'XYZ')[2] (quirk Q6)
    if (kImageTask) {
        const LocalFrame lf = local_frame(R, v, w);
        const float collided = e.aux[6];
        float sa = 0.0f, sd_all = 0.0f, sd_rate = 0.0f;
#pragma unroll
        for (int i = 0; i < A; ++i) {
            sa += ar[i] * ar[i];
            sd_all += sq(ar[i] - e.pa[i]);
            if (i < A - 1) sd_rate += sq(ar[i] - e.pa[i]);
        }
        const float ups_r = sq((up_z + 1.0f) / 2.0f);
        obs[3] = lf.eul[0]; obs[4] = lf.eul[1]; obs[5] = lf.eul[2];
        obs[6] = lf.v.x; obs[7] = lf.v.y; obs[8] = lf.v.z;
        obs[9] = lf.w.x; obs[10] = lf.w.y; obs[11] = lf.w.z;
#pragma unroll
        for (int i = 0; i < 4; ++i) obs[12 + i] = ar[i];  // actions_local aliases the tensor pre_physics_step remapped (avoid.py:162,226)
        if (TASK == AGX_TASK_AVOID) {  // avoid.py:203-295
            obs[0] = p.x - P.target[9]; obs[1] = p.y - P.target[10]; obs[2] = p.z - P.target[11];
            const V3 rel = v3(P.target[9] - p.x, P.target[10] - p.y, P.target[11] - p.z);
            const float heading = yaw_diff(P.target_yaw, yaw);
            const float distance = fsqrt(rel.x * rel.x + rel.y * rel.y + rel.z * rel.z + heading * heading);
            const float pose_r = fdiv(1.0f, 1.0f + sq(1.6f * distance));
            const float spin_r = fdiv(1.0f, 1.0f + sq(w.z * w.z));
            const float effort_r = 0.1f * expf(-sa);
            const float thrust_r = 0.05f * (1.0f - fabsf(0.1533f - ar[A - 1]));
            const float smooth_r = 0.1f * expf(-fsqrt(sd_rate));
            const float alive_r = collided > 0.0f ? -500.0f : 0.5f;
            reward = pose_r + pose_r * (ups_r + spin_r) + effort_r + smooth_r + thrust_r + alive_r;
            reset = (e.progress >= (int64_t)(P.max_episode_length - 1)) ? 1 : 0;
            if (p.z < 0.3f) reset = 1;
            if (p.z > 1.7f) reset = 1;
            if (norm(rel) > 2.0f) reset = 1;
            if (up_z < 0.0f) reset = 1;
            e.terms[0] = pose_r; e.terms[1] = ups_r; e.terms[2] = spin_r; e.terms[3] = effort_r; e.terms[4] = smooth_r;
            e.terms[5] = thrust_r; e.terms[6] = alive_r; e.terms[7] = 0.0f; e.terms[8] = reward;
        } else {  // planning.py:186-307
            const V3 goal = v3(e.aux[0], e.aux[1], e.aux[2]);
            const V3 fg = goal - p;
            const V3 pdl = v3(lf.c * fg.x + lf.s * fg.y, -lf.s * fg.x + lf.c * fg.y, fg.z);
            const float n = norm(pdl);
            const V3 gd = v3(fdiv(pdl.x, n), fdiv(pdl.y, n), fdiv(pdl.z, n));
            const float related = norm(fg);
            obs[0] = gd.x; obs[1] = gd.y; obs[2] = gd.z;
            const float cont = 0.2f * norm(lf.w) + 0.2f * fsqrt(sd_all);
            const float thrust_r = 0.5f * (1.0f - fabsf(0.1533f - ar[A - 1]));
            const V3 prev = v3(e.aux[3], e.aux[4], e.aux[5]);
            const float forward_r = 0.1f * (norm(goal - prev) - related);
            const float heading_r = gd.x * 1.0f + gd.y * 0.0f + gd.z * 0.0f;
            const float speed_r = -0.5f * (1.0f - expf(-2.0f * sq(lf.v.x - 1.0f)));
            const float z_r = fminf(fminf(p.z - 1.8f, 0.0f), 1.2f - p.z);
            const float esdf = e.aux[7];
            const float esdf_r = 0.5f * (1.0f - expf(-0.5f * sq(esdf)));
            const float alive_r = esdf > 0.3f ? 0.0f : -1.0f;
            const bool reach = related < 0.3f;
            const float reach_r = reach ? 200.0f : 0.0f;
            reward = cont + forward_r + alive_r + esdf_r + ups_r + z_r + speed_r + heading_r + thrust_r + reach_r;
            reset = 0;
            if (p.z < 1.5f - 0.3f) reset = 1;
            if (p.z > 1.5f + 0.3f) reset = 1;
            if (p.x < -8.0f - 0.5f) reset = 1;
            if (p.x > 8.0f + 0.5f) reset = 1;
            if (p.y < -4.0f) reset = 1;
            if (p.y > 4.0f) reset = 1;
            if (collided > 0.0f) reset = 1;
            if (reach) reset = 1;
            if (heading_r < 0.25f) reset = 1;
            if (e.progress >= (int64_t)(P.max_episode_length - 1)) reset = 1;
            e.aux[3] = p.x; e.aux[4] = p.y; e.aux[5] = p.z;  // pre_root_positions = root_positions.clone()
            e.terms[0] = cont; e.terms[1] = heading_r; e.terms[2] = speed_r; e.terms[3] = forward_r; e.terms[4] = alive_r;
            e.terms[5] = ups_r; e.terms[6] = z_r; e.terms[7] = esdf_r; e.terms[8] = thrust_r; e.terms[9] = reach_r;
            e.terms[10] = reward;
        }
        if ((P.flags & AGX_FLAG_RESET_ON_COLLISION) && collided > 0.0f) reset = 1;  // customized.py:328-330
    } else {
    float o[18];
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = R[i] + z[i];
    o[9] = p.x + z[9];
    o[10] = p.y + z[10];
    o[11] = p.z + z[11];
    o[12] = v.x + z[12];
    o[13] = v.y + z[13];
    o[14] = v.z + z[14];
    o[15] = w.x + z[15];
    o[16] = w.y + z[16];
    o[17] = w.z + z[17];
    V3 ref0 = v3(0.0f, 0.0f, 0.0f);
    if (TASK == AGX_TASK_TRACKING) {
#pragma unroll
        for (int i = 0; i < 18; ++i) obs[i] = o[i];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const V3 r = lemniscate(e.progress, k, P.dt);
            if (k == 0) ref0 = r;
            obs[18 + 3 * k + 0] = r.x - p.x;
            obs[18 + 3 * k + 1] = r.y - p.y;
            obs[18 + 3 * k + 2] = r.z - p.z;
        }
    } else if (TASK == AGX_TASK_BALLOON) {  // balloon.py:132-145: minus R(q_ball) (the ball keeps its identity quat), minus p_ball
        const float ident[9] = {1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
#pragma unroll
        for (int i = 0; i < 9; ++i) obs[i] = o[i] - ident[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) obs[9 + i] = o[9 + i] - e.aux[i];
#pragma unroll
        for (int i = 12; i < 18; ++i) obs[i] = o[i];
    } else {
#pragma unroll
        for (int i = 0; i < 18; ++i) obs[i] = o[i] - P.target[i];  // hovering.py:356
    }

    if (TASK == AGX_TASK_BALLOON) {
        // -- Balloon.compute_quadcopter_reward (balloon.py:159-225) on the remapped, unclamped actions
        const V3 ball = v3(e.aux[0], e.aux[1], e.aux[2]);
        const V3 rel = ball - p;
        const float check = norm(rel);
        const float nrm = check > 1e-12f ? check : 1e-12f;  // F.normalize eps
        const float dir_yaw = atan2f(fdiv(rel.y, nrm), fdiv(rel.x, nrm));
        const float yd_b = fabsf(yaw_diff(yaw, dir_yaw));
        const float yaw_r = fdiv(1.0f, 1.0f + sq(1.6f * yd_b));
        const V3 prev = v3(e.aux[3], e.aux[4], e.aux[5]);
        const float guidance = 30.0f * (norm(ball - prev) - check);
        const float ups_r = 0.5f * sq((up_z + 1.0f) / 2.0f);
        const float hit_r = check < 0.1f ? 800.0f : 0.0f;
        float sa = 0.0f, sd = 0.0f;
#pragma unroll
        for (int i = 0; i < A; ++i) { sa += ar[i] * ar[i]; sd += sq(ar[i] - e.pa[i]); }
        const float effort_b = 0.1f * expf(-sa);
        const float smooth = 0.1f * expf(-fsqrt(sd));
        reward = guidance + yaw_r + hit_r + smooth + ups_r + effort_b;
        reset = (e.progress >= (int64_t)(P.max_episode_length - 1)) ? 1 : 0;
        if (ar[A - 1] < -1.0f) reset = 1;
        if (ar[A - 1] > 1.0f) reset = 1;
        if (rel.x < -0.2f) reset = 1;
        if (v.x < 0.0f) reset = 1;
        if (check > 4.0f) reset = 1;
        if (p.z < 0.5f) reset = 1;
        if (p.z > 1.5f) reset = 1;
        if (check < 0.1f) reset = 1;
        const float collided = e.aux[6];  // check_collisions ran in env_phys
        if ((P.flags & AGX_FLAG_RESET_ON_COLLISION) && collided > 0.0f) reset = 1;  // customized.py:328-330
        e.aux[3] = p.x; e.aux[4] = p.y; e.aux[5] = p.z;  // pre_root_positions = root_positions.clone()
        e.terms[0] = guidance; e.terms[1] = hit_r; e.terms[2] = smooth; e.terms[3] = effort_b; e.terms[4] = ups_r;
        e.terms[5] = yaw_r; e.terms[6] = 0.0f; e.terms[7] = 0.0f;
    } else {
    // -- compute_quadcopter_reward (hovering.py:371-459; tracking.py:223-296)
    const float c0 = clampf(e.cmd[0], 0.0f, 1.0f), c1 = clampf(e.cmd[1], 0.0f, 1.0f),
                c2 = clampf(e.cmd[2], 0.0f, 1.0f), c3 = clampf(e.cmd[3], 0.0f, 1.0f);
    const float effort = 0.1f * ((1.0f - c0) + (1.0f - c1) + (1.0f - c2) + (1.0f - c3)) / 4.0f;
    float d[AGX_MAX_ACTIONS];
#pragma unroll
    for (int i = 0; i < A; ++i) d[i] = e.a[i] - e.pa[i];
    float cont, thrust_r = 0.0f;
    if (!kThrustMode) {
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < A; ++i) ss += d[i] * d[i];
        cont = 0.2f * expf(-fsqrt(ss));
    } else {
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < A - 1; ++i) ss += d[i] * d[i];
        if (TASK == AGX_TASK_TRACKING)
            cont = 0.1f * expf(-fsqrt(ss)) + fdiv(0.5f, 1.0f + sq(2.0f * d[A - 1]));
        else
            cont = 0.2f * expf(-fsqrt(ss)) + fdiv(0.5f, 1.0f + sq(3.0f * d[A - 1]));
        thrust_r = 0.1f * (1.0f - fabsf(0.1533f - e.a[A - 1]));
    }
    const float yd = yaw_diff(P.target_yaw, yaw) * (1.0f / kPi);
    const float wz2 = w.z * w.z;
    const float ups_r = sq((up_z + 1.0f) / 2.0f);
    if (TASK == AGX_TASK_TRACKING) {
        const V3 dd = ref0 - p;
        const float dist = norm(dd);
        const float dist_r = fdiv(1.0f, 1.0f + sq(1.8f * dist));
        const float yaw_r = fdiv(1.0f, 1.0f + sq(4.0f * yd));
        const float spin_r = fdiv(1.0f, 1.0f + sq(2.0f * wz2));
        reward = cont + effort;
        if (kThrustMode) reward = reward + thrust_r;
        reward = reward + dist_r + dist_r * (spin_r + yaw_r + ups_r);
        reset = (e.progress >= (int64_t)(P.max_episode_length - 1)) ? 1 : 0;
        if (dist > 1.0f) reset = 1;
        e.terms[0] = dist; e.terms[1] = dist_r; e.terms[2] = yaw_r; e.terms[3] = spin_r;
        e.terms[4] = cont; e.terms[5] = thrust_r; e.terms[6] = effort; e.terms[7] = ups_r;
    } else {
        const V3 rel = v3(P.target[9] - p.x, P.target[10] - p.y, P.target[11] - p.z);
        const float pd = norm(rel);
        const float pos_r = fdiv(0.7f, 1.0f + sq(1.6f * pd));
        const float vn = norm(v);
        const float dp = fdiv(rel.x, pd) * fdiv(v.x, vn) + fdiv(rel.y, pd) * fdiv(v.y, vn) + fdiv(rel.z, pd) * fdiv(v.z, vn);
        const float ang = fabsf(acosf(clampf(dp, -1.0f, 1.0f)));
        const float veld = 0.1f * expf(-ang * (1.0f / kPi));
        const float yaw_r = fdiv(1.0f, 1.0f + sq(3.0f * yd));
        const float spin_r = fdiv(1.0f, 1.0f + sq(3.0f * wz2));
        reward = cont + effort;
        if (kThrustMode) reward = reward + thrust_r;
        reward = reward + pos_r + pos_r * (veld + ups_r + spin_r + yaw_r);
        reset = (e.progress >= (int64_t)(P.max_episode_length - 1)) ? 1 : 0;
        if (pd > 4.0f) reset = 1;
        if (rel.z < -2.0f) reset = 1;
        if (rel.z > 2.0f) reset = 1;
        if (up_z < 0.0f) reset = 1;
        e.terms[0] = cont; e.terms[1] = effort; e.terms[2] = thrust_r; e.terms[3] = pos_r;
        e.terms[4] = veld; e.terms[5] = ups_r; e.terms[6] = spin_r; e.terms[7] = yaw_r;
    }
    if (MODE == AGX_CTL_ATTI && e.a[0] < 0.0f) reset = 1;  // hovering.py:442-444
    }  // !balloon
    e.terms[8] = reward;
    }  // !image task
    e.rew = reward;
    e.reset = reset;

    // -- pre_actions = actions.clone() (hovering.py:369); `e.a` leaves as the env's `actions` attribute
#pragma unroll
    for (int i = 0; i < A; ++i) e.pa[i] = e.a[i];
}

// both halves back to back (every task on a step without a render)
template <int TASK, int MODE>
AGX_HD void env_core(const AgxParams& P, const float* z, EnvRegs& e, const SceneRef& sc, float* obs) {
    float R[9];
    env_phys<TASK, MODE>(P, e, sc, R);
    env_task<TASK, MODE>(P, z, e, R, obs);
}

// time_out_buf (hovering.py:304), evaluated after the end-of-step reset zeroed progress
AGX_HD void env_finish(const AgxParams& P, EnvRegs& e) {
    e.timeout = (e.progress > (int64_t)P.max_episode_length) ? 1 : 0;
}

// The whole step for one env, serially (host build of tests/hostsim; the kernel interleaves the same pieces with its
// warp-cooperative reset sampling).
template <int TASK, int MODE>
AGX_HD void env_step(const AgxParams& P, const RandSrc& rnd, const float* z, EnvRegs& e, const SceneRef& sc, float* obs, int phase) {
    float R[9];
    if (phase != AGX_PHASE_TASK) {
        if (e.pending) do_reset<TASK>(P, rnd, 0, e, sc);  // pre_physics_step (hovering.py:209-211, quirk Q1)
        env_phys<TASK, MODE>(P, e, sc, R);
        if (phase == AGX_PHASE_PHYSICS) return;
    } else {
        Q4 q; q.x = e.s[3]; q.y = e.s[4]; q.z = e.s[5]; q.w = e.s[6];
        quat_to_matrix(q, R);
    }
    env_task<TASK, MODE>(P, z, e, R, obs);
    if (e.reset) do_reset<TASK>(P, rnd, 1, e, sc);    // end-of-step reset_idx (hovering.py:300-302): reset_buf stays 1
    env_finish(P, e);
}

}  // namespace agx
