/* gel_math.h -- the fp32 numeric contract of gel's render path, as straight-line functions.
 *
 * Every expression reproduces the reference's IEEE binary32 operation sequence (operand order and
 * association) so that the framebuffer is bit-identical to main.c compiled without contraction:
 *   vdot  main.c:200-203      vunit main.c:205-213      tviewnrm main.c:382-390   tviewtri main.c:372-380
 *   tperspective main.c:302-314   tviewport main.c:288-300   tbarycenter main.c:316-332
 *   tdraw main.c:342-370      pshade main.c:334-340
 * On the device each primitive is an explicit round-to-nearest intrinsic (__fmul_rn, __fadd_rn, ...),
 * which nvcc never contracts into FMA regardless of flags; the library is additionally built with
 * --fmad=false.  Divisions and the square root are the IEEE-exact __fdiv_rn / __fsqrt_rn / __frcp_rn.
 * The same header compiles as plain C++ on the host (tests/emu, -ffp-contract=off) so the operation
 * order can be checked against the oracle without a GPU.
 */
#ifndef GEL_MATH_H
#define GEL_MATH_H

#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define GEL_HD __host__ __device__ __forceinline__
#else
#define GEL_HD static inline
#endif

namespace gel {

#if defined(__CUDA_ARCH__)
GEL_HD float mul(float a, float b) { return __fmul_rn(a, b); }
GEL_HD float add(float a, float b) { return __fadd_rn(a, b); }
GEL_HD float sub(float a, float b) { return __fsub_rn(a, b); }
GEL_HD float dvd(float a, float b) { return __fdiv_rn(a, b); }
GEL_HD float rcp(float a) { return __frcp_rn(a); }          /* == 1.0f / a, correctly rounded */
GEL_HD float root(float a) { return __fsqrt_rn(a); }
GEL_HD int   trunc_i(float a) { return __float2int_rz(a); } /* C cast semantics, main.c:344-347,360-363 */
GEL_HD float i2f(int a) { return __int2float_rn(a); }
#else
GEL_HD float mul(float a, float b) { return a * b; }
GEL_HD float add(float a, float b) { return a + b; }
GEL_HD float sub(float a, float b) { return a - b; }
GEL_HD float dvd(float a, float b) { return a / b; }
GEL_HD float rcp(float a) { return 1.0f / a; }
GEL_HD float root(float a) { return sqrtf(a); }
GEL_HD int   trunc_i(float a) { return (int) a; }
GEL_HD float i2f(int a) { return (float) a; }
#endif

/* vdot, main.c:200-203: (ax*bx + ay*by) + az*bz */
GEL_HD float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return add(add(mul(ax, bx), mul(ay, by)), mul(az, bz));
}

/* Per-view constants derived from the basis; the reference recomputes vdot(x,eye) etc. per corner
 * (main.c:375-377) -- same operands, same result. */
struct ViewConst
{
    float xx, xy, xz, yx, yy, yz, zx, zy, zz;   /* basis rows x, y, z            main.c:510-512 */
    float xe, ye, ze;                           /* vdot(x,eye), vdot(y,eye), vdot(z,eye)       */
    float w, h, x0, y0;                         /* viewport scale/offset         main.c:290-293 */
};

GEL_HD ViewConst view_const(const float* b /* x[3] y[3] z[3] eye[3] */, int xres, int yres)
{
    ViewConst c;
    c.xx = b[0]; c.xy = b[1]; c.xz = b[2];
    c.yx = b[3]; c.yy = b[4]; c.yz = b[5];
    c.zx = b[6]; c.zy = b[7]; c.zz = b[8];
    c.xe = dot3(c.xx, c.xy, c.xz, b[9], b[10], b[11]);
    c.ye = dot3(c.yx, c.yy, c.yz, b[9], b[10], b[11]);
    c.ze = dot3(c.zx, c.zy, c.zz, b[9], b[10], b[11]);
    c.w = dvd(i2f(yres), 1.5f);
    c.h = dvd(i2f(yres), 1.5f);
    c.x0 = dvd(i2f(xres), 2.0f);
    c.y0 = dvd(i2f(yres), 4.0f);
    return c;
}

/* One corner: tviewtri -> tperspective -> tviewport give the screen position; tviewnrm -> tunit give the
 * view normal, of which the path only ever consumes vdot(lights, n) with lights = (0,0,1) (main.c:358,508).
 * out = (screen x, screen y, screen z, shade). */
GEL_HD void transform_corner(const ViewConst& c, float vx, float vy, float vz, float nx, float ny, float nz,
                             float& sx, float& sy, float& sz, float& shade)
{
    /* tviewnrm main.c:382-390 */
    const float ox = dot3(nx, ny, nz, c.xx, c.xy, c.xz);
    const float oy = dot3(nx, ny, nz, c.yx, c.yy, c.yz);
    const float oz = dot3(nx, ny, nz, c.zx, c.zy, c.zz);
    /* vunit main.c:205-213: v * (1.0f / sqrtf((x*x + y*y) + z*z)) */
    const float inv = rcp(root(add(add(mul(ox, ox), mul(oy, oy)), mul(oz, oz))));
    const float ux = mul(ox, inv), uy = mul(oy, inv), uz = mul(oz, inv);
    /* vdot(lights, n) with lights = (0,0,1), evaluated literally   main.c:358 */
    shade = dot3(0.0f, 0.0f, 1.0f, ux, uy, uz);
    /* tviewtri main.c:372-380 */
    const float tx = sub(dot3(vx, vy, vz, c.xx, c.xy, c.xz), c.xe);
    const float ty = sub(dot3(vx, vy, vz, c.yx, c.yy, c.yz), c.ye);
    const float tz = sub(dot3(vx, vy, vz, c.zx, c.zy, c.zz), c.ze);
    /* tperspective main.c:302-314 */
    const float zd = sub(1.0f, dvd(tz, 3.0f));
    const float px = dvd(tx, zd), py = dvd(ty, zd), pz = dvd(tz, zd);
    /* tviewport main.c:288-300 */
    sx = add(mul(c.w, px), c.x0);
    sy = add(mul(c.h, py), c.y0);
    sz = dvd(add(pz, 1.0f), 1.5f);
}

/* Per-triangle invariants of tbarycenter (main.c:319-324, 327-328): the reference recomputes them for
 * every pixel from the same operands, so hoisting them is exact. */
struct TriSetup
{
    float ax, ay, az, bz, cz;
    float v0x, v0y, v1x, v1y;
    float k0, k1;            /* v2.z*v0.z and v2.z*v1.z with v2.z = 0.0f - a.z   main.c:318,321,325-326 */
    float d00, d01, d11, den;
    int x0, y0, x1, y1;      /* inclusive bbox, truncating casts                 main.c:344-347 */
};

GEL_HD TriSetup tri_setup(float ax, float ay, float az, float bx, float by, float bz,
                          float cx, float cy, float cz)
{
    TriSetup s;
    s.ax = ax; s.ay = ay; s.az = az; s.bz = bz; s.cz = cz;
    s.x0 = trunc_i(fminf(ax, fminf(bx, cx)));
    s.y0 = trunc_i(fminf(ay, fminf(by, cy)));
    s.x1 = trunc_i(fmaxf(ax, fmaxf(bx, cx)));
    s.y1 = trunc_i(fmaxf(ay, fmaxf(by, cy)));
    s.v0x = sub(bx, ax); s.v0y = sub(by, ay);
    const float v0z = sub(bz, az);
    s.v1x = sub(cx, ax); s.v1y = sub(cy, ay);
    const float v1z = sub(cz, az);
    s.d00 = dot3(s.v0x, s.v0y, v0z, s.v0x, s.v0y, v0z);
    s.d01 = dot3(s.v0x, s.v0y, v0z, s.v1x, s.v1y, v1z);
    s.d11 = dot3(s.v1x, s.v1y, v1z, s.v1x, s.v1y, v1z);
    s.den = sub(mul(s.d00, s.d11), mul(s.d01, s.d01));
    const float v2z = sub(0.0f, az);
    s.k0 = mul(v2z, v0z);
    s.k1 = mul(v2z, v1z);
    return s;
}

/* Numerators of tbarycenter's v and w at integer pixel (x, y)      main.c:318,321,325-328 */
GEL_HD void bary_numerators(const TriSetup& s, float fx, float fy, float& nv, float& nw)
{
    const float v2x = sub(fx, s.ax);
    const float v2y = sub(fy, s.ay);
    const float d20 = add(add(mul(v2x, s.v0x), mul(v2y, s.v0y)), s.k0);
    const float d21 = add(add(mul(v2x, s.v1x), mul(v2y, s.v1y)), s.k1);
    nv = sub(mul(s.d11, d20), mul(s.d01, d21));
    nw = sub(mul(s.d00, d21), mul(s.d01, d20));
}

/* Full barycentric solve + inside test + depth                     main.c:327-329, 352, 355.
 * Returns true iff the reference's `bc.x >= 0 && bc.y >= 0 && bc.z >= 0` holds. */
GEL_HD bool bary_inside(const TriSetup& s, float nv, float nw, float& v, float& w, float& u, float& z)
{
    v = dvd(nv, s.den);
    w = dvd(nw, s.den);
    u = sub(sub(1.0f, v), w);
    if(!(v >= 0.0f && w >= 0.0f && u >= 0.0f)) return false;
    z = add(add(mul(v, s.bz), mul(w, s.cz)), mul(u, s.az));
    return true;
}

/* Exact early-out: v = nv/den is a NEGATIVE NON-ZERO float (so `v >= 0.0f` is false) whenever nv and den
 * have opposite signs and the quotient cannot round to -0.  With |nv| >= 1e-20 and |den| <= 1e18 the
 * quotient's magnitude is >= 1e-38, far above the 1.4e-45 subnormal floor, so skipping the division is
 * exact.  `sden` is +1/-1 = sign(den), or 0 to disable the shortcut (den zero, NaN or huge). */
GEL_HD float sign_guard(float den)
{
    const float ad = fabsf(den);
    if(!(ad > 0.0f) || !(ad <= 1e18f)) return 0.0f;
    return den > 0.0f ? 1.0f : -1.0f;
}
GEL_HD bool surely_negative(float num, float sden) { return mul(num, sden) < -1e-20f; }

/* Orderable 32-bit key of a float: a > b  <=>  zkey(a) > zkey(b) for all non-NaN a != b (and +0 > -0). */
GEL_HD uint32_t zkey(float z)
{
#if defined(__CUDA_ARCH__)
    const uint32_t b = __float_as_uint(z);
#else
    union { float f; uint32_t u; } cv; cv.f = z; const uint32_t b = cv.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
GEL_HD float zkey_inv(uint32_t k)
{
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}

/* Shading and texel address of a z-passing fragment               main.c:358-363
 * uv = (ta.x, ta.y, tb.x, tb.y, tc.x, tc.y); shade_{a,b,c} = vdot(lights, nrm.{a,b,c}).
 * Weights map (v,w,u) -> (b,c,a) everywhere. */
GEL_HD void fragment_shade(float v, float w, float u, const float* uv, float sa, float sb, float sc,
                           int tw, int th, int& xx, int& yy, int& shading)
{
    const float s = add(add(mul(v, uv[2]), mul(w, uv[4])), mul(u, uv[0]));
    const float t = add(add(mul(v, uv[3]), mul(w, uv[5])), mul(u, uv[1]));
    xx = trunc_i(mul(i2f(tw - 1), add(0.0f, s)));
    yy = trunc_i(mul(i2f(th - 1), sub(1.0f, t)));
    const float intensity = add(add(mul(v, sb), mul(w, sc)), mul(u, sa));
    const float clamped = intensity < 0.0f ? 0.0f : intensity > 1.0f ? 1.0f : intensity;
    shading = trunc_i(mul(255.0f, clamped));
}

/* the same with (float) (tw - 1), (float) (th - 1) converted once by the caller (identical values) */
GEL_HD void fragment_shade_f(float v, float w, float u, const float* uv, float sa, float sb, float sc,
                             float twm1, float thm1, int& xx, int& yy, int& shading)
{
    const float s = add(add(mul(v, uv[2]), mul(w, uv[4])), mul(u, uv[0]));
    const float t = add(add(mul(v, uv[3]), mul(w, uv[5])), mul(u, uv[1]));
    xx = trunc_i(mul(twm1, add(0.0f, s)));
    yy = trunc_i(mul(thm1, sub(1.0f, t)));
    const float intensity = add(add(mul(v, sb), mul(w, sc)), mul(u, sa));
    const float clamped = intensity < 0.0f ? 0.0f : intensity > 1.0f ? 1.0f : intensity;
    shading = trunc_i(mul(255.0f, clamped));
}

/* pshade, main.c:334-340 */
GEL_HD uint32_t pshade(uint32_t p, int shading)
{
    const uint32_t r = ((p >> 16) * (uint32_t) shading) >> 8;
    const uint32_t g = (((p >> 8) & 0xFFu) * (uint32_t) shading) >> 8;
    const uint32_t b = ((p & 0xFFu) * (uint32_t) shading) >> 8;
    return r << 16 | g << 8 | b;
}

/* position-salted checksum term (gelcu.h, gelcu_render hash_out) */
GEL_HD uint32_t salt_mix(uint32_t word, uint32_t index)
{
    uint32_t h = word ^ (index * 0x9E3779B1u);
    h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

} /* namespace gel */
#endif /* GEL_MATH_H */
