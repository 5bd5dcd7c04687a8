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

/* Exact bbox trimming.
 *
 * The reference tests every pixel of a triangle's bounding box (main.c:348-352), and a bbox of a small triangle is mostly
 * empty: its first column x0 = trunc(min x) and first row lie left of / below every vertex.  bbox_trim removes an EDGE
 * (first / last column, first / last row) of the box when every pixel on it provably fails the reference's inside test, so
 * skipping those pixels cannot change a frame.  "Provably" means the kernels' own exact rejection (may_be_inside):
 *     nv < -1e-20 (den <= 1e18)  =>  v < 0        nw likewise  =>  w < 0        nv + nw > den (1 + 1e-5)  =>  u < 0
 * with  nv = fl(fl(A d20) - fl(C d21)),  nw = fl(fl(B d21) - fl(C d20)),  d20 = fl(fl(fl(v2x v0x) + fl(v2y v0y)) + k0)  etc.
 * (A, B, C, D = d11, d00, d01, den after sign normalisation, v2x = fl(x - ax), v2y = fl(y - ay)).
 *
 * Proof.  Let NV(x, y) = A (X v0x + Y v0y + k0) - C (X v1x + Y v1y + k1), X = x - ax, Y = y - ay, in REAL arithmetic on the
 * same float constants; likewise NW and S = NV + NW.  They are affine in (x, y), so on a segment they are bounded by their
 * end-point values.  Standard forward error analysis of the float expressions (each operation: fl(t) = t (1 + e),
 * |e| <= u = 2^-24) gives, for every pixel of the box,
 *     |nv - NV| <= 7u (|A| M20 + |C| M21)      |nw - NW| <= 7u (|B| M21 + |C| M20)      |fl(nv + nw) - S| <= 8u (sum of both)
 * with M20 = max|X| |v0x| + max|Y| |v0y| + |k0| (maxima over the box: attained at its corners), M21 likewise.  Hence for a
 * pixel p on the segment [e1, e2]:  nv(p) <= NV(p) + E <= max(NV(e1), NV(e2)) + E <= max(nv(e1), nv(e2)) + 2E,  and if that is
 * < -1e-20 every pixel of the segment has nv < -1e-20 and is rejected; the same for nw, and with min / > den_hi for the sum.
 * E is taken as 32u (...) -- four times the bound -- plus 1e-24, which also covers the rounding of E's own evaluation and
 * multiplications that underflow (only applied when (|A| + |C|)(M20 + M21)-like magnitudes that bound every intermediate are <= 1e30 -- no overflow --
 * and |A| + |B| + |C| <= 1e12, so errors of underflowing products stay below 1e-30).  Any NaN makes every comparison false: nothing is trimmed.  The four corner values are
 * evaluated with the kernels' operation order (not required by the proof).  Soundness is also checked by brute force on the
 * host (tests/test_emu_math.py: every trimmed pixel is evaluated the reference's way). */
GEL_HD void bbox_trim(float ax, float ay, float v0x, float v0y, float v1x, float v1y, float k0, float k1,
                      float B /* d00 */, float C /* d01 */, float A /* d11 */, float D /* den, > 0 */,
                      int& x0, int& y0, int& x1, int& y1)
{
    if(x0 > x1 || y0 > y1) return;
    const float eps = D <= 1e18f ? -1e-20f : -INFINITY;
    const float den_hi = mul(D, 1.00001f);
    const float xa = sub(i2f(x0), ax), xb = sub(i2f(x1), ax), ya = sub(i2f(y0), ay), yb = sub(i2f(y1), ay);
    const float xa0 = mul(xa, v0x), xa1 = mul(xa, v1x), xb0 = mul(xb, v0x), xb1 = mul(xb, v1x);
    const float ya0 = mul(ya, v0y), ya1 = mul(ya, v1y), yb0 = mul(yb, v0y), yb1 = mul(yb, v1y);
    float nv[4], nw[4], sm[4];                                         /* corners (x0,y0) (x0,y1) (x1,y0) (x1,y1) */
    #define GEL_CORNER(k, cx0, cx1, cy0, cy1) { const float d20 = add(add(cx0, cy0), k0), d21 = add(add(cx1, cy1), k1); \
        nv[k] = sub(mul(A, d20), mul(C, d21)); nw[k] = sub(mul(B, d21), mul(C, d20)); sm[k] = add(nv[k], nw[k]); }
    GEL_CORNER(0, xa0, xa1, ya0, ya1) GEL_CORNER(1, xa0, xa1, yb0, yb1) GEL_CORNER(2, xb0, xb1, ya0, ya1) GEL_CORNER(3, xb0, xb1, yb0, yb1)
    #undef GEL_CORNER
    const float mx = fmaxf(fabsf(xa), fabsf(xb)), my = fmaxf(fabsf(ya), fabsf(yb));
    const float m20 = add(add(mul(mx, fabsf(v0x)), mul(my, fabsf(v0y))), fabsf(k0));
    const float m21 = add(add(mul(mx, fabsf(v1x)), mul(my, fabsf(v1y))), fabsf(k1));
    const float tv_mag = add(mul(fabsf(A), m20), mul(fabsf(C), m21)), tw_mag = add(mul(fabsf(B), m21), mul(fabsf(C), m20));
    /* guards, NaN-proof (a NaN makes the sums NaN and the comparisons false): the magnitudes that bound every intermediate
     * of the corner evaluations are far from overflow, and the Gram terms small enough for underflow errors to vanish */
    if(!(add(tv_mag, tw_mag) <= 1e30f) || !(add(add(fabsf(A), fabsf(B)), fabsf(C)) <= 1e12f) || !(D > 0.0f)) return;
    const float u64 = 3.814697265625e-06f;                               /* 64 u = 2 x 32 u: the "2E" of the proof */
    const float ev2 = add(mul(u64, tv_mag), 1e-24f), ew2 = add(mul(u64, tw_mag), 1e-24f);
    const float tv = sub(eps, ev2), tw = sub(eps, ew2), ts = add(den_hi, add(ev2, ew2));
    /* edge (i, j) is empty when one of the three conditions holds at both of its corners, with the slack */
    #define GEL_EDGE(i, j) (fmaxf(nv[i], nv[j]) < tv || fmaxf(nw[i], nw[j]) < tw || fminf(sm[i], sm[j]) > ts)
    const bool left = GEL_EDGE(0, 1), right = GEL_EDGE(2, 3), bottom = GEL_EDGE(0, 2), top = GEL_EDGE(1, 3);
    #undef GEL_EDGE
    x0 += left ? 1 : 0; x1 -= right ? 1 : 0; y0 += bottom ? 1 : 0; y1 -= top ? 1 : 0;
}

/* Row trimming of one bbox COLUMN (tile pipeline, band rasteriser).
 *
 * A large triangle covers a chord of each of its bbox columns; the rows above and below the chord fail the reference's inside test
 * and the kernel only finds that out by testing them.  row_trim removes, from the two ends of a column's row range, rows that
 * provably fail the kernels' exact cheap rejection (and hence main.c:352) -- the same three conditions as bbox_trim:
 *     nv < eps (eps = -1e-20, den <= 1e18)        nw < eps        nv + nw > den_hi = den (1 + 1e-5)
 *
 * Setting.  Column x, rows y_0 .. y_0 + n (n >= 1).  For one condition let a(j) be the float value the row loop computes at row
 * y_0 + j (nv, nw, or -(nv + nw)), a0 = a(0), a1 = a(n) the two END values, evaluated with the row loop's own operations;
 * "rejected" means a(j) < t0 (t0 = eps, eps, -den_hi).  In real arithmetic on the same float constants the quantity is AFFINE in
 * j:  A(j) = A0 + (j / n)(A1 - A0), and |a(j) - A(j)| <= Et for every pixel of the triangle's bbox, Et <= 8u M (forward error
 * analysis as in bbox_trim; M = the magnitude |A| M20 + |C| M21 (+ the w analogue for the sum) + den).  trim_slack supplies
 * e = 64u M + 1e-24 >= 8 Et, and t = fl(t0 - e).
 *
 * Claim.  With g0 = fl(t - a0) > 0 and dl = fl(fl(a1 - a0) + e):  every integer row j in [0, n] with
 *         j < n g0 / dl   (dl > 0)        or any j at all   (dl <= 0)
 * has a(j) < t0.  Proof: A(j) <= a0 + Et + (j/n)(a1 - a0 + 2Et), so a(j) <= a0 + (j/n) Cf + 2Et with Cf = a1 - a0 + 2Et.  Rounding
 * of g0 and dl costs at most u(|t| + |a0|) + 2u|a1 - a0| + u e <= e/16 (|a| <= M(1 + 8u), u M <= e/64), hence g0 <= t0 - a0 - 2Et - e/2
 * and dl >= Cf + e/2.  If Cf <= 0 then a(j) <= a0 + 2Et < t0.  If Cf > 0 then dl > Cf > 0 and j < n g0 / dl <= n (t0 - a0 - 2Et) / Cf
 * gives (j/n) Cf < t0 - a0 - 2Et, i.e. a(j) < t0.  The quotient is evaluated approximately (MUFU reciprocal, ~1 ulp, three roundings) and multiplied by
 * 0.999 before truncation, so the integer never exceeds the real bound; dl in (0, 1e-30] trims nothing (no reciprocal of a
 * denormal); a NaN anywhere makes every comparison false: nothing is trimmed.  If BOTH ends are below t the whole column goes
 * (A is affine: A(j) <= max(A0, A1)) -- bbox_trim's argument.  By symmetry the same holds from the other end with a0, a1 swapped.
 * Cuts of different conditions combine: rows j < max over the conditions of their lower cuts are each rejected by the condition
 * that attains the maximum; likewise at the top.  Soundness is also brute-forced on the host (tests/test_emu_math.py). */
/* approximate reciprocal for the trimming quotient (never part of the reference's arithmetic); the argument is > 1e-30 where it counts */
#if defined(__CUDA_ARCH__)
GEL_HD float fast_rcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
#else
GEL_HD float fast_rcp(float a) { return 1.0f / a; }
#endif

/* e_v, e_w of the proof for one triangle, from its (frame-clipped) bbox; +inf when the error analysis does not apply (overflow-
 * sized terms, NaN): every trimming comparison is then false.  Same magnitudes as bbox_trim, den added. */
GEL_HD void trim_slack(float ax, float ay, float v0x, float v0y, float v1x, float v1y, float k0, float k1,
                       float B /* d00 */, float C /* d01 */, float A /* d11 */, float D /* den, > 0 */,
                       int x0, int y0, int x1, int y1, float& ev, float& ew)
{
    const float xa = sub(i2f(x0), ax), xb = sub(i2f(x1), ax), ya = sub(i2f(y0), ay), yb = sub(i2f(y1), ay);
    const float mx = fmaxf(fabsf(xa), fabsf(xb)), my = fmaxf(fabsf(ya), fabsf(yb));
    const float m20 = add(add(mul(mx, fabsf(v0x)), mul(my, fabsf(v0y))), fabsf(k0));
    const float m21 = add(add(mul(mx, fabsf(v1x)), mul(my, fabsf(v1y))), fabsf(k1));
    const float tv_mag = add(add(mul(fabsf(A), m20), mul(fabsf(C), m21)), D), tw_mag = add(add(mul(fabsf(B), m21), mul(fabsf(C), m20)), D);
    const bool ok = add(tv_mag, tw_mag) <= 1e30f && add(add(fabsf(A), fabsf(B)), fabsf(C)) <= 1e12f && D > 0.0f;
    const float u64 = 3.814697265625e-06f;                               /* 64 u */
    ev = ok ? add(mul(u64, tv_mag), 1e-24f) : INFINITY;
    ew = ok ? add(mul(u64, tw_mag), 1e-24f) : INFINITY;
}

/* one condition: end values a0 (row 0) and a1 (row n), threshold t = t0 - e, nf = (float) n; raises lo / hi to the rows that can go
 * at either end.  Straight-line (selects only): it runs in the rasteriser's unit prologue with all 32 lanes on different columns. */
GEL_HD void row_trim_cond(float a0, float a1, float t, float e, int n, float nf, int& lo, int& hi)
{
    const float g0 = sub(t, a0), g1 = sub(t, a1);
    const bool f0 = g0 > 0.0f, f1 = g1 > 0.0f;                            /* rejected at the first / last row (NaN: false) */
    const float g = f0 ? g0 : g1;
    const float d = sub(a1, a0);
    const float dl = add(f0 ? d : -d, e);
    int c = trunc_i(mul(mul(mul(nf, g), fast_rcp(dl)), 0.999f));          /* NaN -> 0, huge -> INT_MAX */
    c = c < 0 ? 0 : c > n + 1 ? n + 1 : c;
    c = dl > 1e-30f ? c : dl <= 0.0f ? n + 1 : 0;                         /* no reciprocal of a denormal; NaN: leave the column alone */
    c = (f0 && f1) ? n + 1 : c;                                           /* both ends rejected: the whole column */
    c = (f0 || f1) ? c : 0;
    const int cl = f0 ? c : 0, ch = f0 ? 0 : c;
    lo = cl > lo ? cl : lo; hi = ch > hi ? ch : hi;
}

/* Rows [0, n] of a column -> skip `lo` rows at the bottom and `hi` at the top (lo + hi may exceed n + 1: the column is empty).
 * (nv0, nw0) / (nv1, nw1): the numerators at the first / last row, den_hi = den (1 + 1e-5), eps = -1e-20 or -inf (no guard). */
GEL_HD void row_trim(float nv0, float nw0, float nv1, float nw1, float eps, float den_hi, float ev, float ew, int n, int& lo, int& hi)
{
    lo = 0; hi = 0;
    const float es = add(ev, ew), nf = i2f(n);
    row_trim_cond(nv0, nv1, sub(eps, ev), ev, n, nf, lo, hi);
    row_trim_cond(nw0, nw1, sub(eps, ew), ew, n, nf, lo, hi);
    row_trim_cond(-add(nv0, nw0), -add(nv1, nw1), -add(den_hi, es), es, n, nf, lo, hi);
}

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
