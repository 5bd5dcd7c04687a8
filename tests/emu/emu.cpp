/* TEST-ONLY harness: compiles gel_b200/csrc/gel_math.h for the HOST so the operation order of the device
 * math and the keyed depth resolve (gelcu.cu) can be checked against the oracle in a container with no
 * GPU.  It walks triangles in REVERSE submission order on purpose: the (z, index) key must make the
 * result order-independent.  Never linked into the product; the product has no CPU path. */
#include "../../gel_b200/csrc/gel_math.h"

#include <cstdint>
#include <cstring>
#include <vector>

extern "C" void emu_transform(const float* tv, const float* tn, int ntri, const float* basis, int xres, int yres,
                              float* vew, float* shade)
{
    const gel::ViewConst c = gel::view_const(basis, xres, yres);
    for(int i = 0; i < 3 * ntri; i++)
        gel::transform_corner(c, tv[3 * i], tv[3 * i + 1], tv[3 * i + 2], tn[3 * i], tn[3 * i + 1], tn[3 * i + 2],
                              vew[3 * i], vew[3 * i + 1], vew[3 * i + 2], shade[i]);
}

extern "C" int emu_render(const float* tv, const float* tn, const float* tt, int ntri, const uint32_t* tex, int tw, int th,
                          int xres, int yres, const float* basis, uint32_t* pixel, float* zbuff, int use_sign_guard, int trim_rounds)
{
    const uint64_t CLEAR = (0x00800000ull << 32) | 0xFFFFFFFFull;
    std::vector<float> vew(9 * (size_t) ntri), shade(3 * (size_t) ntri);
    emu_transform(tv, tn, ntri, basis, xres, yres, vew.data(), shade.data());
    std::vector<uint64_t> keys((size_t) xres * yres, CLEAR);
    int flags = 0;
    for(int t = ntri - 1; t >= 0; t--)
    {
        const float* p = &vew[9 * (size_t) t];
        const gel::TriSetup s = gel::tri_setup(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
        int x0 = s.x0, y0 = s.y0, x1 = s.x1, y1 = s.y1;
        if(x0 < 0 || y0 < 0 || x1 > xres - 1 || y1 > yres - 1) { flags |= 1; if(x0 < 0) x0 = 0; if(y0 < 0) y0 = 0; if(x1 > xres - 1) x1 = xres - 1; if(y1 > yres - 1) y1 = yres - 1; }
        const float sden = use_sign_guard ? gel::sign_guard(s.den) : 0.0f;
        if(fabsf(s.den) > 0.0f)                                           /* the kernels' exact bbox trimming, on the sign-normalised terms */
        {
            const float sg = s.den < 0.0f ? -1.0f : 1.0f;
            for(int round = 0; round < trim_rounds; round++)
                gel::bbox_trim(s.ax, s.ay, s.v0x, s.v0y, s.v1x, s.v1y, s.k0, s.k1, s.d00 * sg, s.d01 * sg, s.d11 * sg, s.den * sg, x0, y0, x1, y1);
        }
        for(int x = x0; x <= x1; x++)
            for(int y = y0; y <= y1; y++)
            {
                float nv, nw, v, w, u, z;
                gel::bary_numerators(s, gel::i2f(x), gel::i2f(y), nv, nw);
                if(gel::surely_negative(nv, sden) || gel::surely_negative(nw, sden)) continue;
                if(!gel::bary_inside(s, nv, nw, v, w, u, z)) continue;
                const uint64_t key = ((uint64_t) gel::zkey(z) << 32) | (0xFFFFFFFFu - (uint32_t) t);
                uint64_t& k = keys[(size_t) y + (size_t) x * yres];
                if(key > k) k = key;
            }
    }
    for(int x = 0; x < xres; x++)
        for(int y = 0; y < yres; y++)
        {
            const size_t idx = (size_t) y + (size_t) x * yres;
            const uint64_t key = keys[idx];
            uint32_t colour = 0; float z = -FLT_MAX;
            if(key != CLEAR)
            {
                const uint32_t t = 0xFFFFFFFFu - (uint32_t) key;
                z = gel::zkey_inv((uint32_t) (key >> 32));
                const float* p = &vew[9 * (size_t) t];
                const gel::TriSetup s = gel::tri_setup(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
                float nv, nw, v, w, u, zz;
                gel::bary_numerators(s, gel::i2f(x), gel::i2f(y), nv, nw);
                gel::bary_inside(s, nv, nw, v, w, u, zz);
                const float uv[6] = { tt[9 * (size_t) t], tt[9 * (size_t) t + 1], tt[9 * (size_t) t + 3], tt[9 * (size_t) t + 4], tt[9 * (size_t) t + 6], tt[9 * (size_t) t + 7] };
                int xx, yy, shading;
                gel::fragment_shade(v, w, u, uv, shade[3 * (size_t) t], shade[3 * (size_t) t + 1], shade[3 * (size_t) t + 2], tw, th, xx, yy, shading);
                if(xx < 0 || xx > tw - 1 || yy < 0 || yy > th - 1) { flags |= 2; xx = xx < 0 ? 0 : xx > tw - 1 ? tw - 1 : xx; yy = yy < 0 ? 0 : yy > th - 1 ? th - 1 : yy; }
                colour = gel::pshade(tex[xx + yy * tw], shading);
            }
            pixel[idx] = colour; zbuff[idx] = z;
        }
    return flags;
}

extern "C" uint64_t emu_salted_sum(const uint32_t* w, uint64_t n)
{
    uint64_t s = 0;
    for(uint64_t i = 0; i < n; i++) s += gel::salt_mix(w[i], (uint32_t) i);
    return s;
}


/* Brute-force soundness of bbox_trim: for every triangle (9 screen-space floats) every pixel of its clipped bbox that the
 * trimming removes is evaluated the reference's way (tbarycenter, main.c:316-332, 352); returns how many of them are inside
 * (must be 0).  counts[0] += bbox pixels, counts[1] += pixels left after trimming, counts[2] += pixels inside. */
extern "C" uint64_t emu_trim_soundness(const float* vew, int ntri, int xres, int yres, int rounds, uint64_t* counts)
{
    uint64_t wrong = 0;
    for(int t = 0; t < ntri; t++)
    {
        const float* p = vew + 9 * (size_t) t;
        const gel::TriSetup s = gel::tri_setup(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
        int x0 = s.x0 < 0 ? 0 : s.x0, y0 = s.y0 < 0 ? 0 : s.y0, x1 = s.x1 > xres - 1 ? xres - 1 : s.x1, y1 = s.y1 > yres - 1 ? yres - 1 : s.y1;
        if(x0 > x1 || y0 > y1 || !(fabsf(s.den) > 0.0f)) continue;
        int tx0 = x0, ty0 = y0, tx1 = x1, ty1 = y1;
        const float sg = s.den < 0.0f ? -1.0f : 1.0f;
        for(int round = 0; round < rounds; round++)
            gel::bbox_trim(s.ax, s.ay, s.v0x, s.v0y, s.v1x, s.v1y, s.k0, s.k1, s.d00 * sg, s.d01 * sg, s.d11 * sg, s.den * sg, tx0, ty0, tx1, ty1);
        counts[0] += (uint64_t) (x1 - x0 + 1) * (uint64_t) (y1 - y0 + 1);
        if(tx0 <= tx1 && ty0 <= ty1) counts[1] += (uint64_t) (tx1 - tx0 + 1) * (uint64_t) (ty1 - ty0 + 1);
        for(int x = x0; x <= x1; x++)
            for(int y = y0; y <= y1; y++)
            {
                float nv, nw, v, w, u, z;
                gel::bary_numerators(s, gel::i2f(x), gel::i2f(y), nv, nw);
                const bool inside = gel::bary_inside(s, nv, nw, v, w, u, z);
                counts[2] += inside;
                const bool kept = x >= tx0 && x <= tx1 && y >= ty0 && y <= ty1;
                if(inside && !kept) wrong++;
            }
    }
    return wrong;
}

/* Brute-force soundness of row_trim (gel_math.h), the band rasteriser's per-column row trimming: every triangle's clipped bbox is
 * cut into the rasteriser's 32-row tile segments; for every column of every segment the two end rows are evaluated with the row
 * loop's operations, row_trim says how many rows go at either end, and every row it removes is evaluated the reference's way
 * (tbarycenter + the >= 0 test, main.c:316-332, 352) and the kernels' way (the exact cheap rejection) -- none may be inside, all
 * must be rejected.  Returns the violations; counts[0] += rows tested without trimming, counts[1] += rows left, counts[2] += inside. */
extern "C" uint64_t emu_row_trim_soundness(const float* vew, int ntri, int xres, int yres, uint64_t* counts)
{
    uint64_t wrong = 0;
    for(int t = 0; t < ntri; t++)
    {
        const float* p = vew + 9 * (size_t) t;
        const gel::TriSetup s = gel::tri_setup(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
        const int x0 = s.x0 < 0 ? 0 : s.x0, y0 = s.y0 < 0 ? 0 : s.y0, x1 = s.x1 > xres - 1 ? xres - 1 : s.x1, y1 = s.y1 > yres - 1 ? yres - 1 : s.y1;
        const float ad = fabsf(s.den);
        if(x0 > x1 || y0 > y1 || !(ad > 0.0f)) continue;
        const float sg = s.den < 0.0f ? -1.0f : 1.0f;
        const float B = s.d00 * sg, C = s.d01 * sg, A = s.d11 * sg, D = s.den * sg;
        const float eps = ad <= 1e18f ? -1e-20f : -INFINITY, den_hi = gel::mul(D, 1.00001f);
        float ev, ew;
        gel::trim_slack(s.ax, s.ay, s.v0x, s.v0y, s.v1x, s.v1y, s.k0, s.k1, B, C, A, D, x0, y0, x1, y1, ev, ew);
        auto numer = [&](int x, int y, float& nv, float& nw)
        {
            const float v2x = gel::sub(gel::i2f(x), s.ax), cx0 = gel::mul(v2x, s.v0x), cx1 = gel::mul(v2x, s.v1x);
            const float v2y = gel::sub(gel::i2f(y), s.ay);
            const float d20 = gel::add(gel::add(cx0, gel::mul(v2y, s.v0y)), s.k0), d21 = gel::add(gel::add(cx1, gel::mul(v2y, s.v1y)), s.k1);
            nv = gel::sub(gel::mul(A, d20), gel::mul(C, d21)); nw = gel::sub(gel::mul(B, d21), gel::mul(C, d20));
        };
        for(int ty = y0 / 32; ty <= y1 / 32; ty++)
        {
            const int ya = y0 > ty * 32 ? y0 : ty * 32, yb = y1 < ty * 32 + 31 ? y1 : ty * 32 + 31, n = yb - ya;
            for(int x = x0; x <= x1; x++)
            {
                float nv0, nw0, nv1, nw1;
                numer(x, ya, nv0, nw0); numer(x, yb, nv1, nw1);
                int lo, hi;
                gel::row_trim(nv0, nw0, nv1, nw1, eps, den_hi, ev, ew, n, lo, hi);
                counts[0] += (uint64_t) (n + 1);
                if(lo + hi < n + 1) counts[1] += (uint64_t) (n + 1 - lo - hi);
                for(int y = ya; y <= yb; y++)
                {
                    float nv, nw, v, w, u, z, rv, rw;
                    gel::bary_numerators(s, gel::i2f(x), gel::i2f(y), rv, rw);
                    const bool inside = gel::bary_inside(s, rv, rw, v, w, u, z);
                    counts[2] += inside;
                    if(y - ya >= lo && yb - y >= hi) continue;               /* kept */
                    numer(x, y, nv, nw);
                    const bool rejected = nv < eps || nw < eps || gel::add(nv, nw) > den_hi;
                    if(inside || !rejected) wrong++;
                }
            }
        }
    }
    return wrong;
}
