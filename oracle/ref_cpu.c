/* TEST INFRASTRUCTURE (oracle) -- see ref_cpu.h.  Never linked into the product.
 *
 * A parameterised (resolution, view list) CPU restatement of the reference's per-frame path,
 * /root/reference/main.c:182-225 (vector math), :288-314 (viewport, perspective), :316-370 (barycentric
 * solve, shade, raster), :372-390 (view transforms), :413-417 (clear), :506-522 (camera basis + triangle
 * loop), and of its load-time flow :84-180, :227-286.  Every fp32 expression keeps the reference's
 * operand order and association; build with `-std=c99 -O2 -ffp-contract=off` (oracle/Makefile) so no
 * multiply-add is fused and nothing is re-associated.
 *
 * Written on flat float[3] arrays rather than the reference's by-value structs; the arithmetic performed
 * per element is the same sequence of IEEE binary32 operations.
 */
#define _POSIX_C_SOURCE 200809L
#include "ref_cpu.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- vector helpers: main.c:182-213 ------------------------------------------------------------ */

/* vdot, main.c:200-203: (ax*bx + ay*by) + az*bz, left to right */
static float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* vunit, main.c:205-213: v * (1.0f / sqrtf(v.v)) -- one reciprocal, three multiplies */
static void unit3(const float* v, float* o)
{
    const float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    o[0] = v[0] * inv; o[1] = v[1] * inv; o[2] = v[2] * inv;
}

/* vcross, main.c:188-192 */
static void cross3(const float* a, const float* b, float* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* ---- camera basis: main.c:506-512 -------------------------------------------------------------- */

void ref_view_basis(float xt, float yt, float basis[12])
{
    const float center[3] = { 0.0f, 0.0f, 0.0f };
    const float upward[3] = { 0.0f, 1.0f, 0.0f };
    float* x = basis; float* y = basis + 3; float* z = basis + 6; float* eye = basis + 9;
    float d[3], c[3];
    eye[0] = sinf(xt); eye[1] = sinf(yt); eye[2] = cosf(xt);          /* :509 */
    for(int k = 0; k < 3; k++) d[k] = eye[k] - center[k];             /* vsub, :510 */
    unit3(d, z);
    cross3(upward, z, c);                                             /* :511 */
    unit3(c, x);
    cross3(z, x, y);                                                  /* :512 */
}

/* ---- per-corner transform chain: main.c:372-390, 302-314, 288-300 ------------------------------ */

static void corner_view(const float* v, const float* n, const float* basis, int xres, int yres,
                        float* vew, float* nrm)
{
    const float* x = basis; const float* y = basis + 3; const float* z = basis + 6; const float* eye = basis + 9;
    /* tviewnrm :382-390, then tunit/vunit :205-219 */
    const float nv[3] = { dot3(n, x), dot3(n, y), dot3(n, z) };
    unit3(nv, nrm);
    /* tviewtri :372-380 */
    const float tx = dot3(v, x) - dot3(x, eye);
    const float ty = dot3(v, y) - dot3(y, eye);
    const float tz = dot3(v, z) - dot3(z, eye);
    /* tperspective :302-314 */
    const float c = 3.0f;
    const float zd = 1.0f - tz / c;
    const float px = tx / zd, py = ty / zd, pz = tz / zd;
    /* tviewport :288-300 */
    const float w = yres / 1.5f;
    const float h = yres / 1.5f;
    const float x0 = xres / 2.0f;
    const float y0 = yres / 4.0f;
    vew[0] = w * px + x0;
    vew[1] = h * py + y0;
    vew[2] = (pz + 1.0f) / 1.5f;
}

void ref_transform(const float* tv, const float* tn, int ntri, const float basis[12],
                   int xres, int yres, float* vew, float* nrm)
{
    for(int i = 0; i < ntri; i++)
        for(int k = 0; k < 3; k++)
            corner_view(tv + 9 * i + 3 * k, tn + 9 * i + 3 * k, basis, xres, yres, vew + 9 * i + 3 * k, nrm + 9 * i + 3 * k);
}

/* ---- pshade: main.c:334-340 -------------------------------------------------------------------- */

static uint32_t shade_texel(uint32_t p, int shading)
{
    const uint32_t r = ((p >> 16) * shading) >> 8;          /* no mask: relies on X byte == 0 */
    const uint32_t g = (((p >> 8) & 0xFF) * shading) >> 8;
    const uint32_t b = ((p & 0xFF) * shading) >> 8;
    return r << 16 | g << 8 | b;
}

/* ---- one triangle: tdraw main.c:342-370 with tbarycenter :316-332 inlined per pixel ------------ */

static int draw_triangle(const float* vew, const float* nrm, const float* tex3,
                         const uint32_t* tex, int tw, int th, int xres, int yres,
                         uint32_t* pixel, float* zbuff, RefCounters* cnt)
{
    const float* a = vew; const float* b = vew + 3; const float* c = vew + 6;
    const float lights[3] = { 0.0f, 0.0f, 1.0f };                     /* :508 */
    int x0 = (int) fminf(a[0], fminf(b[0], c[0]));                    /* :344-347, truncating casts */
    int y0 = (int) fminf(a[1], fminf(b[1], c[1]));
    int x1 = (int) fmaxf(a[0], fmaxf(b[0], c[0]));
    int y1 = (int) fmaxf(a[1], fmaxf(b[1], c[1]));
    int clipped = 0;
    if(x0 < 0) { x0 = 0; clipped = 1; }
    if(y0 < 0) { y0 = 0; clipped = 1; }
    if(x1 > xres - 1) { x1 = xres - 1; clipped = 1; }
    if(y1 > yres - 1) { y1 = yres - 1; clipped = 1; }
    /* per-triangle invariants of tbarycenter (same operations on the same operands every pixel) */
    const float v0[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] };    /* :319 */
    const float v1[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };    /* :320 */
    const float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1);   /* :322-324 */
    const float den = d00 * d11 - d01 * d01;                          /* :327-328 */
    const float varying[3] = { dot3(lights, nrm + 3), dot3(lights, nrm + 6), dot3(lights, nrm) };  /* :358 */
    for(int x = x0; x <= x1; x++)
    for(int y = y0; y <= y1; y++)
    {
        const float v2[3] = { (float) x - a[0], (float) y - a[1], 0.0f - a[2] };   /* :318, :321 */
        const float d20 = dot3(v2, v0), d21 = dot3(v2, v1);           /* :325-326 */
        const float v = (d11 * d20 - d01 * d21) / den;                /* :327 */
        const float w = (d00 * d21 - d01 * d20) / den;                /* :328 */
        const float u = 1.0f - v - w;                                 /* :329 */
        if(cnt) cnt->tested++;
        if(v >= 0.0f && w >= 0.0f && u >= 0.0f)                       /* :352 */
        {
            const float z = v * b[2] + w * c[2] + u * a[2];           /* :355 */
            if(cnt) cnt->inside++;
            if(z > zbuff[y + x * yres])                               /* :356 */
            {
                const float bc[3] = { v, w, u };
                int xx = (tw - 1) * (0.0f + (v * tex3[3] + w * tex3[6] + u * tex3[0]));         /* :360 */
                int yy = (th - 1) * (1.0f - (v * tex3[4] + w * tex3[7] + u * tex3[1]));         /* :361 */
                /* Outside [0,tw-1] x [0,th-1] the reference reads out of bounds (undefined).  Like the bbox clipping
                 * above this restatement defines the case the way the product does -- clamp and flag (bit 2) -- so
                 * that it stays a total function; in-domain inputs never come here. */
                if(xx < 0 || xx > tw - 1 || yy < 0 || yy > th - 1)
                {
                    clipped |= 2;
                    xx = xx < 0 ? 0 : xx > tw - 1 ? tw - 1 : xx;
                    yy = yy < 0 ? 0 : yy > th - 1 ? th - 1 : yy;
                }
                const float intensity = dot3(bc, varying);            /* :362 */
                const int shading = 0xFF * (intensity < 0.0f ? 0.0f : intensity > 1.0f ? 1.0f : intensity);
                if(cnt) cnt->zpass++;
                zbuff[y + x * yres] = z;                              /* :365 */
                pixel[y + x * yres] = shade_texel(tex[xx + yy * tw], shading);      /* :366 */
            }
        }
    }
    return clipped;
}

int ref_render(const float* tv, const float* tn, const float* tt, int ntri,
               const uint32_t* tex, int tw, int th, int xres, int yres, const float basis[12],
               uint32_t* pixel, float* zbuff, RefCounters* counters)
{
    const int size = xres * yres;
    int clipped = 0;
    if(counters) memset(counters, 0, sizeof *counters);
    for(int i = 0; i < size; i++) { zbuff[i] = -FLT_MAX; pixel[i] = 0x0; }          /* reset, :413-417 */
    for(int i = 0; i < ntri; i++)                                                    /* :513-522 */
    {
        float vew[9], nrm[9];
        for(int k = 0; k < 3; k++)
            corner_view(tv + 9 * i + 3 * k, tn + 9 * i + 3 * k, basis, xres, yres, vew + 3 * k, nrm + 3 * k);
        clipped |= draw_triangle(vew, nrm, tt + 9 * i, tex, tw, th, xres, yres, pixel, zbuff, counters);
    }
    if(counters)
        for(int i = 0; i < size; i++) counters->lit += zbuff[i] != -FLT_MAX;
    return clipped;
}

/* ---- checksums ---------------------------------------------------------------------------------- */

uint64_t ref_fnv1a64_words(const uint32_t* w, uint64_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for(uint64_t i = 0; i < n; i++) h = (h ^ w[i]) * 0x100000001b3ull;
    return h;
}

uint64_t ref_salted_sum(const uint32_t* w, uint64_t n)
{
    uint64_t s = 0;
    for(uint64_t i = 0; i < n; i++)
    {
        uint32_t h = w[i] ^ ((uint32_t) i * 0x9E3779B1u);
        h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
        s += h;
    }
    return s;
}

/* ---- frames-parallel batch ---------------------------------------------------------------------- */

typedef struct
{
    const float *tv, *tn, *tt; int ntri;
    const uint32_t* tex; int tw, th, xres, yres;
    const float* bases; int nviews, nthreads, tid;
    uint32_t* pixel_out; float* z_out; uint64_t *hash_pixel, *hash_z;
    int clipped;
}
Job;

static void* job_main(void* arg)
{
    Job* j = (Job*) arg;
    const size_t size = (size_t) j->xres * j->yres;
    uint32_t* px_priv = j->pixel_out ? NULL : (uint32_t*) malloc(size * sizeof(uint32_t));
    float* z_priv = j->z_out ? NULL : (float*) malloc(size * sizeof(float));
    for(int v = j->tid; v < j->nviews; v += j->nthreads)
    {
        uint32_t* px = j->pixel_out ? j->pixel_out + size * v : px_priv;
        float* zb = j->z_out ? j->z_out + size * v : z_priv;
        j->clipped |= ref_render(j->tv, j->tn, j->tt, j->ntri, j->tex, j->tw, j->th, j->xres, j->yres,
                                 j->bases + 12 * v, px, zb, NULL);
        if(j->hash_pixel) j->hash_pixel[v] = ref_salted_sum(px, size);
        if(j->hash_z) j->hash_z[v] = ref_salted_sum((const uint32_t*) (const void*) zb, size);
    }
    free(px_priv); free(z_priv);
    return NULL;
}

int ref_render_views(const float* tv, const float* tn, const float* tt, int ntri,
                     const uint32_t* tex, int tw, int th, int xres, int yres,
                     const float* bases, int nviews, int nthreads,
                     uint32_t* pixel_out, float* z_out, uint64_t* hash_pixel, uint64_t* hash_z,
                     double* seconds)
{
    if(nthreads < 1) nthreads = 1;
    if(nthreads > nviews) nthreads = nviews > 0 ? nviews : 1;
    Job* jobs = (Job*) calloc(nthreads, sizeof(Job));
    pthread_t* th_ids = (pthread_t*) calloc(nthreads, sizeof(pthread_t));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for(int t = 0; t < nthreads; t++)
    {
        const Job j = { tv, tn, tt, ntri, tex, tw, th, xres, yres, bases, nviews, nthreads, t,
                        pixel_out, z_out, hash_pixel, hash_z, 0 };
        jobs[t] = j;
        if(t > 0) pthread_create(&th_ids[t], NULL, job_main, &jobs[t]);
    }
    job_main(&jobs[0]);
    int clipped = jobs[0].clipped;
    for(int t = 1; t < nthreads; t++) { pthread_join(th_ids[t], NULL); clipped |= jobs[t].clipped; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if(seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    free(jobs); free(th_ids);
    return clipped;
}

/* ---- load-time flow: main.c:84-180 (OBJ text), :227-286 (soup expansion) ----------------------- */

typedef struct { float* p; int n, cap; } FloatTriples;
typedef struct { int* p; int n, cap; } IntNines;

static void push3(FloatTriples* a, const float* v)
{
    if(a->n == a->cap) a->p = (float*) realloc(a->p, sizeof(float) * 3 * (a->cap = a->cap ? a->cap * 2 : 128));
    memcpy(a->p + 3 * a->n++, v, 3 * sizeof(float));
}

int ref_load_obj(const char* path, float** tv_out, float** tn_out, float** tt_out)
{
    FILE* f = fopen(path, "r");                                       /* oload, :460-469 */
    if(!f) return -1;
    FloatTriples vs = { 0 }, ns = { 0 }, ts = { 0 };
    IntNines fs = { 0 };
    char* line = NULL; size_t cap = 0;
    /* oparse dispatch order, :142-174: "vn", then "vt", then any other 'v', then 'f' */
    while(getline(&line, &cap, f) >= 0)
    {
        float v[3] = { 0.0f, 0.0f, 0.0f };
        if(line[0] == 'v' && line[1] == 'n') { sscanf(line, "vn %f %f %f", v, v + 1, v + 2); push3(&ns, v); }
        else if(line[0] == 'v' && line[1] == 't') { sscanf(line, "vt %f %f %f", v, v + 1, v + 2); push3(&ts, v); }
        else if(line[0] == 'v') { sscanf(line, "v %f %f %f", v, v + 1, v + 2); push3(&vs, v); }
        else if(line[0] == 'f')
        {
            int q[9];   /* file order: va ta na  vb tb nb  vc tc nc, :167 */
            sscanf(line, "f %d/%d/%d %d/%d/%d %d/%d/%d", q, q + 1, q + 2, q + 3, q + 4, q + 5, q + 6, q + 7, q + 8);
            if(fs.n == fs.cap) fs.p = (int*) realloc(fs.p, sizeof(int) * 9 * (fs.cap = fs.cap ? fs.cap * 2 : 128));
            for(int k = 0; k < 9; k++) fs.p[9 * fs.n + k] = q[k] - 1;  /* 1-based -> 0-based, :168-172 */
            fs.n++;
        }
    }
    free(line);
    fclose(f);
    /* vmaxlen :233-240 and the int-truncated scale of tvgen :244 */
    float maxlen = 0.0f;
    for(int i = 0; i < vs.n; i++)
    {
        const float* v = vs.p + 3 * i;
        const float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if(len > maxlen) maxlen = len;
    }
    const int scale = maxlen;
    const float inv = 1.0f / scale;                                   /* :253 */
    const int nt = fs.n;
    float* tv = (float*) malloc(sizeof(float) * 9 * (nt ? nt : 1));
    float* tn = (float*) malloc(sizeof(float) * 9 * (nt ? nt : 1));
    float* tt = (float*) malloc(sizeof(float) * 9 * (nt ? nt : 1));
    for(int i = 0; i < nt; i++)
        for(int k = 0; k < 3; k++)
        {
            const int* q = fs.p + 9 * i + 3 * k;                      /* (v, t, n) of corner k */
            for(int e = 0; e < 3; e++)
            {
                tv[9 * i + 3 * k + e] = vs.p[3 * q[0] + e] * inv;     /* tvgen + tmul, :242-256, 221-225 */
                tt[9 * i + 3 * k + e] = ts.p[3 * q[1] + e];           /* ttgen, :273-286 */
                tn[9 * i + 3 * k + e] = ns.p[3 * q[2] + e];           /* tngen, :258-271 */
            }
        }
    free(vs.p); free(ns.p); free(ts.p); free(fs.p);
    *tv_out = tv; *tn_out = tn; *tt_out = tt;
    return nt;
}

int ref_load_bmp(const char* path, uint32_t** xrgb, int* w_out, int* h_out)
{
    FILE* f = fopen(path, "rb");
    if(!f) return -1;
    unsigned char hdr[54];
    if(fread(hdr, 1, 54, f) != 54 || hdr[0] != 'B' || hdr[1] != 'M') { fclose(f); return -2; }
    const uint32_t off = hdr[10] | hdr[11] << 8 | hdr[12] << 16 | (uint32_t) hdr[13] << 24;
    const int32_t w = (int32_t) (hdr[18] | hdr[19] << 8 | hdr[20] << 16 | (uint32_t) hdr[21] << 24);
    const int32_t hs = (int32_t) (hdr[22] | hdr[23] << 8 | hdr[24] << 16 | (uint32_t) hdr[25] << 24);
    const int bpp = hdr[28] | hdr[29] << 8;
    const uint32_t comp = hdr[30] | hdr[31] << 8 | hdr[32] << 16 | (uint32_t) hdr[33] << 24;
    if(bpp != 24 || comp != 0 || w <= 0 || hs == 0) { fclose(f); return -3; }
    const int h = hs < 0 ? -hs : hs;
    const size_t rowbytes = ((size_t) w * 3 + 3) & ~(size_t) 3;
    unsigned char* row = (unsigned char*) malloc(rowbytes);
    uint32_t* px = (uint32_t*) malloc((size_t) 4 * w * h);
    fseek(f, (long) off, SEEK_SET);
    for(int r = 0; r < h; r++)
    {
        if(fread(row, 1, rowbytes, f) != rowbytes) { free(row); free(px); fclose(f); return -4; }
        const int y = hs < 0 ? r : h - 1 - r;
        for(int x = 0; x < w; x++)
            px[(size_t) y * w + x] = (uint32_t) row[3 * x + 2] << 16 | (uint32_t) row[3 * x + 1] << 8 | row[3 * x];
    }
    free(row); fclose(f);
    *xrgb = px; *w_out = w; *h_out = h;
    return 0;
}

void ref_free(void* p) { free(p); }
