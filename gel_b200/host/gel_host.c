/* gel_host.c -- see gel_host.h.  From-scratch host flow; accepts the same OBJ grammar as the reference's
 * oparse (main.c:129-180): lines dispatched on "vn", "vt", other 'v', 'f'; three floats per vertex line;
 * faces are exactly `f v/t/n v/t/n v/t/n` with 1-based indices.  Numbers go through strtof/strtol, the
 * conversions glibc's sscanf("%f"/"%d") itself performs, so parsed values are bit-identical.
 * Differences, all in territory where the reference has undefined behaviour: a vertex line with fewer than
 * three numbers yields zeros for the missing ones (the reference leaves stale stack values), and an index
 * outside the arrays is an error (the reference reads out of bounds).
 */
#define _DEFAULT_SOURCE
#include "gel_host.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <stdint.h>

typedef struct { float* p; size_t n, cap; } FVec;
typedef struct { int* p; size_t n, cap; } IVec;

static int fvec_push3(FVec* v, const float* x)
{
    if(v->n + 3 > v->cap)
    {
        const size_t cap = v->cap ? v->cap * 2 : 384;
        float* q = (float*) realloc(v->p, cap * sizeof(float));
        if(!q) return -1;
        v->p = q; v->cap = cap;
    }
    memcpy(v->p + v->n, x, 3 * sizeof(float));
    v->n += 3;
    return 0;
}

static int ivec_push9(IVec* v, const int* x)
{
    if(v->n + 9 > v->cap)
    {
        const size_t cap = v->cap ? v->cap * 2 : 1152;
        int* q = (int*) realloc(v->p, cap * sizeof(int));
        if(!q) return -1;
        v->p = q; v->cap = cap;
    }
    memcpy(v->p + v->n, x, 9 * sizeof(int));
    v->n += 9;
    return 0;
}

static void scan3f(const char* s, float* out)
{
    out[0] = out[1] = out[2] = 0.0f;
    for(int k = 0; k < 3; k++)
    {
        char* end;
        const float f = strtof(s, &end);
        if(end == s) return;
        out[k] = f;
        s = end;
    }
}

/* `f %d/%d/%d %d/%d/%d %d/%d/%d` -> q[9] in file order; returns the number of integers converted */
static int scan_face(const char* s, int* q)
{
    int got = 0;
    for(int k = 0; k < 9; k++)
    {
        char* end;
        const long v = strtol(s, &end, 10);
        if(end == s) break;
        q[k] = (int) v; got++;
        s = end;
        if(k % 3 != 2) { if(*s != '/') break; s++; }
    }
    return got;
}

/* one chunk of the file: whole lines in [begin, end) */
enum { PARSE_MAX_THREADS = 32 };
typedef struct { char* begin; char* end; FVec vs, ns, ts; IVec fs; int rc; } ParseJob;

static void* parse_chunk(void* arg)
{
    ParseJob* j = (ParseJob*) arg;
    int rc = 0;
    for(char* line = j->begin; line < j->end && rc == 0; )
    {
        char* nl = (char*) memchr(line, '\n', (size_t) (j->end - line));
        if(nl) *nl = '\0';
        float v[3];
        if(line[0] == 'v' && line[1] == 'n') { scan3f(line + 2, v); rc = fvec_push3(&j->ns, v); }
        else if(line[0] == 'v' && line[1] == 't') { scan3f(line + 2, v); rc = fvec_push3(&j->ts, v); }
        else if(line[0] == 'v') { scan3f(line + 1, v); rc = fvec_push3(&j->vs, v); }
        else if(line[0] == 'f')
        {
            int q[9];
            if(scan_face(line + 1, q) == 9)
            {
                /* file order v/t/n per corner -> the reference's Face { va,vb,vc, ta,tb,tc, na,nb,nc }, 0-based (main.c:165-170) */
                const int face[9] = { q[0] - 1, q[3] - 1, q[6] - 1, q[1] - 1, q[4] - 1, q[7] - 1, q[2] - 1, q[5] - 1, q[8] - 1 };
                rc = ivec_push9(&j->fs, face);
            }
            else rc = -2;
        }
        if(!nl) break;
        line = nl + 1;
    }
    j->rc = rc;
    return NULL;
}

typedef struct
{
    const float *vs, *ns, *ts; const int* fs; int nv, nvt, nvn; float inv;
    float *tv, *tn, *tt; int first, last, rc;
}
ExpandJob;

static void* expand_range(void* arg)
{
    ExpandJob* j = (ExpandJob*) arg;
    for(int i = j->first; i < j->last; i++)
        for(int k = 0; k < 3; k++)
        {
            const int* face = j->fs + 9 * (size_t) i;
            const int qv = face[k], qt = face[3 + k], qn = face[6 + k];   /* v, t, n of corner k */
            if(qv < 0 || qv >= j->nv || qt < 0 || qt >= j->nvt || qn < 0 || qn >= j->nvn) { j->rc = -2; return NULL; }
            for(int e = 0; e < 3; e++)
            {
                j->tv[9 * (size_t) i + 3 * k + e] = j->vs[3 * qv + e] * j->inv;
                j->tt[9 * (size_t) i + 3 * k + e] = j->ts[3 * qt + e];
                j->tn[9 * (size_t) i + 3 * k + e] = j->ns[3 * qn + e];
            }
        }
    return NULL;
}

int gel_obj_parse(const char* path, GelObj* out)
{
    memset(out, 0, sizeof *out);
    FILE* f = fopen(path, "r");
    if(!f) return -1;
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* text = (char*) malloc((size_t) size + 1);
    if(!text) { fclose(f); return -3; }
    const size_t got = fread(text, 1, (size_t) size, f);
    fclose(f);
    text[got] = '\0';

    /* Large files are cut into chunks at line boundaries and parsed by several threads; the
     * pieces are joined in file order, so the arrays are the ones a single pass produces. */
    int nthreads = 1;
    if(got >= ((size_t) 4 << 20))
    {
        const char* env = getenv("GEL_PARSE_THREADS");
        long want = env ? strtol(env, NULL, 10) : sysconf(_SC_NPROCESSORS_ONLN);
        nthreads = want < 1 ? 1 : want > PARSE_MAX_THREADS ? PARSE_MAX_THREADS : (int) want;
    }
    ParseJob jobs[PARSE_MAX_THREADS];
    pthread_t tid[PARSE_MAX_THREADS];
    memset(jobs, 0, sizeof jobs);
    char* cut = text;
    for(int t = 0; t < nthreads; t++)
    {
        char* end = t == nthreads - 1 ? text + got : text + got / (size_t) nthreads * (size_t) (t + 1);
        if(end < cut) end = cut;
        if(t != nthreads - 1)
        {
            char* nl = (char*) memchr(end, '\n', (size_t) (text + got - end));      /* a chunk ends after a newline */
            end = nl ? nl + 1 : text + got;
        }
        jobs[t].begin = cut; jobs[t].end = end;
        cut = end;
    }
    int spawned = 0;
    for(int t = 1; t < nthreads; t++)
    {
        if(pthread_create(&tid[t], NULL, parse_chunk, &jobs[t]) != 0) break;
        spawned = t;
    }
    for(int t = spawned + 1; t < nthreads; t++) parse_chunk(&jobs[t]);              /* threads that could not start: inline */
    parse_chunk(&jobs[0]);
    for(int t = 1; t <= spawned; t++) pthread_join(tid[t], NULL);
    free(text);

    FVec vs = { 0 }, ns = { 0 }, ts = { 0 };
    IVec fs = { 0 };
    int rc = 0;
    for(int t = 0; t < nthreads; t++) if(jobs[t].rc) rc = jobs[t].rc;
    if(rc == 0 && nthreads == 1) { vs = jobs[0].vs; ns = jobs[0].ns; ts = jobs[0].ts; fs = jobs[0].fs; memset(&jobs[0], 0, sizeof jobs[0]); }
    else if(rc == 0)
    {
        size_t nvs = 0, nns = 0, nts = 0, nfs = 0;
        for(int t = 0; t < nthreads; t++) { nvs += jobs[t].vs.n; nns += jobs[t].ns.n; nts += jobs[t].ts.n; nfs += jobs[t].fs.n; }
        vs.p = (float*) malloc(sizeof(float) * (nvs ? nvs : 1)); ns.p = (float*) malloc(sizeof(float) * (nns ? nns : 1));
        ts.p = (float*) malloc(sizeof(float) * (nts ? nts : 1)); fs.p = (int*) malloc(sizeof(int) * (nfs ? nfs : 1));
        if(!vs.p || !ns.p || !ts.p || !fs.p) rc = -3;
        for(int t = 0; t < nthreads && rc == 0; t++)
        {
            memcpy(vs.p + vs.n, jobs[t].vs.p, sizeof(float) * jobs[t].vs.n); vs.n += jobs[t].vs.n;
            memcpy(ns.p + ns.n, jobs[t].ns.p, sizeof(float) * jobs[t].ns.n); ns.n += jobs[t].ns.n;
            memcpy(ts.p + ts.n, jobs[t].ts.p, sizeof(float) * jobs[t].ts.n); ts.n += jobs[t].ts.n;
            memcpy(fs.p + fs.n, jobs[t].fs.p, sizeof(int) * jobs[t].fs.n); fs.n += jobs[t].fs.n;
        }
    }
    for(int t = 0; t < nthreads; t++) { free(jobs[t].vs.p); free(jobs[t].ns.p); free(jobs[t].ts.p); free(jobs[t].fs.p); }
    if(rc) { free(vs.p); free(ns.p); free(ts.p); free(fs.p); return rc == -2 ? -2 : -3; }

    out->v = vs.p; out->vn = ns.p; out->vt = ts.p; out->faces = fs.p;
    out->nv = (int) (vs.n / 3); out->nvn = (int) (ns.n / 3); out->nvt = (int) (ts.n / 3); out->nfaces = (int) (fs.n / 9);
    out->parse_threads = nthreads;
    return 0;
}

void gel_obj_free(GelObj* o)
{
    free(o->v); free(o->vt); free(o->vn); free(o->faces);
    memset(o, 0, sizeof *o);
}

int gel_obj_expand(const GelObj* obj, GelMesh* out)
{
    memset(out, 0, sizeof *out);
    int rc = 0;
    pthread_t tid[PARSE_MAX_THREADS];
    const int nthreads = obj->parse_threads > 0 ? obj->parse_threads : 1;
    const FVec vs = { obj->v, 0, 0 }, ns = { obj->vn, 0, 0 }, ts = { obj->vt, 0, 0 };
    const IVec fs = { obj->faces, 0, 0 };
    const int nv = obj->nv, nvn = obj->nvn, nvt = obj->nvt, nt = obj->nfaces;
    /* vmaxlen (main.c:233-240) and tvgen's integer scale (main.c:244) */
    float maxlen = 0.0f;
    for(int i = 0; i < nv; i++)
    {
        const float* p = vs.p + 3 * i;
        const float len = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        if(len > maxlen) maxlen = len;
    }
    const int scale = (int) maxlen;
    if(nt > 0 && scale == 0) return -4;
    const float inv = nt > 0 ? 1.0f / scale : 1.0f;
    const size_t soup_tris = nt > 0 ? (size_t) nt : 1;
    out->tv = (float*) malloc(sizeof(float) * 9 * soup_tris);
    out->tn = (float*) malloc(sizeof(float) * 9 * soup_tris);
    out->tt = (float*) malloc(sizeof(float) * 9 * soup_tris);
    if(!out->tv || !out->tn || !out->tt) rc = -3;
    if(rc == 0)
    {
        /* soup expansion (tvgen / ttgen / tngen, main.c:242-286): triangle ranges are independent */
        ExpandJob ej[PARSE_MAX_THREADS];
        const int nexp = nt >= (1 << 16) ? nthreads : 1;
        for(int t = 0; t < nexp; t++)
        {
            ExpandJob e = { vs.p, ns.p, ts.p, fs.p, nv, nvt, nvn, inv, out->tv, out->tn, out->tt,
                            (int) ((long long) nt * t / nexp), (int) ((long long) nt * (t + 1) / nexp), 0 };
            ej[t] = e;
        }
        int started = 0;
        for(int t = 1; t < nexp; t++)
        {
            if(pthread_create(&tid[t], NULL, expand_range, &ej[t]) != 0) break;
            started = t;
        }
        for(int t = started + 1; t < nexp; t++) expand_range(&ej[t]);
        expand_range(&ej[0]);
        for(int t = 1; t <= started; t++) pthread_join(tid[t], NULL);
        for(int t = 0; t < nexp; t++) if(ej[t].rc) rc = ej[t].rc;
    }
    if(rc) { gel_mesh_free(out); return rc; }
    out->ntri = nt; out->nv = nv; out->nvt = nvt; out->nvn = nvn;
    return 0;
}

int gel_obj_load(const char* path, GelMesh* out)
{
    GelObj obj;
    memset(out, 0, sizeof *out);
    int rc = gel_obj_parse(path, &obj);
    if(rc == 0) rc = gel_obj_expand(&obj, out);
    gel_obj_free(&obj);
    return rc;
}

void gel_mesh_free(GelMesh* m)
{
    free(m->tv); free(m->tn); free(m->tt);
    memset(m, 0, sizeof *m);
}

static uint32_t le32(const unsigned char* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t) p[3] << 24; }

/* Uncompressed Windows BMP -> XRGB8888 top-down with the X byte 0, i.e. what the reference gets from
 * IMG_Load + SDL_ConvertSurface(RGB888) (main.c:473-480): 24-bit BGR, 32-bit BGRX/BGRA (alpha dropped by the
 * conversion) and 8-bit palettised (palette entries B, G, R, -).  Every size read from the file is validated against
 * the file length before anything is allocated or read. */
int gel_bmp_load(const char* path, GelTexture* out)
{
    memset(out, 0, sizeof *out);
    FILE* f = fopen(path, "rb");
    if(!f) return -1;
    unsigned char hdr[54];
    if(fread(hdr, 1, 54, f) != 54 || hdr[0] != 'B' || hdr[1] != 'M') { fclose(f); return -2; }
    if(fseek(f, 0, SEEK_END) != 0) { fclose(f); return -4; }
    const long flen = ftell(f);
    const uint32_t data_off = le32(hdr + 10), dib = le32(hdr + 14), comp = le32(hdr + 30);
    const int32_t w = (int32_t) le32(hdr + 18), hs = (int32_t) le32(hdr + 22);
    const int bpp = hdr[28] | hdr[29] << 8;
    /* BI_RGB, or BI_BITFIELDS on 32 bits (the masks SDL writes: the layout is still B, G, R, A in memory) */
    const int layout_ok = comp == 0 || (comp == 3 && bpp == 32);
    if(dib < 40 || !layout_ok || (bpp != 8 && bpp != 24 && bpp != 32) || w <= 0 || hs == 0 || hs == INT32_MIN) { fclose(f); return -2; }
    const size_t h = (size_t) (hs < 0 ? -(int64_t) hs : hs);
    if((size_t) w > (size_t) 1 << 15 || h > (size_t) 1 << 15) { fclose(f); return -2; }          /* 32768^2 texels at most */
    const size_t stride = ((size_t) w * (size_t) (bpp / 8) + 3) & ~(size_t) 3;
    if(flen < 54 || (size_t) data_off < 14 + (size_t) dib || (size_t) data_off > (size_t) flen || stride * h > (size_t) flen - data_off) { fclose(f); return -4; }
    uint32_t palette[256];
    memset(palette, 0, sizeof palette);
    if(bpp == 8)
    {
        uint32_t ncol = le32(hdr + 46);
        if(ncol == 0 || ncol > 256) ncol = 256;
        const size_t pal_off = 14 + (size_t) dib;
        if(pal_off + 4 * (size_t) ncol > (size_t) data_off) ncol = (uint32_t) (((size_t) data_off - pal_off) / 4);
        unsigned char quad[4];
        if(fseek(f, (long) pal_off, SEEK_SET) != 0) { fclose(f); return -4; }
        for(uint32_t k = 0; k < ncol; k++)
        {
            if(fread(quad, 1, 4, f) != 4) { fclose(f); return -4; }
            palette[k] = (uint32_t) quad[2] << 16 | (uint32_t) quad[1] << 8 | quad[0];
        }
    }
    unsigned char* row = (unsigned char*) malloc(stride);
    uint32_t* px = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) w * h);
    if(!row || !px) { free(row); free(px); fclose(f); return -3; }
    if(fseek(f, (long) data_off, SEEK_SET) != 0) { free(row); free(px); fclose(f); return -4; }
    for(size_t r = 0; r < h; r++)
    {
        if(fread(row, 1, stride, f) != stride) { free(row); free(px); fclose(f); return -4; }
        uint32_t* dst = px + (hs < 0 ? r : h - 1 - r) * (size_t) w;    /* file is bottom-up unless height < 0 */
        if(bpp == 24)
            for(int32_t x = 0; x < w; x++) dst[x] = (uint32_t) row[3 * x + 2] << 16 | (uint32_t) row[3 * x + 1] << 8 | row[3 * x];   /* X byte = 0 */
        else if(bpp == 32)
            for(int32_t x = 0; x < w; x++) dst[x] = (uint32_t) row[4 * x + 2] << 16 | (uint32_t) row[4 * x + 1] << 8 | row[4 * x];
        else
            for(int32_t x = 0; x < w; x++) dst[x] = palette[row[x]];
    }
    free(row);
    fclose(f);
    out->pixels = px; out->w = w; out->h = (int) h;
    return 0;
}

void gel_texture_free(GelTexture* t) { free(t->pixels); memset(t, 0, sizeof *t); }

void gel_view_basis(float xt, float yt, float basis[12])
{
    float* x = basis; float* y = basis + 3; float* z = basis + 6; float* eye = basis + 9;
    eye[0] = sinf(xt); eye[1] = sinf(yt); eye[2] = cosf(xt);                       /* main.c:509 */
    /* z = vunit(vsub(eye, center)), center = 0                                       main.c:510 */
    const float dx = eye[0] - 0.0f, dy = eye[1] - 0.0f, dz = eye[2] - 0.0f;
    const float iz = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    z[0] = dx * iz; z[1] = dy * iz; z[2] = dz * iz;
    /* x = vunit(vcross(upward, z)), upward = (0,1,0)                                 main.c:511 */
    const float ux = 0.0f, uy = 1.0f, uz = 0.0f;
    const float cx = uy * z[2] - uz * z[1], cy = uz * z[0] - ux * z[2], cz = ux * z[1] - uy * z[0];
    const float ix = 1.0f / sqrtf(cx * cx + cy * cy + cz * cz);
    x[0] = cx * ix; x[1] = cy * ix; x[2] = cz * ix;
    /* y = vcross(z, x)                                                                main.c:512 */
    y[0] = z[1] * x[2] - z[2] * x[1];
    y[1] = z[2] * x[0] - z[0] * x[2];
    y[2] = z[0] * x[1] - z[1] * x[0];
}

void gel_input_step(float* xt, float* yt, int dx, int dy)
{
    const float sens = 0.005f;
    *xt -= sens * dx;
    *yt += sens * dy;
}

uint64_t gel_fnv1a64_words(const uint32_t* w, uint64_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for(uint64_t i = 0; i < n; i++) h = (h ^ w[i]) * 0x100000001b3ull;
    return h;
}

void gel_upright(const uint32_t* canvas, int xres, int yres, uint32_t* upright)
{
    for(int wy = 0; wy < yres; wy++)
        for(int wx = 0; wx < xres; wx++)
            upright[(size_t) wy * xres + wx] = canvas[(size_t) (yres - 1 - wy) + (size_t) wx * yres];
}

int gel_write_ppm(const char* path, const uint32_t* canvas, int xres, int yres)
{
    FILE* f = fopen(path, "wb");
    if(!f) return -1;
    fprintf(f, "P6\n%d %d\n255\n", xres, yres);
    unsigned char* row = (unsigned char*) malloc((size_t) 3 * xres);
    for(int wy = 0; wy < yres; wy++)
    {
        for(int wx = 0; wx < xres; wx++)
        {
            const uint32_t p = canvas[(size_t) (yres - 1 - wy) + (size_t) wx * yres];
            row[3 * wx] = (unsigned char) (p >> 16); row[3 * wx + 1] = (unsigned char) (p >> 8); row[3 * wx + 2] = (unsigned char) p;
        }
        fwrite(row, 1, (size_t) 3 * xres, f);
    }
    free(row);
    fclose(f);
    return 0;
}
