/* gel.c -- headless `gel`: the reference's host flow (main.c:486-531) with the per-frame render path
 * (main.c:505-522) handed to the GPU through the C ABI in include/gelcu.h.
 *
 *   reference main()                          here
 *   ----------------------------------------  -------------------------------------------------------
 *   oload / oparse                            gel_obj_parse                       (gel_host.c)
 *   tvgen / ttgen / tngen                     gelcu_set_mesh_indexed: on the device (--soups: gel_obj_expand on the host
 *                                             + gelcu_set_mesh, the reference's own three soups)
 *   sload                                     gel_bmp_load
 *   ssetup(800, 600)                          gelcu_create(device, xres, yres)    default 800x600, --res
 *   iinit / ipump (mouse)                     scripted input: --mouse DX,DY per frame, or --sweep N
 *   slock .. reset .. tdraw loop .. sunlock   gelcu_render(views...); --region: gelcu_render_region into reused frame slots
 *                                             (only each view's screen region crosses PCIe)
 *   schurn / spresent                         per-frame JSON line (FNV-1a-64, non-zero pixels), optional
 *                                             raw dump (--dump) or upright PPM (--ppm); with --sink rgb8 the
 *                                             un-rotation + 24-bit pack happen on the device
 *                                             (gelcu_render_rgb8) and --ppm writes the buffer as it arrives
 *   60 fps cap                                none (headless)
 *
 * Exit status and messages follow the reference: wrong argument count prints the usage line and returns
 * 1 (main.c:488-492); an unreadable OBJ prints "could not open <path>" and exits 1 (main.c:463-467).
 * There is no CPU render fallback: without a CUDA device the program fails.
 */
#define _DEFAULT_SOURCE
#include "gel_host.h"
#include "gelcu.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct
{
    int device, first, count, xres, yres, batch, readback, sink, region;
    const GelObj* obj; const GelMesh* mesh; const GelTexture* tex; const gelcu_view* views;
    gelcu_rect* rects;       /* region mode: one per frame slot, carried from chunk to chunk */
    uint32_t* pixels;        /* count frames (readback); with sink: count upright 24-bit frames */
    uint64_t* hashes;        /* 2 per view */
    float device_ms; double wall_s; int rc; char err[512];
    gelcu_stats stats;
}
Shard;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void* shard_main(void* arg)
{
    Shard* s = (Shard*) arg;
    gelcu_ctx* ctx = NULL;
    s->rc = gelcu_create(&ctx, s->device, s->xres, s->yres);
    if(s->rc == 0)
        s->rc = s->mesh ? gelcu_set_mesh(ctx, s->mesh->tv, s->mesh->tn, s->mesh->tt, s->mesh->ntri)
                        : gelcu_set_mesh_indexed(ctx, s->obj->v, s->obj->nv, s->obj->vt, s->obj->nvt, s->obj->vn, s->obj->nvn, s->obj->faces, s->obj->nfaces);
    if(s->rc == 0) s->rc = gelcu_set_texture(ctx, s->tex->pixels, s->tex->w, s->tex->h);
    if(s->rc == 0 && s->batch > 0) s->rc = gelcu_set_option(ctx, "batch_views", s->batch);
    if(s->rc == 0)
    {
        const double t0 = now_s();
        if(s->region && s->readback) s->rc = gelcu_render_region(ctx, s->views + s->first, s->count, s->pixels, NULL, s->rects, s->sink, s->hashes, &s->device_ms);
        else if(s->sink && s->readback) s->rc = gelcu_render_rgb8(ctx, s->views + s->first, s->count, (uint8_t*) s->pixels, s->hashes, &s->device_ms);
        else s->rc = gelcu_render(ctx, s->views + s->first, s->count, s->readback ? s->pixels : NULL, NULL, s->hashes, &s->device_ms);
        s->wall_s = now_s() - t0;
        gelcu_get_stats(ctx, &s->stats);
    }
    if(s->rc != 0) snprintf(s->err, sizeof s->err, "%s", gelcu_last_error());
    gelcu_destroy(ctx);
    return NULL;
}

int main(int argc, char* argv[])
{
    const char* positional[2] = { NULL, NULL };
    int npos = 0, xres = 800, yres = 600, frames = 1, dx = 0, dy = 0, sweep = 0, gpus = 1, batch = 0, readback = 1, sink = 0, region = 0, soups = 0;
    const char* dump_path = NULL; const char* ppm_prefix = NULL;
    for(int i = 1; i < argc; i++)
    {
        if(!strcmp(argv[i], "--res") && i + 1 < argc) { if(sscanf(argv[++i], "%dx%d", &xres, &yres) != 2) npos = 99; }
        else if(!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
        else if(!strcmp(argv[i], "--mouse") && i + 1 < argc) { if(sscanf(argv[++i], "%d,%d", &dx, &dy) != 2) npos = 99; }
        else if(!strcmp(argv[i], "--sweep") && i + 1 < argc) sweep = atoi(argv[++i]);
        else if(!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if(!strcmp(argv[i], "--batch") && i + 1 < argc) batch = atoi(argv[++i]);
        else if(!strcmp(argv[i], "--dump") && i + 1 < argc) dump_path = argv[++i];
        else if(!strcmp(argv[i], "--ppm") && i + 1 < argc) ppm_prefix = argv[++i];
        else if(!strcmp(argv[i], "--no-readback")) readback = 0;
        else if(!strcmp(argv[i], "--region")) region = 1;
        else if(!strcmp(argv[i], "--soups")) soups = 1;
        else if(!strcmp(argv[i], "--sink") && i + 1 < argc) { if(!strcmp(argv[++i], "rgb8")) sink = 1; else npos = 99; }
        else if(argv[i][0] == '-' && argv[i][1] == '-') npos = 99;
        else if(npos < 2) positional[npos++] = argv[i];
        else npos = 99;
    }
    if(npos != 2)
    {
        puts("args: path/to/obj path/to/bmp");
        puts("      [--res WxH] [--frames N] [--mouse DX,DY] [--sweep N] [--gpus G] [--batch B]");
        puts("      [--dump frames.raw] [--ppm prefix] [--no-readback] [--sink rgb8] [--region] [--soups]");
        return 1;
    }
    GelObj obj; GelMesh mesh; GelTexture tex;
    memset(&mesh, 0, sizeof mesh);
    int rc = gel_obj_parse(positional[0], &obj);
    if(rc == -1) { printf("could not open %s\n", positional[0]); exit(1); }
    if(rc == 0 && soups) rc = gel_obj_expand(&obj, &mesh);              /* the reference's host-side tvgen / ttgen / tngen */
    if(rc != 0) { printf("could not parse %s (error %d)\n", positional[0], rc); exit(1); }
    rc = gel_bmp_load(positional[1], &tex);
    if(rc != 0) { printf("could not load %s (error %d; need an uncompressed 8-, 24- or 32-bit BMP)\n", positional[1], rc); exit(1); }

    /* scripted input -> one basis per frame (main.c:501, 506-512) */
    const int nviews = sweep > 0 ? sweep : frames;
    if(nviews <= 0 || gpus <= 0) { puts("nothing to render"); return 1; }
    gelcu_view* views = (gelcu_view*) malloc(sizeof(gelcu_view) * (size_t) nviews);
    float xt = 0.0f, yt = 0.0f;
    for(int k = 0; k < nviews; k++)
    {
        if(sweep > 0) { xt = (float) (2.0 * M_PI * k / sweep); yt = 0.0f; }
        gel_view_basis(xt, yt, (float*) &views[k]);
        if(sweep == 0) gel_input_step(&xt, &yt, dx, dy);
    }
    const int ndev = gelcu_device_count();
    if(ndev <= 0) { printf("no CUDA device: %s\n", gelcu_last_error()); exit(1); }
    if(gpus > ndev) gpus = ndev;
    if(gpus > nviews) gpus = nviews;

    const size_t frame = (size_t) xres * yres;
    const size_t frame_bytes = frame * (sink ? 3 : 4);                   /* sink: upright 24-bit frames come back */
    FILE* dump = dump_path ? fopen(dump_path, "wb") : NULL;
    /* frames come back in chunks so that long sweeps do not need nviews frames of host memory */
    const int chunk = readback ? (int) (((size_t) 1 << 30) / frame_bytes > 0 ? ((size_t) 1 << 30) / frame_bytes : 1) : nviews;
    uint64_t* hashes = (uint64_t*) calloc((size_t) 2 * nviews, sizeof(uint64_t));
    /* frame slots: one allocation reused by every chunk (with --region the slots carry their rectangles from chunk to chunk,
     * so a slot is only reset where the previous view drew and the new one does not) */
    const int slots = nviews < chunk * gpus ? nviews : chunk * gpus;
    uint32_t* pixels = NULL;
    gelcu_rect* rects = (gelcu_rect*) calloc((size_t) slots, sizeof(gelcu_rect));
    if(readback && gelcu_host_alloc((void**) &pixels, frame_bytes * (size_t) slots) != 0) { printf("host alloc failed: %s\n", gelcu_last_error()); exit(1); }
    if(readback && region) { memset(pixels, 0, frame_bytes * (size_t) slots); for(int k = 0; k < slots; k++) { rects[k].x0 = 0; rects[k].y0 = 0; rects[k].x1 = -1; rects[k].y1 = -1; } }
    double dev_ms_max_sum = 0.0, wall_sum = 0.0;
    uint64_t launches = 0;
    int status = 0;
    for(int base = 0; base < nviews; base += chunk * gpus)
    {
        const int n = nviews - base < chunk * gpus ? nviews - base : chunk * gpus;
        Shard shards[64]; pthread_t th[64];
        const int g_used = gpus > 64 ? 64 : gpus;
        for(int g = 0; g < g_used; g++)
        {
            const int lo = (int) ((long long) n * g / g_used), hi = (int) ((long long) n * (g + 1) / g_used);
            Shard s = { g, base + lo, hi - lo, xres, yres, batch, readback, sink, region, &obj, soups ? &mesh : NULL, &tex, views, rects + lo,
                        readback ? (uint32_t*) ((uint8_t*) pixels + frame_bytes * lo) : NULL, hashes + 2 * (size_t) (base + lo), 0.0f, 0.0, 0, "", { 0 } };
            shards[g] = s;
            pthread_create(&th[g], NULL, shard_main, &shards[g]);
        }
        float ms_max = 0.0f; double wall_max = 0.0;
        for(int g = 0; g < g_used; g++)
        {
            pthread_join(th[g], NULL);
            if(shards[g].rc < 0) { printf("gpu %d: %s\n", g, shards[g].err); exit(1); }
            if(shards[g].rc > 0) { fprintf(stderr, "gpu %d warning: %s\n", g, shards[g].err); status = 2; }
            if(shards[g].device_ms > ms_max) ms_max = shards[g].device_ms;
            if(shards[g].wall_s > wall_max) wall_max = shards[g].wall_s;
            launches += shards[g].stats.kernels_launched;
        }
        dev_ms_max_sum += ms_max; wall_sum += wall_max;
        for(int k = 0; k < n; k++)
        {
            if(readback && sink)
            {
                /* the frame is already upright RGB: checksum over its bytes, PPM = header + the buffer */
                const uint8_t* rgb = (const uint8_t*) pixels + frame_bytes * k;
                uint64_t h = 0xcbf29ce484222325ull; size_t nonzero = 0;
                for(size_t i = 0; i < frame_bytes; i++) h = (h ^ rgb[i]) * 0x100000001b3ull;
                for(size_t i = 0; i < frame; i++) nonzero += (rgb[3 * i] | rgb[3 * i + 1] | rgb[3 * i + 2]) != 0;
                printf("{\"frame\": %d, \"rgb_fnv\": \"%016llx\", \"nonzero\": %zu, \"checksum\": \"%016llx\"}\n", base + k,
                       (unsigned long long) h, nonzero, (unsigned long long) hashes[2 * (size_t) (base + k)]);
                if(dump) fwrite(rgb, 1, frame_bytes, dump);
                if(ppm_prefix)
                {
                    char path[1024];
                    snprintf(path, sizeof path, "%s%04d.ppm", ppm_prefix, base + k);
                    FILE* f = fopen(path, "wb");
                    if(f) { fprintf(f, "P6\n%d %d\n255\n", xres, yres); fwrite(rgb, 1, frame_bytes, f); fclose(f); }
                }
            }
            else if(readback)
            {
                const uint32_t* px = pixels + frame * k;
                size_t nonzero = 0;
                for(size_t i = 0; i < frame; i++) nonzero += px[i] != 0;
                printf("{\"frame\": %d, \"fnv\": \"%016llx\", \"nonzero\": %zu, \"checksum\": \"%016llx\"}\n", base + k,
                       (unsigned long long) gel_fnv1a64_words(px, frame), nonzero, (unsigned long long) hashes[2 * (size_t) (base + k)]);
                if(dump) fwrite(px, 4, frame, dump);
                if(ppm_prefix)
                {
                    char path[1024];
                    snprintf(path, sizeof path, "%s%04d.ppm", ppm_prefix, base + k);
                    gel_write_ppm(path, px, xres, yres);
                }
            }
            else
                printf("{\"frame\": %d, \"checksum\": \"%016llx\", \"zchecksum\": \"%016llx\"}\n", base + k,
                       (unsigned long long) hashes[2 * (size_t) (base + k)], (unsigned long long) hashes[2 * (size_t) (base + k) + 1]);
        }
    }
    gelcu_host_free(pixels);
    free(rects);
    if(dump) fclose(dump);
    printf("{\"summary\": true, \"views\": %d, \"triangles\": %d, \"res\": \"%dx%d\", \"gpus\": %d, \"device_ms\": %.4f, "
           "\"frames_per_s_device\": %.2f, \"mtri_per_s_device\": %.3f, \"wall_s\": %.4f, \"gpu_launches\": %llu}\n",
           nviews, obj.nfaces, xres, yres, gpus, dev_ms_max_sum,
           dev_ms_max_sum > 0 ? nviews / (dev_ms_max_sum * 1e-3) : 0.0,
           dev_ms_max_sum > 0 ? (double) obj.nfaces * nviews / (dev_ms_max_sum * 1e-3) / 1e6 : 0.0, wall_sum,
           (unsigned long long) launches);
    free(hashes); free(views);
    gel_mesh_free(&mesh); gel_obj_free(&obj); gel_texture_free(&tex);
    return status;
}
