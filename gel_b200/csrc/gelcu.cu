/* gelcu.cu -- host side of the C ABI in include/gelcu.h: contexts, device buffers, batching, stream/event
 * plumbing and the kernel launches.  The kernels themselves are in gel_kernels.cuh.
 *
 * Replaces /root/reference main.c:505-522 (reset, per-triangle transform, tdraw) for batches of views.
 * Per batch of B views, all on one stream with no host sync inside:
 *     memset heads/cursors/flags -> K1 transform_kernel -> K2 bin_kernel -> K3 raster_kernel
 * (K3 also resets the tiles no triangle touched, as background stores between its work items).
 * Meshes of tiny triangles take the DIRECT pipeline instead (gel_direct.cuh):
 *     K1 -> D0 clear keys -> D1 near triangles -> D2 hi-Z -> D3 parked triangles -> D5 resolve/shade.
 * Frames are double-buffered in HBM so the device->host copy of batch b overlaps the kernels of batch b+1.
 */
#include "gel_direct.cuh"
#include "gel_sink.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace gelk;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
    return fail(e_ == cudaErrorMemoryAllocation ? GELCU_E_NOMEM : GELCU_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while(0)

template<typename T> void dfree(T*& p) { if(p) cudaFree(p); p = nullptr; }

} /* namespace */

struct gelcu_ctx
{
    int device = 0, xres = 0, yres = 0, tiles_x = 0, tiles_y = 0, ntiles = 0, num_sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr, side_stream = nullptr;   /* side: HBM-bound fill overlapping the raster kernels */
    cudaEvent_t side_go = nullptr, side_done = nullptr;
    /* mesh: distinct (position, normal) corners + per-triangle indices and texture coordinates */
    int ntri = 0, nuniq = 0; bool have_mesh = false, keys_dirty = true;
    float4 *d_vpos = nullptr, *d_vnrm = nullptr; uint32_t *d_i0 = nullptr, *d_i1 = nullptr, *d_i2 = nullptr; float2* d_uv = nullptr; uint4* d_trec = nullptr; bool trec_compact = false, allow_compact = true;
    /* texture */
    uint32_t* d_tex = nullptr; int tw = 0, th = 0;
    /* per-batch work buffers */
    int batch_opt = 0, batch = 0, cap_e = 0, cap_d = 0, ctas_per_sm = 1024 / RASTER_THREADS, stage_timing = 1;
    float4* d_xf = nullptr; uint32_t* d_entries = nullptr; uint4* d_descs = nullptr; int *d_heads = nullptr, *d_cursors = nullptr, *d_tile_lit = nullptr, *d_lit_list = nullptr; uint32_t* d_vstat = nullptr; uint4* d_far = nullptr;
    /* direct pipeline */
    unsigned long long* d_keys = nullptr; uint32_t* d_hiz = nullptr; uint4* d_parked = nullptr; int *d_far_count = nullptr, *d_region = nullptr; int hbx = 0, hby = 0;
    int pipeline_opt = 0, pipeline_auto = 1, work_pipeline = 0;   /* 0 auto, 1 tile, 2 direct */
    double mean_tri_px = 0.0;
    uint32_t* d_flags = nullptr; unsigned long long* d_hash = nullptr; int* d_work = nullptr;
    uint32_t* d_pixel[2] = { nullptr, nullptr }; float* d_z[2] = { nullptr, nullptr };
    uint8_t* d_rgb[2] = { nullptr, nullptr };   /* frame sink: upright 24-bit frames, allocated on first use */
    gelcu_view* d_views = nullptr; int views_cap = 0;
    int* h_cursors = nullptr; uint32_t* h_flags = nullptr; int hcap = 0;
    uint32_t* h_vinit = nullptr;   /* pinned initial per-view statistics (VIEW_STAT_WORDS each) x MAX_BATCH */
    std::vector<cudaEvent_t> ev;   /* EV_PER_BATCH per batch */
    cudaEvent_t render_done[2] = { nullptr, nullptr }, copy_done[2] = { nullptr, nullptr };
    int last_batch_views = 0, last_buf = 0;
    gelcu_stats stats = {};
};

namespace {

void free_work(gelcu_ctx* c)
{
    dfree(c->d_xf); dfree(c->d_entries); dfree(c->d_descs); dfree(c->d_heads); dfree(c->d_cursors); dfree(c->d_tile_lit); dfree(c->d_lit_list); dfree(c->d_vstat); dfree(c->d_far); dfree(c->d_keys); dfree(c->d_hiz); dfree(c->d_parked); dfree(c->d_far_count); dfree(c->d_region);
    dfree(c->d_flags); dfree(c->d_hash); dfree(c->d_work);
    dfree(c->d_pixel[0]); dfree(c->d_pixel[1]); dfree(c->d_z[0]); dfree(c->d_z[1]); dfree(c->d_rgb[0]); dfree(c->d_rgb[1]);
    c->batch = 0; c->cap_e = 0; c->cap_d = 0;
}

#ifndef GEL_RESOLVE_CTAS
#define GEL_RESOLVE_CTAS 1024
#endif
constexpr int RESOLVE_CTAS = GEL_RESOLVE_CTAS;  /* most CTAs per view in the direct pipeline's resolve pass (each walks strips of 8 columns) */
constexpr int EV_PER_BATCH = 5;   /* start, after K1, after bin/clear, after the dominant raster kernel, end */

int active_pipeline(const gelcu_ctx* c) { return c->pipeline_opt ? c->pipeline_opt : c->pipeline_auto; }

size_t per_view_bytes(const gelcu_ctx* c, int cap_e, int cap_d)
{
    const size_t frame = (size_t) c->xres * c->yres;
    size_t b = 2 * frame * 8 + (size_t) c->nuniq * 16 + 64;
    if(active_pipeline(c) == 2) b += frame * 8 + (size_t) c->hbx * c->hby * 4 + (size_t) c->ntri * 16;
    else b += (size_t) cap_e * 4 + (size_t) cap_d * 16 + (size_t) c->ntiles * (NCHAIN + 2) * 4;
    return b;
}

int ensure_work(gelcu_ctx* c, int B, int cap_e, int cap_d)
{
    const int pipe = active_pipeline(c);
    if(c->batch >= B && c->cap_e >= cap_e && c->cap_d >= cap_d && c->d_xf && c->work_pipeline == pipe) return GELCU_OK;
    free_work(c);
    const size_t frame = (size_t) c->xres * c->yres;
    CU(cudaMalloc(&c->d_xf, sizeof(float4) * std::max<size_t>(1, (size_t) B * c->nuniq)));
    CU(cudaMalloc(&c->d_vstat, sizeof(uint32_t) * VIEW_STAT_WORDS * B));
    CU(cudaMalloc(&c->d_flags, sizeof(uint32_t) * B));
    CU(cudaMalloc(&c->d_hash, sizeof(unsigned long long) * 2 * B));
    CU(cudaMalloc(&c->d_cursors, sizeof(int) * 4 * B));
    if(pipe == 2)
    {
        CU(cudaMalloc(&c->d_keys, sizeof(unsigned long long) * B * frame));
        c->keys_dirty = true;                                   /* filled with "no winner" before the first batch */
        CU(cudaMalloc(&c->d_hiz, sizeof(uint32_t) * (size_t) B * c->hbx * c->hby));
        CU(cudaMalloc(&c->d_parked, sizeof(uint4) * std::max<size_t>(1, (size_t) B * c->ntri)));
        CU(cudaMalloc(&c->d_far_count, sizeof(int) * (size_t) B * (c->ntri / DIRECT_TRIS_PER_WARP + DIRECT_WARPS + 1)));
        CU(cudaMalloc(&c->d_region, sizeof(int) * REGION_WORDS * B));
    }
    else
    {
        CU(cudaMalloc(&c->d_entries, sizeof(uint32_t) * std::max<size_t>(1, (size_t) B * cap_e)));
        CU(cudaMalloc(&c->d_descs, sizeof(uint4) * std::max<size_t>(1, (size_t) B * cap_d)));
        CU(cudaMalloc(&c->d_heads, sizeof(int) * (size_t) B * c->ntiles * NCHAIN));
        CU(cudaMalloc(&c->d_tile_lit, sizeof(int) * (size_t) B * c->ntiles));
        CU(cudaMalloc(&c->d_lit_list, sizeof(int) * (size_t) B * c->ntiles));
        CU(cudaMalloc(&c->d_far, sizeof(uint4) * (size_t) FAR_CAP * c->num_sms * 16));   /* one scratch per resident rasteriser CTA */
        CU(cudaMalloc(&c->d_work, 2 * sizeof(int)));
    }
    for(int k = 0; k < 2; k++)
    {
        CU(cudaMalloc(&c->d_pixel[k], sizeof(uint32_t) * B * frame));
        CU(cudaMalloc(&c->d_z[k], sizeof(float) * B * frame));
    }
    c->batch = B; c->cap_e = cap_e; c->cap_d = cap_d; c->work_pipeline = pipe;
    return GELCU_OK;
}

int default_batch(const gelcu_ctx* c, int cap_e, int cap_d)
{
    if(c->batch_opt > 0) return std::min(c->batch_opt, MAX_BATCH);
    const size_t budget = (size_t) 24 << 30;
    return (int) std::min<size_t>(MAX_BATCH, std::max<size_t>(1, budget / per_view_bytes(c, cap_e, cap_d)));
}

/* Enqueues the kernels for `n` views starting at d_views + first into frame buffer `buf`. */
int enqueue_batch(gelcu_ctx* c, int first, int n, int buf, bool want_hash, bool want_rgb, cudaEvent_t* ev)
{
    cudaStream_t s = c->stream;
    const int pipe = c->work_pipeline;
    CU(cudaMemcpyAsync(c->d_vstat, c->h_vinit, sizeof(uint32_t) * VIEW_STAT_WORDS * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(c->d_cursors, 0, sizeof(int) * 4 * n, s));
    CU(cudaMemsetAsync(c->d_flags, 0, sizeof(uint32_t) * n, s));
    if(want_hash) CU(cudaMemsetAsync(c->d_hash, 0, sizeof(unsigned long long) * 2 * n, s));
    if(pipe != 2)
    {
        CU(cudaMemsetAsync(c->d_heads, 0xFF, sizeof(int) * (size_t) n * c->ntiles * NCHAIN, s));
        CU(cudaMemsetAsync(c->d_tile_lit, 0, sizeof(int) * (size_t) n * c->ntiles, s));
        CU(cudaMemsetAsync(c->d_work, 0, 2 * sizeof(int), s));
    }
    CU(cudaEventRecord(ev[0], s));
    if(c->nuniq > 0)
    {
        transform_kernel<<<dim3((c->nuniq + 256 * XF_PER_THREAD - 1) / (256 * XF_PER_THREAD), n), 256, 0, s>>>(c->d_views + first, c->d_vpos, c->d_vnrm, c->d_xf, c->d_vstat, c->nuniq, c->xres, c->yres);
        c->stats.kernels_launched++;
    }
    CU(cudaEventRecord(ev[1], s));
    if(pipe == 2)
    {
        DirectParams dp = { c->d_xf, c->d_i0, c->d_i1, c->d_i2, c->d_trec, c->d_tex, c->tw, c->th, c->d_keys, c->d_hiz, c->d_parked, c->d_far_count, c->d_region, c->d_vstat,
                            c->d_pixel[buf], c->d_z[buf], c->d_hash, c->d_flags, c->ntri, c->nuniq, c->xres, c->yres, c->hbx, c->hby, n };
        const int tris_per_cta = DIRECT_WARPS * DIRECT_TRIS_PER_WARP;
        const dim3 rgrid((c->ntri + tris_per_cta - 1) / tris_per_cta, n);
        direct_clear_kernel<<<dim3(64, n), 256, 0, s>>>(dp);
        c->stats.kernels_launched++;
        CU(cudaEventRecord(ev[2], s));
        CU(cudaEventRecord(c->side_go, s));                              /* the region is known from here on */
        if(c->ntri > 0)
        {
            direct_raster_kernel<0><<<rgrid, DIRECT_THREADS, 0, s>>>(dp);
            c->stats.kernels_launched++;
        }
        CU(cudaEventRecord(ev[3], s));
        /* everything outside the view's region is reset by pure stores on the side stream, submitted AFTER the near
         * pass: D1 (one warp per CTA, 56 registers) leaves room for one fill CTA per SM, so this HBM traffic runs
         * underneath the instruction-bound raster kernels */
        CU(cudaStreamWaitEvent(c->side_stream, c->side_go, 0));
        const dim3 fgrid((c->yres + 1023) / 1024, c->xres, n);
        if(want_hash) direct_fill_kernel<true><<<fgrid, 256, 0, c->side_stream>>>(dp);
        else direct_fill_kernel<false><<<fgrid, 256, 0, c->side_stream>>>(dp);
        CU(cudaEventRecord(c->side_done, c->side_stream));
        c->stats.kernels_launched++;
        if(c->ntri > 0)
        {
            direct_hiz_kernel<<<dim3(32, n), 256, 0, s>>>(dp);
            direct_raster_kernel<1><<<rgrid, DIRECT_THREADS, 0, s>>>(dp);
            c->stats.kernels_launched += 2;
        }
        const dim3 sgrid(std::min(RESOLVE_CTAS, (c->xres + 7) / 8), n);   /* one strip of 8 columns per CTA when the grid allows; CTAs past the region's last strip exit at once */
        if(c->trec_compact) { if(want_hash) direct_resolve_kernel<true, true><<<sgrid, 256, 0, s>>>(dp); else direct_resolve_kernel<false, true><<<sgrid, 256, 0, s>>>(dp); }
        else { if(want_hash) direct_resolve_kernel<true, false><<<sgrid, 256, 0, s>>>(dp); else direct_resolve_kernel<false, false><<<sgrid, 256, 0, s>>>(dp); }
        c->stats.kernels_launched++;
        CU(cudaStreamWaitEvent(s, c->side_done, 0));
    }
    else
    {
        if(c->ntri > 0)
        {
            BinParams bp = { c->d_xf, c->d_i0, c->d_i1, c->d_i2, c->d_entries, c->d_descs, c->d_heads, c->d_cursors, c->d_tile_lit, c->d_lit_list, c->d_flags,
                             c->ntri, c->nuniq, c->xres, c->yres, c->tiles_x, c->tiles_y, c->ntiles, c->cap_e, c->cap_d };
            bin_kernel<<<dim3((c->ntri + BIN_CHUNK - 1) / BIN_CHUNK, n), BIN_THREADS, 0, s>>>(bp);
            c->stats.kernels_launched++;
        }
        CU(cudaEventRecord(ev[2], s));
        RasterParams rp = { c->d_xf, c->d_i0, c->d_i1, c->d_i2, c->d_uv, c->d_entries, c->d_descs, c->d_heads, c->d_cursors, c->d_lit_list, c->d_tile_lit, c->d_vstat, c->d_far,
                            c->d_tex, c->tw, c->th, c->d_pixel[buf], c->d_z[buf], c->d_hash, c->d_flags, c->d_work,
                            c->ntri, c->nuniq, c->xres, c->yres, c->tiles_x, c->tiles_y, c->ntiles, c->cap_e, c->cap_d, n };
        const int grid = c->num_sms * std::min(c->ctas_per_sm, 16);
        if(want_hash) raster_kernel<true><<<grid, RASTER_THREADS, sizeof(RasterSmem), s>>>(rp);
        else raster_kernel<false><<<grid, RASTER_THREADS, sizeof(RasterSmem), s>>>(rp);
        c->stats.kernels_launched++;
        CU(cudaEventRecord(ev[3], s));
    }
    CU(cudaEventRecord(ev[4], s));
    if(want_rgb)
    {
        /* frame sink: un-rotated 24-bit copy for the device -> host transfer; after the path's last
         * event, so the render figures (ms_total, device_ms) mean the same with and without it */
        sink_rgb8_kernel<<<dim3((c->xres + SINK_TX - 1) / SINK_TX, (c->yres + SINK_TY - 1) / SINK_TY, n), SINK_THREADS, 0, s>>>(c->d_pixel[buf], c->d_rgb[buf], c->xres, c->yres);
        c->stats.kernels_launched++;
    }
    CU(cudaGetLastError());
    return GELCU_OK;
}

int ensure_events(gelcu_ctx* c, int nbatches)
{
    while((int) c->ev.size() < EV_PER_BATCH * nbatches)
    {
        cudaEvent_t e; CU(cudaEventCreate(&e)); c->ev.push_back(e);
    }
    return GELCU_OK;
}

int ensure_host(gelcu_ctx* c, int n)
{
    if(c->hcap >= n) return GELCU_OK;
    if(c->h_cursors) cudaFreeHost(c->h_cursors);
    if(c->h_flags) cudaFreeHost(c->h_flags);
    c->h_cursors = nullptr; c->h_flags = nullptr; c->hcap = 0;
    CU(cudaMallocHost(&c->h_cursors, sizeof(int) * 4 * n));
    CU(cudaMallocHost(&c->h_flags, sizeof(uint32_t) * n));
    c->hcap = n;
    return GELCU_OK;
}

int check_ready(gelcu_ctx* c)
{
    if(!c) return fail(GELCU_E_INVALID, "null context");
    if(!c->have_mesh) return fail(GELCU_E_INVALID, "gelcu_set_mesh has not been called");
    if(!c->d_tex) return fail(GELCU_E_INVALID, "gelcu_set_texture has not been called");
    CU(cudaSetDevice(c->device));
    return GELCU_OK;
}

} /* namespace */

extern "C" {

const char* gelcu_last_error(void) { return g_err.c_str(); }

int gelcu_device_count(void)
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess) return fail(GELCU_E_NOGPU, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int gelcu_create(gelcu_ctx** out, int device, int xres, int yres)
{
    if(!out) return fail(GELCU_E_INVALID, "null out pointer");
    *out = nullptr;
    /* tile coordinates travel as bytes inside K2: at most 256 tiles of 32 pixels per axis */
    if(xres <= 0 || yres <= 0 || xres > 256 * TW || yres > 256 * TH) return fail(GELCU_E_INVALID, "resolution %dx%d out of range (max %dx%d)", xres, yres, 256 * TW, 256 * TH);
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess || n <= 0)
        return fail(GELCU_E_NOGPU, "no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count is 0");
    if(device < 0 || device >= n) return fail(GELCU_E_INVALID, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    CU(cudaFuncSetAttribute(raster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(RasterSmem)));
    CU(cudaFuncSetAttribute(raster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(RasterSmem)));
    gelcu_ctx* c = new gelcu_ctx();
    c->device = device; c->xres = xres; c->yres = yres;
    c->tiles_x = (xres + TW - 1) / TW; c->tiles_y = (yres + TH - 1) / TH; c->ntiles = c->tiles_x * c->tiles_y;
    c->hbx = (xres + 7) / 8; c->hby = (yres + 7) / 8;
    cudaDeviceProp prop;
    if(cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    int resident = 0;   /* persistent rasteriser: exactly as many CTAs as fit */
    if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, raster_kernel<false>, RASTER_THREADS, sizeof(RasterSmem)) == cudaSuccess && resident > 0)
        c->ctas_per_sm = std::min(resident, 16);
    cudaError_t s1 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    cudaError_t s2 = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if(s2 == cudaSuccess) s2 = cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking);
    if(s1 == cudaSuccess) s1 = cudaEventCreateWithFlags(&c->side_go, cudaEventDisableTiming);
    if(s2 == cudaSuccess) s2 = cudaEventCreateWithFlags(&c->side_done, cudaEventDisableTiming);
    for(int k = 0; k < 2 && s1 == cudaSuccess && s2 == cudaSuccess; k++)
    {
        s1 = cudaEventCreateWithFlags(&c->render_done[k], cudaEventDisableTiming);
        s2 = cudaEventCreateWithFlags(&c->copy_done[k], cudaEventDisableTiming);
    }
    if(s1 != cudaSuccess || s2 != cudaSuccess) { delete c; return fail(GELCU_E_CUDA, "stream/event creation failed"); }
    if(cudaMallocHost(&c->h_vinit, sizeof(uint32_t) * VIEW_STAT_WORDS * MAX_BATCH) != cudaSuccess) { delete c; return fail(GELCU_E_NOMEM, "pinned allocation failed"); }
    for(int v = 0; v < MAX_BATCH; v++)
    {
        uint32_t* w = c->h_vinit + VIEW_STAT_WORDS * v;
        w[0] = 0xFFFFFFFFu; w[1] = 0u;                                     /* depth range        */
        w[2] = 0x7FFFFFFFu; w[3] = 0x80000000u; w[4] = 0x7FFFFFFFu; w[5] = 0x80000000u;   /* screen bbox (int min/max) */
        w[6] = 0u; w[7] = 0u;                                              /* parked triangles   */
    }
    *out = c;
    return GELCU_OK;
}

int gelcu_tile_grid(gelcu_ctx* c, int* tile_w, int* tile_h, int* tiles_x, int* tiles_y)
{
    if(!c) return fail(GELCU_E_INVALID, "null context");
    if(tile_w) *tile_w = TW;
    if(tile_h) *tile_h = TH;
    if(tiles_x) *tiles_x = c->tiles_x;
    if(tiles_y) *tiles_y = c->tiles_y;
    return GELCU_OK;
}

int gelcu_set_mesh(gelcu_ctx* c, const float* tv, const float* tn, const float* tt, int ntri)
{
    if(!c) return fail(GELCU_E_INVALID, "null context");
    if(ntri < 0 || (ntri > 0 && (!tv || !tn || !tt))) return fail(GELCU_E_INVALID, "bad mesh arguments");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    /* Load-time layout: the reference's soup repeats every shared corner (main.c:242-286); the transform is
     * a pure function of (position, normal), so identical corners are merged here (bitwise equality) and
     * the per-frame kernels run once per distinct corner and index it per triangle. */
    const size_t ncorner = (size_t) ntri * 3;
    std::vector<float4> vpos, vnrm;
    std::vector<uint32_t> idx[3];
    for(int k = 0; k < 3; k++) idx[k].resize(ntri);
    size_t tsize = 16; while(tsize < ncorner * 2) tsize <<= 1;
    std::vector<uint32_t> table(tsize, 0xFFFFFFFFu);
    for(size_t cidx = 0; cidx < ncorner; cidx++)
    {
        uint32_t key[6];
        memcpy(key, tv + 3 * cidx, 12); memcpy(key + 3, tn + 3 * cidx, 12);
        uint64_t h = 0xcbf29ce484222325ull;
        for(int k = 0; k < 6; k++) h = (h ^ key[k]) * 0x100000001b3ull;
        size_t slot = (size_t) (h ^ (h >> 29)) & (tsize - 1);
        uint32_t found = 0xFFFFFFFFu;
        for(;; slot = (slot + 1) & (tsize - 1))
        {
            const uint32_t v = table[slot];
            if(v == 0xFFFFFFFFu) break;
            uint32_t other[6];
            memcpy(other, &vpos[v], 12); memcpy(other + 3, &vnrm[v], 12);
            if(memcmp(other, key, 24) == 0) { found = v; break; }
        }
        if(found == 0xFFFFFFFFu)
        {
            found = (uint32_t) vpos.size();
            table[slot] = found;
            vpos.push_back(make_float4(tv[3 * cidx], tv[3 * cidx + 1], tv[3 * cidx + 2], 0.0f));
            vnrm.push_back(make_float4(tn[3 * cidx], tn[3 * cidx + 1], tn[3 * cidx + 2], 0.0f));
        }
        idx[cidx % 3][cidx / 3] = found;
    }
    /* Pipeline choice: mean projected triangle area (model units -> pixels at depth 0: yres/2 px per unit, main.c:290,
     * 302-314).  Tiny triangles -> direct pipeline; otherwise the tile pipeline. */
    double area = 0.0;
    for(int t = 0; t < ntri; t++)
    {
        const float* q = tv + 9 * (size_t) t;
        const double ux = q[3] - q[0], uy = q[4] - q[1], uz = q[5] - q[2], vx = q[6] - q[0], vy = q[7] - q[1], vz = q[8] - q[2];
        const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
        const double a2 = cx * cx + cy * cy + cz * cz;
        if(a2 == a2 && a2 < 1e30) area += 0.5 * sqrt(a2);
    }
    c->mean_tri_px = ntri > 0 ? area / ntri * (0.5 * c->yres) * (0.5 * c->yres) : 0.0;
    c->pipeline_auto = (ntri >= 65536 && c->mean_tri_px < 32.0) ? 2 : 1;
    std::vector<float2> uv(ncorner);
    for(size_t cidx = 0; cidx < ncorner; cidx++) uv[cidx] = make_float2(tt[3 * cidx], tt[3 * cidx + 1]);   /* tt.z is never read, main.c:360-361 */

    /* the direct pipeline's resolve pass gathers per winning pixel: one 64-byte record per triangle (vertex indices +
     * texture coordinates) costs three 16-byte loads in one cache line instead of six loads in four arrays */
    const bool compact = c->allow_compact && vpos.size() < ((size_t) 1 << TREC_COMPACT_BITS);   /* three 21-bit indices fit one 64-bit word: 32-byte record */
    const int quads = compact ? 2 : TREC_QUADS;
    std::vector<uint4> trec((size_t) quads * ntri, make_uint4(0u, 0u, 0u, 0u));
    for(int t = 0; t < ntri; t++)
    {
        uint32_t w[8];
        memcpy(w, &uv[3 * (size_t) t], 24);
        if(compact)
        {
            const unsigned long long packed = (unsigned long long) idx[0][t] | (unsigned long long) idx[1][t] << 21 | (unsigned long long) idx[2][t] << 42;
            trec[2 * (size_t) t] = make_uint4((uint32_t) packed, (uint32_t) (packed >> 32), w[0], w[1]);
            trec[2 * (size_t) t + 1] = make_uint4(w[2], w[3], w[4], w[5]);
        }
        else
        {
            trec[(size_t) TREC_QUADS * t] = make_uint4(idx[0][t], idx[1][t], idx[2][t], 0u);
            trec[(size_t) TREC_QUADS * t + 1] = make_uint4(w[0], w[1], w[2], w[3]);
            trec[(size_t) TREC_QUADS * t + 2] = make_uint4(w[4], w[5], 0u, 0u);
        }
    }

    free_work(c);
    dfree(c->d_vpos); dfree(c->d_vnrm); dfree(c->d_i0); dfree(c->d_i1); dfree(c->d_i2); dfree(c->d_uv); dfree(c->d_trec);
    c->ntri = ntri; c->nuniq = (int) vpos.size(); c->have_mesh = true;
    const size_t nu = std::max<size_t>(1, vpos.size()), nt = std::max<size_t>(1, (size_t) ntri);
    CU(cudaMalloc(&c->d_vpos, sizeof(float4) * nu)); CU(cudaMalloc(&c->d_vnrm, sizeof(float4) * nu));
    CU(cudaMalloc(&c->d_i0, 4 * nt)); CU(cudaMalloc(&c->d_i1, 4 * nt)); CU(cudaMalloc(&c->d_i2, 4 * nt));
    CU(cudaMalloc(&c->d_uv, sizeof(float2) * 3 * nt));
    CU(cudaMalloc(&c->d_trec, sizeof(uint4) * std::max<size_t>(1, trec.size())));
    c->trec_compact = compact;
    if(ntri > 0)
    {
        CU(cudaMemcpy(c->d_vpos, vpos.data(), sizeof(float4) * vpos.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_vnrm, vnrm.data(), sizeof(float4) * vnrm.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_i0, idx[0].data(), 4 * (size_t) ntri, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_i1, idx[1].data(), 4 * (size_t) ntri, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_i2, idx[2].data(), 4 * (size_t) ntri, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_uv, uv.data(), sizeof(float2) * uv.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_trec, trec.data(), sizeof(uint4) * trec.size(), cudaMemcpyHostToDevice));
    }
    return GELCU_OK;
}

int gelcu_set_texture(gelcu_ctx* c, const uint32_t* xrgb, int w, int h)
{
    if(!c) return fail(GELCU_E_INVALID, "null context");
    if(!xrgb || w <= 0 || h <= 0) return fail(GELCU_E_INVALID, "bad texture arguments");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    dfree(c->d_tex);
    CU(cudaMalloc(&c->d_tex, sizeof(uint32_t) * (size_t) w * h));
    CU(cudaMemcpy(c->d_tex, xrgb, sizeof(uint32_t) * (size_t) w * h, cudaMemcpyHostToDevice));
    c->tw = w; c->th = h;
    return GELCU_OK;
}

int gelcu_set_option(gelcu_ctx* c, const char* name, int value)
{
    if(!c || !name) return fail(GELCU_E_INVALID, "null argument");
    if(!strcmp(name, "batch_views")) { if(value < 0) return fail(GELCU_E_INVALID, "batch_views < 0"); c->batch_opt = value; cudaSetDevice(c->device); cudaDeviceSynchronize(); free_work(c); }
    else if(!strcmp(name, "raster_ctas_per_sm")) { if(value < 1 || value > 16) return fail(GELCU_E_INVALID, "raster_ctas_per_sm out of [1,16]"); c->ctas_per_sm = value; }
    else if(!strcmp(name, "stage_timing")) c->stage_timing = value != 0;
    else if(!strcmp(name, "compact_records")) c->allow_compact = value != 0;      /* takes effect at the next gelcu_set_mesh */
    else if(!strcmp(name, "pipeline")) { if(value < 0 || value > 2) return fail(GELCU_E_INVALID, "pipeline must be 0 (auto), 1 (tile) or 2 (direct)"); c->pipeline_opt = value; }
    else return fail(GELCU_E_INVALID, "unknown option '%s'", name);
    return GELCU_OK;
}

int gelcu_get_stats(gelcu_ctx* c, gelcu_stats* out)
{
    if(!c || !out) return fail(GELCU_E_INVALID, "null argument");
    *out = c->stats;
    return GELCU_OK;
}

namespace {

int render_impl(gelcu_ctx* c, const gelcu_view* views, int nviews,
                uint32_t* pixel_out, float* z_out, uint8_t* rgb_out, uint64_t* hash_out, float* device_ms)
{
    int rc = check_ready(c);
    if(rc) return rc;
    if(nviews < 0 || (nviews > 0 && !views)) return fail(GELCU_E_INVALID, "bad views argument");
    if(device_ms) *device_ms = 0.0f;
    c->stats = gelcu_stats();
    c->stats.unique_vertices = c->nuniq; c->stats.triangles = c->ntri;
    if(nviews == 0) return GELCU_OK;
    const size_t frame = (size_t) c->xres * c->yres;
    if(c->views_cap < nviews)
    {
        dfree(c->d_views);
        CU(cudaMalloc(&c->d_views, sizeof(gelcu_view) * nviews));
        c->views_cap = nviews;
    }
    rc = ensure_host(c, nviews); if(rc) return rc;
    const int nchunks = (c->ntri + BIN_CHUNK - 1) / BIN_CHUNK;
    int cap_e = c->cap_e > 0 ? c->cap_e : (int) std::min<size_t>((size_t) 1 << 30, (size_t) 2 * c->ntri + 4096);
    int cap_d = c->cap_d > 0 ? c->cap_d : (int) std::min<size_t>((size_t) 1 << 30, (size_t) 32 * nchunks + 4096);

    for(int attempt = 0; attempt < 4; attempt++)
    {
        const int B = std::min(nviews, std::max(c->batch, default_batch(c, cap_e, cap_d)));
        rc = ensure_work(c, B, cap_e, cap_d); if(rc) return rc;
        if(rgb_out && !c->d_rgb[0])
            for(int k = 0; k < 2; k++) CU(cudaMalloc(&c->d_rgb[k], 3 * frame * (size_t) c->batch));
        /* frames that go back to the host: a call is cut into at least four batches so that the copy of one batch
         * runs under the rendering of the next (the copy is the longer of the two by an order of magnitude) */
        int bsz = c->batch;
        if((pixel_out || z_out || rgb_out) && c->batch_opt == 0) bsz = std::min(c->batch, std::max(8, (nviews + 3) / 4));
        const int nb = (nviews + bsz - 1) / bsz;
        rc = ensure_events(c, nb); if(rc) return rc;
        c->stats.kernels_launched = 0; c->stats.h2d_bytes = 0; c->stats.d2h_bytes = 0; c->stats.batches = nb; c->stats.views = nviews;
        CU(cudaMemcpyAsync(c->d_views, views, sizeof(gelcu_view) * nviews, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += sizeof(gelcu_view) * (size_t) nviews;
        if(c->d_keys && c->keys_dirty)
        {
            /* the resolve pass leaves the key buffer all "no winner"; only a fresh allocation or a call that failed
             * half way needs the whole buffer written */
            direct_keys_init_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(c->d_keys, (size_t) c->batch * frame);
            CU(cudaGetLastError());
            c->stats.kernels_launched++;
        }
        c->keys_dirty = true;

        auto issue_copies = [&](int b) -> int {
            const int buf = b & 1, first = b * bsz, n = std::min(bsz, nviews - first);
            CU(cudaStreamWaitEvent(c->copy_stream, c->render_done[buf], 0));
            if(pixel_out) { CU(cudaMemcpyAsync(pixel_out + frame * first, c->d_pixel[buf], 4 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * frame * n; }
            if(z_out) { CU(cudaMemcpyAsync(z_out + frame * first, c->d_z[buf], 4 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * frame * n; }
            if(rgb_out) { CU(cudaMemcpyAsync(rgb_out + 3 * frame * first, c->d_rgb[buf], 3 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 3 * frame * n; }
            CU(cudaEventRecord(c->copy_done[buf], c->copy_stream));
            return GELCU_OK;
        };

        for(int b = 0; b < nb; b++)
        {
            const int buf = b & 1, first = b * bsz, n = std::min(bsz, nviews - first);
            if(b >= 2) CU(cudaStreamWaitEvent(c->stream, c->copy_done[buf], 0));
            rc = enqueue_batch(c, first, n, buf, hash_out != nullptr, rgb_out != nullptr, &c->ev[EV_PER_BATCH * b]); if(rc) return rc;
            /* small per-batch results ride the render stream (the next batch overwrites their device copies) */
            CU(cudaMemcpyAsync(c->h_cursors + 4 * first, c->d_cursors, sizeof(int) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaMemcpyAsync(c->h_flags + first, c->d_flags, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
            if(hash_out) { CU(cudaMemcpyAsync(hash_out + 2 * (size_t) first, c->d_hash, 16 * (size_t) n, cudaMemcpyDeviceToHost, c->stream)); c->stats.d2h_bytes += 16 * (size_t) n; }
            CU(cudaEventRecord(c->render_done[buf], c->stream));
            if(b >= 1 && (pixel_out || z_out || rgb_out)) { rc = issue_copies(b - 1); if(rc) return rc; }
            c->last_batch_views = n; c->last_buf = buf;
        }
        if(pixel_out || z_out || rgb_out) { rc = issue_copies(nb - 1); if(rc) return rc; }
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaStreamSynchronize(c->copy_stream));
        CU(cudaStreamSynchronize(c->side_stream));
        c->keys_dirty = false;

        uint32_t flags = 0; int need_e = 0, need_d = 0; uint64_t entries = 0;
        for(int v = 0; v < nviews; v++)
        {
            flags |= c->h_flags[v];
            need_e = std::max(need_e, c->h_cursors[4 * v]); need_d = std::max(need_d, c->h_cursors[4 * v + 1]);
            entries += (uint64_t) c->h_cursors[4 * v];
        }
        if(flags & FLAG_OVERFLOW)
        {
            /* an entry or segment pool ran out: grow both to the measured need and render the call again */
            cap_e = (int) std::min<size_t>((size_t) 1 << 30, std::max<size_t>(cap_e, (size_t) need_e + need_e / 8 + 1024));
            cap_d = (int) std::min<size_t>((size_t) 1 << 30, std::max<size_t>(cap_d, (size_t) need_d + need_d / 8 + 1024));
            if(per_view_bytes(c, cap_e, cap_d) > ((size_t) 150 << 30))
                return fail(GELCU_E_NOMEM, "bin lists need %d entries / %d segments per view, beyond device memory", need_e, need_d);
            const int keep = c->batch_opt;
            free_work(c);
            c->batch_opt = keep;
            continue;
        }
        float ms = 0.0f, t = 0.0f;
        CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[EV_PER_BATCH * (nb - 1) + 4]));
        c->stats.ms_total = ms;
        if(c->stage_timing)
            for(int b = 0; b < nb; b++)
            {
                cudaEvent_t* e = &c->ev[EV_PER_BATCH * b];
                CU(cudaEventElapsedTime(&t, e[0], e[1])); c->stats.ms_transform += t;
                CU(cudaEventElapsedTime(&t, e[1], e[2])); c->stats.ms_bin += t;
                CU(cudaEventElapsedTime(&t, e[2], e[4])); c->stats.ms_raster += t;
                CU(cudaEventElapsedTime(&t, e[2], e[3])); c->stats.ms_dominant += t;
            }
        if(device_ms) *device_ms = ms;
        c->stats.bin_entries = entries;
        c->stats.pipeline = (uint32_t) c->work_pipeline;
        c->stats.flags = flags & ~FLAG_OVERFLOW;
        if(c->stats.flags) return fail(GELCU_W_CLIPPED, "input left the reference's defined domain (flags 0x%x: 1 = bbox off screen, 2 = texel out of range); clipped", c->stats.flags);
        return GELCU_OK;
    }
    return fail(GELCU_E_NOMEM, "bin list capacity did not converge");
}

} /* namespace */

int gelcu_render(gelcu_ctx* c, const gelcu_view* views, int nviews,
                 uint32_t* pixel_out, float* z_out, uint64_t* hash_out, float* device_ms)
{
    return render_impl(c, views, nviews, pixel_out, z_out, nullptr, hash_out, device_ms);
}

int gelcu_render_rgb8(gelcu_ctx* c, const gelcu_view* views, int nviews, uint8_t* rgb_out, uint64_t* hash_out, float* device_ms)
{
    if(!rgb_out && nviews > 0) return fail(GELCU_E_INVALID, "null rgb_out");
    return render_impl(c, views, nviews, nullptr, nullptr, rgb_out, hash_out, device_ms);
}

int gelcu_read_frame(gelcu_ctx* c, int slot, uint32_t* pixel_out, float* z_out)
{
    int rc = check_ready(c);
    if(rc) return rc;
    if(slot < 0 || slot >= c->last_batch_views) return fail(GELCU_E_INVALID, "slot %d outside the last batch (%d views)", slot, c->last_batch_views);
    const size_t frame = (size_t) c->xres * c->yres;
    if(pixel_out) CU(cudaMemcpy(pixel_out, c->d_pixel[c->last_buf] + frame * slot, 4 * frame, cudaMemcpyDeviceToHost));
    if(z_out) CU(cudaMemcpy(z_out, c->d_z[c->last_buf] + frame * slot, 4 * frame, cudaMemcpyDeviceToHost));
    return GELCU_OK;
}

int gelcu_debug_transform(gelcu_ctx* c, const gelcu_view* view, float* vew, float* shade)
{
    int rc = check_ready(c);
    if(rc) return rc;
    if(!view) return fail(GELCU_E_INVALID, "null view");
    rc = gelcu_render(c, view, 1, nullptr, nullptr, nullptr, nullptr);
    if(rc < 0) return rc;
    std::vector<float4> xf(std::max(1, c->nuniq));
    std::vector<uint32_t> idx[3];
    CU(cudaMemcpy(xf.data(), c->d_xf, sizeof(float4) * c->nuniq, cudaMemcpyDeviceToHost));
    for(int k = 0; k < 3; k++)
    {
        idx[k].resize(std::max(1, c->ntri));
        CU(cudaMemcpy(idx[k].data(), k == 0 ? c->d_i0 : k == 1 ? c->d_i1 : c->d_i2, 4 * (size_t) c->ntri, cudaMemcpyDeviceToHost));
    }
    for(int t = 0; t < c->ntri; t++)
        for(int k = 0; k < 3; k++)
        {
            const float4 o = xf[idx[k][t]];
            if(vew) { vew[9 * t + 3 * k] = o.x; vew[9 * t + 3 * k + 1] = o.y; vew[9 * t + 3 * k + 2] = o.z; }
            if(shade) shade[3 * t + k] = o.w;
        }
    return GELCU_OK;
}

int gelcu_debug_bins(gelcu_ctx* c, const gelcu_view* view, int* counts, int* entries, int cap, int* total)
{
    int rc = check_ready(c);
    if(rc) return rc;
    if(!view) return fail(GELCU_E_INVALID, "null view");
    const int keep = c->pipeline_opt;
    c->pipeline_opt = 1;                                   /* lists only exist in the tile pipeline */
    rc = gelcu_render(c, view, 1, nullptr, nullptr, nullptr, nullptr);
    c->pipeline_opt = keep;
    if(rc < 0) return rc;
    const int ne = c->h_cursors[0], nd = c->h_cursors[1];
    std::vector<int> heads((size_t) c->ntiles * NCHAIN);
    std::vector<uint4> desc(std::max(1, nd)); std::vector<uint32_t> ent(std::max(1, ne));
    CU(cudaMemcpy(heads.data(), c->d_heads, sizeof(int) * heads.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(desc.data(), c->d_descs, sizeof(uint4) * nd, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ent.data(), c->d_entries, sizeof(uint32_t) * ne, cudaMemcpyDeviceToHost));
    if(total) *total = ne;
    int w = 0;
    for(int t = 0; t < c->ntiles; t++)
    {
        std::vector<int> tris;
        for(int ch = 0; ch < NCHAIN; ch++)
            for(int cur = heads[(size_t) t * NCHAIN + ch]; cur >= 0; cur = (int) desc[cur].x)
                for(uint32_t k = 0; k < desc[cur].z; k++) tris.push_back((int) ent[desc[cur].y + k]);
        std::sort(tris.begin(), tris.end());
        if(counts) counts[t] = (int) tris.size();
        for(size_t k = 0; k < tris.size() && entries && w < cap; k++) entries[w++] = tris[k];
    }
    return GELCU_OK;
}

int gelcu_host_alloc(void** p, size_t bytes)
{
    if(!p) return fail(GELCU_E_INVALID, "null pointer");
    CU(cudaMallocHost(p, bytes ? bytes : 1));
    return GELCU_OK;
}

void gelcu_host_free(void* p) { if(p) cudaFreeHost(p); }

void gelcu_destroy(gelcu_ctx* c)
{
    if(!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_work(c);
    dfree(c->d_vpos); dfree(c->d_vnrm); dfree(c->d_i0); dfree(c->d_i1); dfree(c->d_i2); dfree(c->d_uv); dfree(c->d_trec);
    dfree(c->d_tex); dfree(c->d_views);
    if(c->h_cursors) cudaFreeHost(c->h_cursors);
    if(c->h_flags) cudaFreeHost(c->h_flags);
    if(c->h_vinit) cudaFreeHost(c->h_vinit);
    for(cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for(int k = 0; k < 2; k++) { if(c->render_done[k]) cudaEventDestroy(c->render_done[k]); if(c->copy_done[k]) cudaEventDestroy(c->copy_done[k]); }
    if(c->stream) cudaStreamDestroy(c->stream);
    if(c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if(c->side_stream) cudaStreamDestroy(c->side_stream);
    if(c->side_go) cudaEventDestroy(c->side_go);
    if(c->side_done) cudaEventDestroy(c->side_done);
    delete c;
}

} /* extern "C" */
