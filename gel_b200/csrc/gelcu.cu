/* gelcu.cu -- host side of the C ABI in include/gelcu.h: contexts, device buffers, batching, stream/event
 * plumbing and the kernel launches.  The kernels themselves are in gel_kernels.cuh.
 *
 * Replaces /root/reference main.c:505-522 (reset, per-triangle transform, tdraw) for batches of views.
 * Per batch of B views, all on one stream with no host sync inside:
 *     batch_init_kernel -> K1 transform_kernel -> K2 bin_kernel -> K3 raster_kernel
 * (K3 also resets the tiles no triangle touched, as background stores between its work items).
 * Meshes of tiny triangles take the DIRECT pipeline instead (gel_direct.cuh):
 *     K1 -> D0 clear keys -> D1 near triangles -> D2 hi-Z -> D3 parked triangles -> D5 resolve/shade.
 * Frames are double-buffered in HBM so the device->host copy of batch b overlaps the kernels of batch b+1.
 */
#include "gel_mesh.cuh"
#include "gel_sink.cuh"
#include "gel_band.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>
#include <cfloat>
#include <cudaTypedefs.h>   /* PFN_cuTensorMapEncodeTiled */

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

struct Rect { int x0, y0, x1, y1; bool empty() const { return x1 < x0 || y1 < y0; } };   /* inclusive */

} /* namespace */

#ifndef GEL_RASTER_MODE
#define GEL_RASTER_MODE 1
#endif

struct gelcu_ctx
{
    int device = 0, xres = 0, yres = 0, tiles_x = 0, tiles_y = 0, ntiles = 0, num_sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr, side_stream = nullptr, hi_stream = nullptr, aux_stream = nullptr;   /* side / hi: HBM-bound fill beside the raster kernels (hi = higher priority); aux: small result copies */
    cudaEvent_t side_go = nullptr, side_done = nullptr, stats_go = nullptr, stats_ready[2] = { nullptr, nullptr };
    int fill_mode = 0, fill_ctas = 1, red_hint = 1, store_hint = 0, fill_sleep_ns = 0, fill_after = 0;   /* fill_after: trailing fill starts 0 = as the near pass drains, 1 = when it has ended, 2 = after the parked pass */   /* direct pipeline, background reset: 0 = plain grid after the near pass, 1 = persistent grid under it with evict-first stores, 2 = the same without the hint */
    /* mesh: distinct (position, normal) corners + per-triangle indices and texture coordinates */
    int ntri = 0, nuniq = 0; bool have_mesh = false, keys_dirty = true;
    float4 *d_vpos = nullptr, *d_vnrm = nullptr; uint32_t *d_i0 = nullptr, *d_i1 = nullptr, *d_i2 = nullptr; float2* d_uv = nullptr; uint4* d_trec = nullptr; bool trec_compact = false, allow_compact = true;
    /* texture */
    uint32_t* d_tex = nullptr; int tw = 0, th = 0;
    /* per-batch work buffers */
    int batch_opt = 0, batch = 0, cap_e = 0, cap_d = 0, ctas_per_sm = 1024 / RASTER_THREADS, stage_timing = 1;
    float4* d_xf = nullptr; uint32_t* d_entries = nullptr; uint4* d_descs = nullptr; int *d_heads = nullptr, *d_cursors = nullptr, *d_tile_lit = nullptr; uint32_t* d_lit_list = nullptr; uint32_t* d_vstat = nullptr; uint4* d_far = nullptr; float4* d_vrec = nullptr; float4* d_vconst = nullptr;   /* d_vconst: [view][4] per-view constants of K1 */   /* d_vrec: per-(view, triangle) records K2 leaves for K3 */
    /* direct pipeline */
    unsigned long long* d_keys = nullptr; uint32_t* d_hiz = nullptr; uint4* d_parked = nullptr; int *d_far_count = nullptr, *d_region = nullptr; int hbx = 0, hby = 0;
    int pipeline_opt = 0, pipeline_auto = 1, work_pipeline = 0;   /* 0 auto, 1 tile, 2 direct */
    double mean_tri_px = 0.0;
    uint32_t* d_flags = nullptr; unsigned long long* d_hash = nullptr; int* d_work = nullptr;
    uint32_t* d_pixel[2] = { nullptr, nullptr }; float* d_z[2] = { nullptr, nullptr };
    CUtensorMap tm_pixel[2] = {}, tm_z[2] = {}; bool tma_ok = false; int tma_reset = 1, raster_mode = GEL_RASTER_MODE, band_ctas_per_sm = 1024 / RASTER_THREADS;   /* raster_mode: 0 = a CTA per tile (raster_kernel), 1 = a warp per band (raster_band_kernel) */   /* tile pipeline: the frame buffers as TMA tensors (reset of untouched tiles) */
    uint8_t* d_rgb[2] = { nullptr, nullptr };   /* frame sink: upright 24-bit frames, allocated on first use */
    gelcu_view* d_views = nullptr; int views_cap = 0;
    int* h_cursors = nullptr; uint32_t* h_flags = nullptr; uint32_t* h_vstat = nullptr; int hcap = 0;
    std::vector<cudaEvent_t> ev;   /* EV_PER_BATCH per batch */
    cudaEvent_t render_done[2] = { nullptr, nullptr }, copy_done[2] = { nullptr, nullptr };
    int last_batch_views = 0, last_buf = 0;
    gelcu_stats stats = {};
    /* small calls (the interactive drop-in: one view per call) replay a captured CUDA graph instead of ~20 API calls */
    int use_graph = 1; unsigned state_gen = 0; bool keys_dirty_at_entry = false;                              /* state_gen: bumped whenever a buffer, the mesh, the texture or an option changes */
    cudaGraphExec_t graph_exec = nullptr; unsigned graph_gen = 0; int graph_n = 0, graph_flags = -1; uint64_t graph_kernels = 0, graph_d2h = 0;
    gelcu_view* h_views = nullptr; unsigned long long* h_hash = nullptr;   /* pinned staging for the graph's fixed-address copies */
    std::vector<std::pair<int, Rect> > pending_rects, done_rects;   /* region output: (view, rectangle) copied, resets pending / done */
};

namespace {

void free_bins(gelcu_ctx* c)
{
    dfree(c->d_entries); dfree(c->d_descs);
    c->cap_e = 0; c->cap_d = 0;
}

void free_work(gelcu_ctx* c)
{
    free_bins(c);
    dfree(c->d_xf); dfree(c->d_heads); dfree(c->d_cursors); dfree(c->d_tile_lit); dfree(c->d_lit_list); dfree(c->d_vstat); dfree(c->d_far); dfree(c->d_vrec); dfree(c->d_vconst); dfree(c->d_keys); dfree(c->d_hiz); dfree(c->d_parked); dfree(c->d_far_count); dfree(c->d_region);
    dfree(c->d_flags); dfree(c->d_hash); dfree(c->d_work);
    dfree(c->d_pixel[0]); dfree(c->d_pixel[1]); dfree(c->d_z[0]); dfree(c->d_z[1]); dfree(c->d_rgb[0]); dfree(c->d_rgb[1]);
    c->batch = 0;
}

/* Waits for everything THIS context has in flight.  Never cudaDeviceSynchronize(): a device-wide wait from one host thread is
 * an error -- and invalidates the capture -- while another context on the same device is capturing its small-call graph in
 * another thread ("operation not permitted when stream is capturing"). */
cudaError_t sync_ctx(gelcu_ctx* c)
{
    cudaStream_t all[5] = { c->stream, c->copy_stream, c->side_stream, c->hi_stream, c->aux_stream };
    cudaError_t first = cudaSuccess;
    for(cudaStream_t s : all)
        if(s) { const cudaError_t e = cudaStreamSynchronize(s); if(e != cudaSuccess && first == cudaSuccess) first = e; }
    return first;
}

/* the rasterisers are compiled for GEL_RASTER_MINB resident CTAs per SM: their shared memory has to allow as many (228 KB per SM, 1 KB reserved per CTA) */
static_assert((sizeof(BandSmem) + 1024) * GEL_RASTER_MINB <= 233472, "raster_band_kernel: shared memory leaves fewer resident CTAs than the launch bounds assume");
/* (raster_kernel, the CTA-per-tile form kept as raster_mode 0, runs 7 CTAs per SM since the reset patterns grew to 2 x 2 KB: gelcu_create asks the occupancy API) */

#ifndef GEL_RESOLVE_CTAS
#define GEL_RESOLVE_CTAS 1024
#endif
constexpr int RESOLVE_CTAS = GEL_RESOLVE_CTAS;  /* most CTAs per view in the direct pipeline's resolve pass (each walks strips of 8 columns) */
constexpr int GRAPH_MAX_VIEWS = 4;   /* calls of up to this many views (one batch) replay a captured graph */
constexpr int EV_PER_BATCH = 5;   /* start, after K1, after bin/clear, after the dominant raster kernel, end */

int active_pipeline(const gelcu_ctx* c) { return c->pipeline_opt ? c->pipeline_opt : c->pipeline_auto; }

size_t per_view_bytes(const gelcu_ctx* c, int cap_e, int cap_d)
{
    const size_t frame = (size_t) c->xres * c->yres;
    size_t b = 2 * frame * 8 + (size_t) c->nuniq * 16 + 64;
    if(active_pipeline(c) == 2) b += frame * 8 + (size_t) c->hbx * c->hby * 4 + (size_t) c->ntri * 16;
    else b += (size_t) cap_e * 4 + (size_t) cap_d * 16 + (size_t) c->ntiles * (NCHAIN + 2) * 4 + (size_t) c->ntri * VREC_QUADS * 16;
    return b;
}

/* A batch's frames (pixel or z: 32-bit words, index y + x*yres + view*xres*yres) as a 3-D TMA tensor (y, x, view) with boxes of
 * 32 rows x RESET_BOX_COLS columns: what the tile rasteriser's reset stores address.  The encoder is a driver entry point
 * fetched at run time (no link against libcuda).  Needs 16-byte row pitches: yres % 4 == 0. */
bool make_frame_map(CUtensorMap* tm, void* base, int xres, int yres, int B)
{
    static PFN_cuTensorMapEncodeTiled_v12000 encode = []() {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }();
    if(!encode || (yres & 3) != 0 || B < 1) return false;
    const cuuint64_t dims[3] = { (cuuint64_t) yres, (cuuint64_t) xres, (cuuint64_t) B };
    const cuuint64_t strides[2] = { (cuuint64_t) yres * 4, (cuuint64_t) xres * yres * 4 };
    const cuuint32_t box[3] = { (cuuint32_t) std::min(TH, yres), (cuuint32_t) std::min(RESET_BOX_COLS, xres), 1 }, estr[3] = { 1, 1, 1 };
    if(box[0] != (cuuint32_t) TH || box[1] != (cuuint32_t) RESET_BOX_COLS) return false;   /* frames smaller than a box keep the store loop */
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* Work buffers for batches of up to B views.  They are kept between calls: a steady stream of calls with the same (or a
 * smaller) number of views allocates nothing.  A larger batch or another pipeline rebuilds everything; a bin-pool overflow
 * (tile pipeline) only replaces the two pools. */
int ensure_work(gelcu_ctx* c, int B, int cap_e, int cap_d)
{
    const int pipe = active_pipeline(c);
    if(!(c->batch >= B && c->d_xf && c->work_pipeline == pipe))
    {
        free_work(c);
        const size_t frame = (size_t) c->xres * c->yres;
        CU(cudaMalloc(&c->d_xf, sizeof(float4) * std::max<size_t>(1, (size_t) B * c->nuniq)));
        CU(cudaMalloc(&c->d_vstat, sizeof(uint32_t) * VIEW_STAT_WORDS * B));
        CU(cudaMalloc(&c->d_vconst, sizeof(float4) * 4 * B));
        CU(cudaMalloc(&c->d_flags, sizeof(uint32_t) * B));
        CU(cudaMalloc(&c->d_hash, sizeof(unsigned long long) * 2 * B));
        CU(cudaMalloc(&c->d_cursors, sizeof(int) * 4 * B));
        if(pipe == 2)
        {
            CU(cudaMalloc(&c->d_keys, sizeof(unsigned long long) * B * frame));
            c->keys_dirty = true;                                   /* filled with "no winner" before the first batch */
            CU(cudaMalloc(&c->d_hiz, sizeof(uint32_t) * (size_t) B * c->hbx * c->hby));
            CU(cudaMalloc(&c->d_parked, sizeof(uint4) * std::max<size_t>(1, (size_t) B * c->ntri)));
            CU(cudaMalloc(&c->d_far_count, sizeof(int) * (size_t) B * (c->ntri / DIRECT_TRIS_PER_WARP_MIN + DIRECT_WARPS + 1)));
            CU(cudaMalloc(&c->d_region, sizeof(int) * REGION_WORDS * B));
        }
        else
        {
            CU(cudaMalloc(&c->d_heads, sizeof(int) * (size_t) B * c->ntiles * NCHAIN));
            CU(cudaMalloc(&c->d_tile_lit, sizeof(int) * (size_t) B * c->ntiles));
            CU(cudaMalloc(&c->d_lit_list, sizeof(uint32_t) * (size_t) B * c->ntiles));
            CU(cudaMalloc(&c->d_far, sizeof(uint4) * (size_t) FAR_CAP * c->num_sms * 16));   /* one scratch per resident rasteriser CTA */
            CU(cudaMalloc(&c->d_work, 4 * sizeof(int)));
            CU(cudaMalloc(&c->d_vrec, sizeof(float4) * VREC_QUADS * std::max<size_t>(1, (size_t) B * c->ntri)));
        }
        for(int k = 0; k < 2; k++)
        {
            CU(cudaMalloc(&c->d_pixel[k], sizeof(uint32_t) * B * frame));
            CU(cudaMalloc(&c->d_z[k], sizeof(float) * B * frame));
        }
        c->tma_ok = pipe != 2;
        for(int k = 0; k < 2 && c->tma_ok; k++)
            c->tma_ok = make_frame_map(&c->tm_pixel[k], c->d_pixel[k], c->xres, c->yres, B) && make_frame_map(&c->tm_z[k], c->d_z[k], c->xres, c->yres, B);
        c->batch = B; c->work_pipeline = pipe; c->state_gen++;
    }
    if(pipe != 2 && !(c->cap_e >= cap_e && c->cap_d >= cap_d && c->d_entries))
    {
        free_bins(c);
        CU(cudaMalloc(&c->d_entries, sizeof(uint32_t) * std::max<size_t>(1, (size_t) c->batch * cap_e)));
        CU(cudaMalloc(&c->d_descs, sizeof(uint4) * std::max<size_t>(1, (size_t) c->batch * cap_d)));
        c->cap_e = cap_e; c->cap_d = cap_d; c->state_gen++;
    }
    return GELCU_OK;
}

int default_batch(const gelcu_ctx* c, int cap_e, int cap_d)
{
    /* the shade passes address the batch's frames (and the direct pipeline its transformed vertices) with 32-bit element indices */
    size_t most = std::max<size_t>(1, std::min<size_t>(MAX_BATCH, 0xFFFFFFFFull / std::max<size_t>(1, (size_t) c->xres * c->yres)));
    if(active_pipeline(c) == 2) most = std::max<size_t>(1, std::min<size_t>(most, 0xFFFFFFFFull / std::max<size_t>(1, (size_t) c->nuniq)));
    if(c->batch_opt > 0) return (int) std::min<size_t>((size_t) c->batch_opt, most);
    const size_t budget = (size_t) 24 << 30;
    return (int) std::min<size_t>(most, std::max<size_t>(1, budget / per_view_bytes(c, cap_e, cap_d)));
}

/* Enqueues the kernels for `n` views starting at d_views + first into frame buffer `buf`. */
int enqueue_batch(gelcu_ctx* c, int first, int n, int buf, bool want_hash, bool want_rgb, bool want_stats, cudaEvent_t* ev /* null: no timing events (graph capture) */)
{
    cudaStream_t s = c->stream;
    const int pipe = c->work_pipeline;
    {
        const size_t cells = pipe != 2 ? (size_t) n * c->ntiles * NCHAIN : (size_t) n;
        const int grid = (int) std::min<size_t>((size_t) c->num_sms * 8, std::max<size_t>(1, (cells / 4 + 255) / 256));
        batch_init_kernel<<<grid, 256, 0, s>>>(c->d_vstat, c->d_cursors, c->d_flags, c->d_hash, pipe != 2 ? c->d_heads : nullptr, c->d_tile_lit, c->d_work,
                                               n, c->ntiles, want_hash ? 1 : 0, c->d_views + first, c->d_vconst, c->xres, c->yres);
        c->stats.kernels_launched++;
    }
    if(ev) CU(cudaEventRecord(ev[0], s));
    if(c->nuniq > 0)
    {
        transform_kernel<<<dim3((c->nuniq + 256 * XF_PER_THREAD - 1) / (256 * XF_PER_THREAD), n), 256, 0, s>>>(c->d_vconst, c->d_vpos, c->d_vnrm, c->d_xf, c->d_vstat, c->nuniq, c->xres, c->yres);
        c->stats.kernels_launched++;
    }
    if(ev) CU(cudaEventRecord(ev[1], s));
    if(want_stats)
    {
        /* region output: the per-view vertex statistics (complete after K1) go to the host on their own stream while
         * the raster kernels run; the host turns them into the rectangles it copies (region_from_stats) */
        CU(cudaEventRecord(c->stats_go, s));
        CU(cudaStreamWaitEvent(c->aux_stream, c->stats_go, 0));
        CU(cudaMemcpyAsync(c->h_vstat + (size_t) VIEW_STAT_WORDS * first, c->d_vstat, sizeof(uint32_t) * VIEW_STAT_WORDS * n, cudaMemcpyDeviceToHost, c->aux_stream));
        CU(cudaEventRecord(c->stats_ready[buf], c->aux_stream));
        if(!ev) CU(cudaStreamWaitEvent(s, c->stats_ready[buf], 0));       /* under capture every forked stream has to join again */
        c->stats.d2h_bytes += sizeof(uint32_t) * VIEW_STAT_WORDS * (size_t) n;
    }
    if(pipe == 2)
    {
        DirectParams dp = { c->d_xf, c->d_i0, c->d_i1, c->d_i2, c->d_trec, c->d_tex, c->tw, c->th, c->d_keys, c->d_hiz, c->d_parked, c->d_far_count, c->d_region, c->d_vstat,
                            c->d_pixel[buf], c->d_z[buf], c->d_hash, c->d_flags, c->ntri, c->nuniq, c->xres, c->yres, c->hbx, c->hby, n, DIRECT_TRIS_PER_WARP };
        /* a warp streams 1024 consecutive triangles; small batches (few views, small meshes) take shorter runs so that the
         * grid still has a few warps for every warp slot of the machine */
        while(dp.tpw > DIRECT_TRIS_PER_WARP_MIN && (long long) ((c->ntri + dp.tpw - 1) / dp.tpw) * n < 3LL * c->num_sms * 32) dp.tpw >>= 1;
        const int tris_per_cta = DIRECT_WARPS * dp.tpw;
        const dim3 rgrid((c->ntri + tris_per_cta - 1) / tris_per_cta, n);
        direct_clear_kernel<<<dim3(64, n), 256, 0, s>>>(dp);
        c->stats.kernels_launched++;
        if(ev) CU(cudaEventRecord(ev[2], s));
        CU(cudaEventRecord(c->side_go, s));                              /* the region is known from here on */
        /* Reset (main.c:413-417) of everything outside the view's region: pure stores, beside the raster kernels.
         *   fill_mode 0: a plain grid on the side stream, submitted AFTER the near pass -- it fills in as the near pass
         *                drains and runs beside the hi-Z / parked / resolve kernels;
         *   fill_mode 1: a few persistent CTAs per SM on a HIGHER-PRIORITY stream, submitted BEFORE the near pass, with
         *                evict-first stores: the 51 MB per cfg-3 frame stream out underneath the issue-bound near pass
         *                (which uses a fifth of the DRAM bandwidth) without evicting the L2-resident key buffer;
         *   fill_mode 2: mode 1 without the cache hint (the control of that experiment). */
        /*   fill_mode 3 / 4: TMA bulk stores (direct_fill_bulk_kernel) from a persistent grid, submitted before (3) or after (4)
         *                the near pass; 5 = 3 with the evict-first hint.  Frames whose height is not a multiple of 4, and calls
         *                that want checksums, use the store loops. */
        const bool bulk_ok = (c->yres & 3) == 0 && !want_hash;
        const int fmode = (c->fill_mode >= 3 && !bulk_ok) ? 0 : c->fill_mode;
        const bool early_fill = fmode == 1 || fmode == 2 || fmode == 3 || fmode == 5;
        cudaStream_t fs = early_fill ? c->hi_stream : c->side_stream;
        auto launch_fill = [&]() -> int {
            CU(cudaStreamWaitEvent(fs, c->side_go, 0));
            if(fmode >= 3)
            {
                direct_fill_bulk_kernel<<<c->num_sms * std::max(1, std::min(c->fill_ctas, 8)), 32, 0, fs>>>(dp, fmode == 5 ? 1 : 0);
            }
            else if(early_fill)
            {
                const int grid = c->num_sms * std::max(1, std::min(c->fill_ctas, 8));
                if(fmode == 1) { if(want_hash) direct_fill_persistent_kernel<true, true><<<grid, 256, 0, fs>>>(dp, c->fill_sleep_ns); else direct_fill_persistent_kernel<false, true><<<grid, 256, 0, fs>>>(dp, c->fill_sleep_ns); }
                else { if(want_hash) direct_fill_persistent_kernel<true, false><<<grid, 256, 0, fs>>>(dp, c->fill_sleep_ns); else direct_fill_persistent_kernel<false, false><<<grid, 256, 0, fs>>>(dp, c->fill_sleep_ns); }
            }
            else
            {
                const dim3 fgrid((c->yres + 1023) / 1024, c->xres, n);
                if(c->store_hint & 1) { if(want_hash) direct_fill_kernel<true, true><<<fgrid, 256, 0, fs>>>(dp); else direct_fill_kernel<false, true><<<fgrid, 256, 0, fs>>>(dp); }
                else { if(want_hash) direct_fill_kernel<true, false><<<fgrid, 256, 0, fs>>>(dp); else direct_fill_kernel<false, false><<<fgrid, 256, 0, fs>>>(dp); }
            }
            CU(cudaEventRecord(c->side_done, fs));
            c->stats.kernels_launched++;
            return GELCU_OK;
        };
        if(early_fill) { const int rc = launch_fill(); if(rc) return rc; }
        if(c->ntri > 0)
        {
            if(c->red_hint) direct_raster_kernel<0, true><<<rgrid, DIRECT_THREADS, 0, s>>>(dp);
            else direct_raster_kernel<0, false><<<rgrid, DIRECT_THREADS, 0, s>>>(dp);
            c->stats.kernels_launched++;
        }
        if(ev) CU(cudaEventRecord(ev[3], s));
        if(!early_fill && c->fill_after == 1) CU(cudaEventRecord(c->side_go, s));      /* the fill waits for the END of the near pass */
        if(!early_fill && c->fill_after != 2) { const int rc = launch_fill(); if(rc) return rc; }
        if(c->ntri > 0)
        {
            direct_hiz_kernel<<<dim3(32, n), 256, 0, s>>>(dp);
            if(c->red_hint) direct_raster_kernel<1, true><<<rgrid, DIRECT_THREADS, 0, s>>>(dp);
            else direct_raster_kernel<1, false><<<rgrid, DIRECT_THREADS, 0, s>>>(dp);
            c->stats.kernels_launched += 2;
        }
        if(!early_fill && c->fill_after == 2) { CU(cudaEventRecord(c->side_go, s)); const int rc = launch_fill(); if(rc) return rc; }
        const dim3 sgrid(std::min(RESOLVE_CTAS, (c->xres + 7) / 8), n);   /* one strip of 8 columns per CTA when the grid allows; CTAs past the region's last strip exit at once */
        {
            const int variant = (want_hash ? 4 : 0) | (c->trec_compact ? 2 : 0) | ((c->store_hint & 2) ? 1 : 0);
            switch(variant)
            {
                case 0: direct_resolve_kernel<false, false, false><<<sgrid, 256, 0, s>>>(dp); break;
                case 1: direct_resolve_kernel<false, false, true><<<sgrid, 256, 0, s>>>(dp); break;
                case 2: direct_resolve_kernel<false, true, false><<<sgrid, 256, 0, s>>>(dp); break;
                case 3: direct_resolve_kernel<false, true, true><<<sgrid, 256, 0, s>>>(dp); break;
                case 4: direct_resolve_kernel<true, false, false><<<sgrid, 256, 0, s>>>(dp); break;
                case 5: direct_resolve_kernel<true, false, true><<<sgrid, 256, 0, s>>>(dp); break;
                case 6: direct_resolve_kernel<true, true, false><<<sgrid, 256, 0, s>>>(dp); break;
                default: direct_resolve_kernel<true, true, true><<<sgrid, 256, 0, s>>>(dp); break;
            }
        }
        c->stats.kernels_launched++;
        CU(cudaStreamWaitEvent(s, c->side_done, 0));                    /* the batch ends when its frames are complete: resolve AND fill */
    }
    else
    {
        if(c->ntri > 0)
        {
            BinParams bp = { c->d_xf, c->d_i0, c->d_i1, c->d_i2, c->d_uv, c->d_vrec, c->d_entries, c->d_descs, c->d_heads, c->d_cursors, c->d_tile_lit, c->d_lit_list, c->d_work, c->d_flags,
                             c->ntri, c->nuniq, c->xres, c->yres, c->tiles_x, c->tiles_y, c->ntiles, c->cap_e, c->cap_d };
            bin_kernel<<<dim3((c->ntri + BIN_CHUNK - 1) / BIN_CHUNK, n), BIN_THREADS, 0, s>>>(bp);
            c->stats.kernels_launched++;
        }
        if(ev) CU(cudaEventRecord(ev[2], s));
        RasterParams rp = { c->d_vrec, c->d_entries, c->d_descs, c->d_heads, c->d_cursors, c->d_lit_list, c->d_tile_lit, c->d_vstat, c->d_far,
                            c->d_tex, c->tw, c->th, c->d_pixel[buf], c->d_z[buf], c->d_hash, c->d_flags, c->d_work,
                            c->ntri, c->nuniq, c->xres, c->yres, c->tiles_x, c->tiles_y, c->ntiles, c->cap_e, c->cap_d, n };
        rp.tma_reset = (c->tma_ok && c->tma_reset) ? 1 : 0;
        if(rp.tma_reset) { rp.tm_pixel = c->tm_pixel[buf]; rp.tm_z = c->tm_z[buf]; }
        if(c->raster_mode == 1)
        {
            const int grid = c->num_sms * std::min(c->band_ctas_per_sm, 16);
            if(want_hash) raster_band_kernel<true><<<grid, RASTER_THREADS, sizeof(BandSmem), s>>>(rp);
            else raster_band_kernel<false><<<grid, RASTER_THREADS, sizeof(BandSmem), s>>>(rp);
        }
        else
        {
            const int grid = c->num_sms * std::min(c->ctas_per_sm, 16);
            if(want_hash) raster_kernel<true><<<grid, RASTER_THREADS, sizeof(RasterSmem), s>>>(rp);
            else raster_kernel<false><<<grid, RASTER_THREADS, sizeof(RasterSmem), s>>>(rp);
        }
        c->stats.kernels_launched++;
        if(ev) CU(cudaEventRecord(ev[3], s));
    }
    if(ev) CU(cudaEventRecord(ev[4], s));
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
    if(c->h_vstat) cudaFreeHost(c->h_vstat);
    c->h_cursors = nullptr; c->h_flags = nullptr; c->h_vstat = nullptr; c->hcap = 0;
    CU(cudaMallocHost(&c->h_cursors, sizeof(int) * 4 * n));
    CU(cudaMallocHost(&c->h_flags, sizeof(uint32_t) * n));
    CU(cudaMallocHost(&c->h_vstat, sizeof(uint32_t) * VIEW_STAT_WORDS * n));
    c->hcap = n; c->state_gen++;
    return GELCU_OK;
}

/* Pipeline choice: mean projected triangle area (model units -> pixels at depth 0: yres/2 px per unit, main.c:290,
 * 302-314).  Tiny triangles -> direct pipeline; otherwise the tile pipeline. */
int choose_pipeline(int ntri, double mean_tri_px) { return (ntri >= 65536 && mean_tri_px < 32.0) ? 2 : 1; }

void drop_mesh(gelcu_ctx* c)
{
    c->have_mesh = false; c->ntri = 0; c->nuniq = 0; c->state_gen++;
    free_work(c);
    dfree(c->d_vpos); dfree(c->d_vnrm); dfree(c->d_i0); dfree(c->d_i1); dfree(c->d_i2); dfree(c->d_uv); dfree(c->d_trec);
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
    CU(cudaFuncSetAttribute(raster_band_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(BandSmem)));
    CU(cudaFuncSetAttribute(raster_band_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(BandSmem)));
    gelcu_ctx* c = new gelcu_ctx();
    c->device = device; c->xres = xres; c->yres = yres;
    c->tiles_x = (xres + TW - 1) / TW; c->tiles_y = (yres + TH - 1) / TH; c->ntiles = c->tiles_x * c->tiles_y;
    c->hbx = (xres + 7) / 8; c->hby = (yres + 7) / 8;
    cudaDeviceProp prop;
    if(cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    int resident = 0;   /* persistent rasteriser: exactly as many CTAs as fit */
    if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, raster_kernel<false>, RASTER_THREADS, sizeof(RasterSmem)) == cudaSuccess && resident > 0)
        c->ctas_per_sm = std::min(resident, 16);
    if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, raster_band_kernel<false>, RASTER_THREADS, sizeof(BandSmem)) == cudaSuccess && resident > 0)
        c->band_ctas_per_sm = std::min(resident, 16);
    cudaError_t s1 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    cudaError_t s2 = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if(s2 == cudaSuccess) s2 = cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking);
    if(s2 == cudaSuccess) s2 = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
    int prio_least = 0, prio_greatest = 0;
    if(s2 == cudaSuccess) s2 = cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if(s2 == cudaSuccess) s2 = cudaStreamCreateWithPriority(&c->hi_stream, cudaStreamNonBlocking, prio_greatest);
    if(s1 == cudaSuccess) s1 = cudaEventCreateWithFlags(&c->side_go, cudaEventDisableTiming);
    if(s2 == cudaSuccess) s2 = cudaEventCreateWithFlags(&c->side_done, cudaEventDisableTiming);
    if(s2 == cudaSuccess) s2 = cudaEventCreateWithFlags(&c->stats_go, cudaEventDisableTiming);
    for(int k = 0; k < 2 && s1 == cudaSuccess && s2 == cudaSuccess; k++)
    {
        s1 = cudaEventCreateWithFlags(&c->render_done[k], cudaEventDisableTiming);
        s2 = cudaEventCreateWithFlags(&c->copy_done[k], cudaEventDisableTiming);
        if(s2 == cudaSuccess) s2 = cudaEventCreateWithFlags(&c->stats_ready[k], cudaEventDisableTiming);
    }
    if(s1 != cudaSuccess || s2 != cudaSuccess) { delete c; return fail(GELCU_E_CUDA, "stream/event creation failed"); }
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
    CU(sync_ctx(c));
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
    double area = 0.0;
    for(int t = 0; t < ntri; t++)
    {
        const float* q = tv + 9 * (size_t) t;
        const double ux = q[3] - q[0], uy = q[4] - q[1], uz = q[5] - q[2], vx = q[6] - q[0], vy = q[7] - q[1], vz = q[8] - q[2];
        const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
        const double a2 = cx * cx + cy * cy + cz * cz;
        if(a2 == a2 && a2 < 1e30) area += 0.5 * sqrt(a2);
    }
    const double mean_px = ntri > 0 ? area / ntri * (0.5 * c->yres) * (0.5 * c->yres) : 0.0;
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

    /* the context has no mesh until the last upload has succeeded (a failed allocation must not leave half a mesh) */
    drop_mesh(c);
    const size_t nu = std::max<size_t>(1, vpos.size()), nt = std::max<size_t>(1, (size_t) ntri);
    CU(cudaMalloc(&c->d_vpos, sizeof(float4) * nu)); CU(cudaMalloc(&c->d_vnrm, sizeof(float4) * nu));
    CU(cudaMalloc(&c->d_i0, 4 * nt)); CU(cudaMalloc(&c->d_i1, 4 * nt)); CU(cudaMalloc(&c->d_i2, 4 * nt));
    CU(cudaMalloc(&c->d_uv, sizeof(float2) * 3 * nt));
    CU(cudaMalloc(&c->d_trec, sizeof(uint4) * std::max<size_t>(1, trec.size())));
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
    c->mean_tri_px = mean_px; c->pipeline_auto = choose_pipeline(ntri, mean_px);
    c->trec_compact = compact;
    c->ntri = ntri; c->nuniq = (int) vpos.size(); c->have_mesh = true;
    return GELCU_OK;
}

int gelcu_set_mesh_indexed(gelcu_ctx* c, const float* v, int nv, const float* vt, int nvt, const float* vn, int nvn, const int* faces, int nfaces)
{
    if(!c) return fail(GELCU_E_INVALID, "null context");
    if(nv < 0 || nvt < 0 || nvn < 0 || nfaces < 0 || (nfaces > 0 && (!v || !vt || !vn || !faces || nv == 0 || nvt == 0 || nvn == 0)))
        return fail(GELCU_E_INVALID, "bad indexed mesh arguments");
    CU(cudaSetDevice(c->device));
    CU(sync_ctx(c));
    drop_mesh(c);
    /* staging: the OBJ arrays as they are (36 bytes per face + the vertex lines, against 108 bytes per face of soups) */
    struct Staging
    {
        float *v = nullptr, *vt = nullptr, *vn = nullptr; int* faces = nullptr; unsigned long long* keys = nullptr; uint32_t *ranks = nullptr, *count = nullptr, *result = nullptr; double* area = nullptr;
        ~Staging() { cudaFree(v); cudaFree(vt); cudaFree(vn); cudaFree(faces); cudaFree(keys); cudaFree(ranks); cudaFree(count); cudaFree(result); cudaFree(area); }
    } st;
    const size_t ncorner = (size_t) nfaces * 3;
    size_t tsize = 1024; while(tsize < ncorner * 2) tsize <<= 1;
    CU(cudaMalloc(&st.v, sizeof(float) * 3 * std::max(1, nv))); CU(cudaMalloc(&st.vt, sizeof(float) * 3 * std::max(1, nvt))); CU(cudaMalloc(&st.vn, sizeof(float) * 3 * std::max(1, nvn)));
    CU(cudaMalloc(&st.faces, sizeof(int) * 9 * std::max<size_t>(1, (size_t) nfaces)));
    CU(cudaMalloc(&st.keys, sizeof(unsigned long long) * tsize)); CU(cudaMalloc(&st.ranks, sizeof(uint32_t) * tsize));
    CU(cudaMalloc(&st.count, sizeof(uint32_t) * std::max(1, nv))); CU(cudaMalloc(&st.result, sizeof(uint32_t) * 4)); CU(cudaMalloc(&st.area, sizeof(double)));
    cudaStream_t s = c->stream;
    if(nv) CU(cudaMemcpyAsync(st.v, v, sizeof(float) * 3 * (size_t) nv, cudaMemcpyHostToDevice, s));
    if(nvt) CU(cudaMemcpyAsync(st.vt, vt, sizeof(float) * 3 * (size_t) nvt, cudaMemcpyHostToDevice, s));
    if(nvn) CU(cudaMemcpyAsync(st.vn, vn, sizeof(float) * 3 * (size_t) nvn, cudaMemcpyHostToDevice, s));
    if(nfaces) CU(cudaMemcpyAsync(st.faces, faces, sizeof(int) * 9 * (size_t) nfaces, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(st.keys, 0xFF, sizeof(unsigned long long) * tsize, s));
    CU(cudaMemsetAsync(st.count, 0, sizeof(uint32_t) * std::max(1, nv), s));
    CU(cudaMemsetAsync(st.result, 0, sizeof(uint32_t) * 4, s));
    CU(cudaMemsetAsync(st.area, 0, sizeof(double), s));
    MeshBuild mb = { st.v, st.vt, st.vn, st.faces, nv, nvt, nvn, nfaces, st.keys, st.ranks, (uint32_t) (tsize - 1), st.count, st.result, st.area };
    const int grid = c->num_sms * 8;
    if(nv) mesh_maxlen_kernel<<<grid, 256, 0, s>>>(mb);
    if(nfaces) mesh_pairs_kernel<<<grid, 256, 0, s>>>(mb);
    mesh_scan_kernel<<<1, 1024, 0, s>>>(mb);
    uint32_t result[4] = { 0, 0, 0, 0 };
    CU(cudaMemcpyAsync(result, st.result, sizeof result, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    if(result[1] & MESH_BAD_INDEX) return fail(GELCU_E_INVALID, "a face index lies outside its v / vt / vn array (the reference reads out of bounds there, main.c:248-250)");
    float maxlen; memcpy(&maxlen, &result[0], 4);
    /* tvgen: `const int scale = vmaxlen(obj.vsv)` then `1.0f / scale` (main.c:244,253) */
    if(nfaces > 0 && !(maxlen >= 1.0f && maxlen < 2147483648.0f)) return fail(GELCU_E_INVALID, "max |v| = %g: the reference's scale 1.0f / (int) max|v| is undefined (main.c:244,253)", (double) maxlen);
    const int scale = nfaces > 0 ? (int) maxlen : 1;
    const float inv = 1.0f / (float) scale;
    const int nuniq = (int) result[2];
    const bool compact = c->allow_compact && (size_t) nuniq < ((size_t) 1 << TREC_COMPACT_BITS);
    const size_t nu = std::max<size_t>(1, (size_t) nuniq), nt = std::max<size_t>(1, (size_t) nfaces);
    CU(cudaMalloc(&c->d_vpos, sizeof(float4) * nu)); CU(cudaMalloc(&c->d_vnrm, sizeof(float4) * nu));
    CU(cudaMalloc(&c->d_i0, 4 * nt)); CU(cudaMalloc(&c->d_i1, 4 * nt)); CU(cudaMalloc(&c->d_i2, 4 * nt));
    CU(cudaMalloc(&c->d_uv, sizeof(float2) * 3 * nt));
    CU(cudaMalloc(&c->d_trec, sizeof(uint4) * (size_t) (compact ? 2 : TREC_QUADS) * nt));
    double area = 0.0;
    if(nfaces)
    {
        MeshOut mo = { c->d_vpos, c->d_vnrm, c->d_i0, c->d_i1, c->d_i2, c->d_uv, c->d_trec, compact ? 1 : 0, inv };
        mesh_emit_kernel<<<grid, 256, 0, s>>>(mb, mo);
        CU(cudaMemcpyAsync(&area, st.area, sizeof(double), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        CU(cudaGetLastError());
    }
    const double mean_px = nfaces > 0 ? area / nfaces * (0.5 * c->yres) * (0.5 * c->yres) : 0.0;
    c->mean_tri_px = mean_px; c->pipeline_auto = choose_pipeline(nfaces, mean_px);
    c->trec_compact = compact;
    c->ntri = nfaces; c->nuniq = nuniq; c->have_mesh = true;
    return GELCU_OK;
}

int gelcu_set_texture(gelcu_ctx* c, const uint32_t* xrgb, int w, int h)
{
    if(!c) return fail(GELCU_E_INVALID, "null context");
    if(!xrgb || w <= 0 || h <= 0) return fail(GELCU_E_INVALID, "bad texture arguments");
    CU(cudaSetDevice(c->device));
    CU(sync_ctx(c));
    dfree(c->d_tex);
    CU(cudaMalloc(&c->d_tex, sizeof(uint32_t) * (size_t) w * h));
    CU(cudaMemcpy(c->d_tex, xrgb, sizeof(uint32_t) * (size_t) w * h, cudaMemcpyHostToDevice));
    c->tw = w; c->th = h; c->state_gen++;
    return GELCU_OK;
}

int gelcu_set_option(gelcu_ctx* c, const char* name, int value)
{
    if(!c || !name) return fail(GELCU_E_INVALID, "null argument");
    c->state_gen++;
    if(!strcmp(name, "graph_small_calls")) c->use_graph = value != 0;
    else if(!strcmp(name, "batch_views")) { if(value < 0) return fail(GELCU_E_INVALID, "batch_views < 0"); c->batch_opt = value; cudaSetDevice(c->device); sync_ctx(c); free_work(c); }
    else if(!strcmp(name, "tma_reset")) c->tma_reset = value ? 1 : 0;
    else if(!strcmp(name, "raster_mode")) { if(value < 0 || value > 1) return fail(GELCU_E_INVALID, "raster_mode must be 0 (a CTA per tile) or 1 (a warp per band)"); c->raster_mode = value; }
    else if(!strcmp(name, "raster_ctas_per_sm")) { if(value < 1 || value > 16) return fail(GELCU_E_INVALID, "raster_ctas_per_sm out of [1,16]"); c->ctas_per_sm = value; c->band_ctas_per_sm = value; }
    else if(!strcmp(name, "band_carveout"))
    {
        /* shared-memory carve-out hint (percent of 228 KB, -1 = driver default) of the band rasteriser: 8 CTAs per SM need the 228 KB carve-out (28 KB of L1), 7 fit 196 KB (60 KB of L1) */
        if(value < -1 || value > 100) return fail(GELCU_E_INVALID, "band_carveout out of [-1,100]");
        cudaSetDevice(c->device);
        CU(cudaFuncSetAttribute(raster_band_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CU(cudaFuncSetAttribute(raster_band_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
    }
    else if(!strcmp(name, "stage_timing")) c->stage_timing = value != 0;
    else if(!strcmp(name, "compact_records")) c->allow_compact = value != 0;      /* takes effect at the next gelcu_set_mesh */
    else if(!strcmp(name, "fill_mode")) { if(value < 0 || value > 5) return fail(GELCU_E_INVALID, "fill_mode must be 0..5"); c->fill_mode = value; }
    else if(!strcmp(name, "fill_ctas_per_sm")) { if(value < 1 || value > 8) return fail(GELCU_E_INVALID, "fill_ctas_per_sm out of [1,8]"); c->fill_ctas = value; }
    else if(!strcmp(name, "red_hint")) c->red_hint = value != 0;
    else if(!strcmp(name, "near_carveout"))
    {
        /* shared-memory carve-out hint (percent of 228 KB, -1 = driver default) of the direct pipeline's near pass: the kernel needs 164 KB at 32 CTAs per SM; what the carve-out leaves is L1 */
        if(value < -1 || value > 100) return fail(GELCU_E_INVALID, "near_carveout out of [-1,100]");
        cudaSetDevice(c->device);
        CU(cudaFuncSetAttribute(direct_raster_kernel<0, true>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CU(cudaFuncSetAttribute(direct_raster_kernel<0, false>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
    }
    else if(!strcmp(name, "fill_after")) { if(value < 0 || value > 2) return fail(GELCU_E_INVALID, "fill_after must be 0, 1 or 2"); c->fill_after = value; }
    else if(!strcmp(name, "store_hint")) { if(value < 0 || value > 3) return fail(GELCU_E_INVALID, "store_hint must be 0..3 (bit 0: fill stores, bit 1: resolve stores evict-first)"); c->store_hint = value; }
    else if(!strcmp(name, "fill_sleep_ns")) { if(value < 0 || value > 1000000) return fail(GELCU_E_INVALID, "fill_sleep_ns out of [0, 1000000]"); c->fill_sleep_ns = value; }
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

/* ---- host side of the region output (gelcu_render_region) ---- */

Rect clip_rect(const gelcu_rect& r, int xres, int yres)
{
    Rect o = { std::max(r.x0, 0), std::max(r.y0, 0), std::min(r.x1, xres - 1), std::min(r.y1, yres - 1) };
    return o;
}

/* reset values (main.c:413-417) into rows [ya, yb] of columns [xa, xb] of a sideways frame */
void host_reset_sideways(uint32_t* pixel, float* z, int yres, int xa, int xb, int ya, int yb)
{
    if(xa > xb || ya > yb) return;
    for(int x = xa; x <= xb; x++)
    {
        const size_t base = (size_t) x * yres + ya;
        if(pixel) memset(pixel + base, 0, sizeof(uint32_t) * (size_t) (yb - ya + 1));
        if(z) std::fill(z + base, z + base + (yb - ya + 1), -FLT_MAX);
    }
}

/* zeros into columns [xa, xb] of the upright rows that hold screen rows [ya, yb] (row wy = yres - 1 - y) */
void host_reset_upright(uint8_t* rgb, int xres, int yres, int xa, int xb, int ya, int yb)
{
    if(xa > xb || ya > yb) return;
    for(int y = ya; y <= yb; y++)
        memset(rgb + ((size_t) (yres - 1 - y) * xres + xa) * 3, 0, 3 * (size_t) (xb - xa + 1));
}

/* the part of `old` that `now` does not cover goes back to reset values (at most four strips) */
void host_reset_stale(void* pixel, float* z, bool rgb8, int xres, int yres, const Rect& old, const Rect& now)
{
    if(old.empty()) return;
    auto strip = [&](int xa, int xb, int ya, int yb) {
        if(rgb8) host_reset_upright((uint8_t*) pixel, xres, yres, xa, xb, ya, yb);
        else host_reset_sideways((uint32_t*) pixel, z, yres, xa, xb, ya, yb);
    };
    if(now.empty()) { strip(old.x0, old.x1, old.y0, old.y1); return; }
    strip(old.x0, std::min(old.x1, now.x0 - 1), old.y0, old.y1);                         /* columns left of the new rectangle  */
    strip(std::max(old.x0, now.x1 + 1), old.x1, old.y0, old.y1);                         /* columns right of it                */
    const int xa = std::max(old.x0, now.x0), xb = std::min(old.x1, now.x1);              /* shared columns: rows below / above */
    strip(xa, xb, old.y0, std::min(old.y1, now.y0 - 1));
    strip(xa, xb, std::max(old.y0, now.y1 + 1), old.y1);
}

int render_impl(gelcu_ctx* c, const gelcu_view* views, int nviews,
                uint32_t* pixel_out, float* z_out, uint8_t* rgb_out, gelcu_rect* rect_io, uint64_t* hash_out, float* device_ms)
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
        c->views_cap = nviews; c->state_gen++;
    }
    rc = ensure_host(c, nviews); if(rc) return rc;
    const int nchunks = (c->ntri + BIN_CHUNK - 1) / BIN_CHUNK;
    int cap_e = c->cap_e > 0 ? c->cap_e : (int) std::min<size_t>((size_t) 1 << 30, (size_t) 4 * c->ntri + 4096);
    int cap_d = c->cap_d > 0 ? c->cap_d : (int) std::min<size_t>((size_t) 1 << 30, (size_t) 32 * nchunks + 4096);
    const bool want_frames = pixel_out || z_out || rgb_out;
    const bool want_region = rect_io != nullptr;

    for(int attempt = 0; attempt < 6; attempt++)
    {
        const int B = std::min(nviews, std::max(c->batch, default_batch(c, cap_e, cap_d)));
        rc = ensure_work(c, B, cap_e, cap_d); if(rc) return rc;
        if(rgb_out && !c->d_rgb[0])
        { for(int k = 0; k < 2; k++) CU(cudaMalloc(&c->d_rgb[k], 3 * frame * (size_t) c->batch)); c->state_gen++; }
        /* frames that go back to the host: a call is cut into at least four batches so that the copy of one batch
         * runs under the rendering of the next (the copy is the longer of the two by an order of magnitude) */
        int bsz = std::min(c->batch, nviews);
        if(want_frames && c->batch_opt == 0) bsz = std::min(c->batch, std::max(8, (nviews + 3) / 4));
        const int nb = (nviews + bsz - 1) / bsz;
        rc = ensure_events(c, nb); if(rc) return rc;
        c->stats.kernels_launched = 0; c->stats.h2d_bytes = 0; c->stats.d2h_bytes = 0; c->stats.batches = nb; c->stats.views = nviews;
        c->stats.h2d_bytes += sizeof(gelcu_view) * (size_t) nviews;
        c->keys_dirty_at_entry = c->d_keys && c->keys_dirty;
        if(c->d_keys && c->keys_dirty)
        {
            /* the resolve pass leaves the key buffer all "no winner"; only a fresh allocation or a call that failed
             * half way needs the whole buffer written */
            direct_keys_init_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(c->d_keys, (size_t) c->batch * frame);
            CU(cudaGetLastError());
            c->stats.kernels_launched++;
        }
        c->keys_dirty = true;

        bool graphed_call = false;
        auto issue_copies = [&](int b) -> int {
            const int buf = b & 1, first = b * bsz, n = std::min(bsz, nviews - first);
            CU(cudaStreamWaitEvent(c->copy_stream, c->render_done[buf], 0));
            if(want_region)
            {
                /* only each view's region crosses PCIe (a strided copy per frame); the strips of the caller's frame that
                 * the previous occupant lit and this view does not are reset here, on the host, while the copies run */
                /* (a graphed call has no usable event inside the graph: it waits for the whole small render) */
                CU(cudaEventSynchronize(graphed_call ? c->render_done[buf] : c->stats_ready[buf]));
                for(int v = 0; v < n; v++)
                {
                    Rect now;
                    if(!region_from_stats(c->h_vstat + (size_t) VIEW_STAT_WORDS * (first + v), c->xres, c->yres, now.x0, now.x1, now.y0, now.y1)) now = Rect{ 0, 0, -1, -1 };
                    if(!now.empty())
                    {
                        const size_t w = (size_t) (now.x1 - now.x0 + 1), h = (size_t) (now.y1 - now.y0 + 1);
                        if(rgb_out)
                        {
                            const size_t off = ((size_t) (c->yres - 1 - now.y1) * c->xres + now.x0) * 3, pitch = (size_t) c->xres * 3;
                            CU(cudaMemcpy2DAsync(rgb_out + 3 * frame * (first + v) + off, pitch, c->d_rgb[buf] + 3 * frame * v + off, pitch, 3 * w, h, cudaMemcpyDeviceToHost, c->copy_stream));
                            c->stats.d2h_bytes += 3 * w * h;
                        }
                        else
                        {
                            const size_t off = (size_t) now.x0 * c->yres + now.y0, pitch = (size_t) c->yres * 4;
                            if(pixel_out) { CU(cudaMemcpy2DAsync(pixel_out + frame * (first + v) + off, pitch, c->d_pixel[buf] + frame * v + off, pitch, 4 * h, w, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * w * h; }
                            if(z_out) { CU(cudaMemcpy2DAsync(z_out + frame * (first + v) + off, pitch, c->d_z[buf] + frame * v + off, pitch, 4 * h, w, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * w * h; }
                        }
                    }
                    c->pending_rects.push_back(std::make_pair(first + v, now));
                }
            }
            else
            {
                if(pixel_out) { CU(cudaMemcpyAsync(pixel_out + frame * first, c->d_pixel[buf], 4 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * frame * n; }
                if(z_out) { CU(cudaMemcpyAsync(z_out + frame * first, c->d_z[buf], 4 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * frame * n; }
                if(rgb_out) { CU(cudaMemcpyAsync(rgb_out + 3 * frame * first, c->d_rgb[buf], 3 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 3 * frame * n; }
            }
            CU(cudaEventRecord(c->copy_done[buf], c->copy_stream));
            return GELCU_OK;
        };
        /* host-side resets for the views whose copies have been issued (they touch only pixels outside the new rectangles) */
        auto reset_stale = [&]() {
            for(const auto& pr : c->pending_rects)
            {
                const int k = pr.first;
                const Rect old = clip_rect(rect_io[k], c->xres, c->yres);
                host_reset_stale(rgb_out ? (void*) (rgb_out + 3 * frame * k) : (void*) (pixel_out ? pixel_out + frame * k : nullptr),
                                 z_out ? z_out + frame * k : nullptr, rgb_out != nullptr, c->xres, c->yres, old, pr.second);
            }
            c->done_rects.insert(c->done_rects.end(), c->pending_rects.begin(), c->pending_rects.end());
            c->pending_rects.clear();
        };
        c->pending_rects.clear(); c->done_rects.clear();

        const bool graphed = c->use_graph && nb == 1 && nviews <= GRAPH_MAX_VIEWS && !want_region && !(c->d_keys && c->keys_dirty_at_entry);   /* region calls overlap their copies with the render instead */
        graphed_call = graphed;
        if(graphed)
        {
            /* Small call: everything up to the frames' own copies is ONE graph launch.  The graph is captured from the same
             * enqueue code (kernels, the cross-stream fill, the small result copies) the first time a call of this shape
             * arrives and replayed until a buffer, the mesh, the texture or an option changes (state_gen).  Its copies use
             * fixed pinned staging: the views go through h_views, the checksums through h_hash. */
            const int flags_key = (hash_out ? 1 : 0) | (rgb_out ? 2 : 0) | (want_region ? 4 : 0);
            if(!c->h_views) { CU(cudaMallocHost(&c->h_views, sizeof(gelcu_view) * GRAPH_MAX_VIEWS)); CU(cudaMallocHost(&c->h_hash, 16 * GRAPH_MAX_VIEWS)); }
            if(!c->graph_exec || c->graph_gen != c->state_gen || c->graph_n != nviews || c->graph_flags != flags_key)
            {
                if(c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
                const uint64_t k0 = c->stats.kernels_launched, b0 = c->stats.d2h_bytes;
                cudaGraph_t graph = nullptr;
                CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                int crc = GELCU_OK;
                cudaError_t ce = cudaMemcpyAsync(c->d_views, c->h_views, sizeof(gelcu_view) * nviews, cudaMemcpyHostToDevice, c->stream);
                if(ce == cudaSuccess) crc = enqueue_batch(c, 0, nviews, 0, hash_out != nullptr, rgb_out != nullptr, want_region, nullptr);
                if(ce == cudaSuccess && crc == GELCU_OK) ce = cudaMemcpyAsync(c->h_cursors, c->d_cursors, sizeof(int) * 4 * nviews, cudaMemcpyDeviceToHost, c->stream);
                if(ce == cudaSuccess && crc == GELCU_OK) ce = cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(uint32_t) * nviews, cudaMemcpyDeviceToHost, c->stream);
                if(ce == cudaSuccess && crc == GELCU_OK && hash_out) { ce = cudaMemcpyAsync(c->h_hash, c->d_hash, 16 * (size_t) nviews, cudaMemcpyDeviceToHost, c->stream); c->stats.d2h_bytes += 16 * (size_t) nviews; }
                const cudaError_t ee = cudaStreamEndCapture(c->stream, &graph);
                if(crc != GELCU_OK) { if(graph) cudaGraphDestroy(graph); return crc; }
                if(ce != cudaSuccess || ee != cudaSuccess) { if(graph) cudaGraphDestroy(graph); return fail(GELCU_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ee)); }
                const cudaError_t ie = cudaGraphInstantiate(&c->graph_exec, graph, 0);
                cudaGraphDestroy(graph);
                if(ie != cudaSuccess) { c->graph_exec = nullptr; return fail(GELCU_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
                c->graph_gen = c->state_gen; c->graph_n = nviews; c->graph_flags = flags_key;
                c->graph_kernels = c->stats.kernels_launched - k0; c->graph_d2h = c->stats.d2h_bytes - b0;
                c->stats.kernels_launched = k0; c->stats.d2h_bytes = b0;
            }
            memcpy(c->h_views, views, sizeof(gelcu_view) * nviews);
            CU(cudaEventRecord(c->ev[0], c->stream));
            CU(cudaGraphLaunch(c->graph_exec, c->stream));
            CU(cudaEventRecord(c->ev[4], c->stream));
            CU(cudaEventRecord(c->render_done[0], c->stream));
            c->stats.kernels_launched += c->graph_kernels; c->stats.d2h_bytes += c->graph_d2h;
            c->last_batch_views = nviews; c->last_buf = 0;
        }
        else
        for(int b = 0; b < nb; b++)
        {
            const int buf = b & 1, first = b * bsz, n = std::min(bsz, nviews - first);
            if(b == 0) CU(cudaMemcpyAsync(c->d_views, views, sizeof(gelcu_view) * nviews, cudaMemcpyHostToDevice, c->stream));
            if(b >= 2) CU(cudaStreamWaitEvent(c->stream, c->copy_done[buf], 0));
            if(want_region && b >= 1) CU(cudaStreamWaitEvent(c->stream, c->stats_ready[(b - 1) & 1], 0));   /* d_vstat is reused */
            rc = enqueue_batch(c, first, n, buf, hash_out != nullptr, rgb_out != nullptr, want_region, &c->ev[EV_PER_BATCH * b]); if(rc) return rc;
            /* small per-batch results ride the render stream (the next batch overwrites their device copies) */
            CU(cudaMemcpyAsync(c->h_cursors + 4 * first, c->d_cursors, sizeof(int) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaMemcpyAsync(c->h_flags + first, c->d_flags, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
            if(hash_out) { CU(cudaMemcpyAsync(hash_out + 2 * (size_t) first, c->d_hash, 16 * (size_t) n, cudaMemcpyDeviceToHost, c->stream)); c->stats.d2h_bytes += 16 * (size_t) n; }
            CU(cudaEventRecord(c->render_done[buf], c->stream));
            if(b >= 1 && want_frames) { rc = issue_copies(b - 1); if(rc) return rc; if(want_region) reset_stale(); }
            c->last_batch_views = n; c->last_buf = buf;
        }
        if(want_frames) { rc = issue_copies(nb - 1); if(rc) return rc; if(want_region) reset_stale(); }
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaStreamSynchronize(c->copy_stream));
        CU(cudaStreamSynchronize(c->side_stream));
        CU(cudaStreamSynchronize(c->hi_stream));
        CU(cudaStreamSynchronize(c->aux_stream));
        c->keys_dirty = false;
        if(graphed && hash_out) memcpy(hash_out, c->h_hash, 16 * (size_t) nviews);

        uint32_t flags = 0; int need_e = 0, need_d = 0; uint64_t entries = 0;
        for(int v = 0; v < nviews; v++)
        {
            flags |= c->h_flags[v];
            need_e = std::max(need_e, c->h_cursors[4 * v]); need_d = std::max(need_d, c->h_cursors[4 * v + 1]);
            entries += (uint64_t) c->h_cursors[4 * v];
        }
        if(flags & FLAG_OVERFLOW)
        {
            /* an entry or segment pool ran out: both grow geometrically past the measured need (so a slowly changing view
             * does not overflow again a few frames later), only the pools are reallocated, and the call is rendered again.
             * The frames already copied are overwritten by the second pass; their rectangles are the same. */
            cap_e = (int) std::min<size_t>((size_t) 1 << 30, std::max<size_t>((size_t) 2 * cap_e, (size_t) need_e + need_e / 2 + 1024));
            cap_d = (int) std::min<size_t>((size_t) 1 << 30, std::max<size_t>((size_t) 2 * cap_d, (size_t) need_d + need_d / 2 + 1024));
            if(per_view_bytes(c, cap_e, cap_d) * (size_t) c->batch > ((size_t) 150 << 30))
                return fail(GELCU_E_NOMEM, "bin lists need %d entries / %d segments per view, beyond device memory", need_e, need_d);
            continue;
        }
        if(want_region)
            for(const auto& pr : c->done_rects)
            {
                gelcu_rect& r = rect_io[pr.first];
                r.x0 = pr.second.x0; r.y0 = pr.second.y0; r.x1 = pr.second.x1; r.y1 = pr.second.y1;
            }
        /* device time = the batches' own spans (first kernel start to frames complete), summed: stalls of the render stream
         * between batches -- waiting for a frame buffer whose copy is still running -- are not kernel time */
        float ms = 0.0f, t = 0.0f;
        for(int b = 0; b < nb; b++)
        {
            cudaEvent_t* e = &c->ev[EV_PER_BATCH * b];
            CU(cudaEventElapsedTime(&t, e[0], e[4])); ms += t;
            if(c->stage_timing && !graphed)
            {
                CU(cudaEventElapsedTime(&t, e[0], e[1])); c->stats.ms_transform += t;
                CU(cudaEventElapsedTime(&t, e[1], e[2])); c->stats.ms_bin += t;
                CU(cudaEventElapsedTime(&t, e[2], e[4])); c->stats.ms_raster += t;
                CU(cudaEventElapsedTime(&t, e[2], e[3])); c->stats.ms_dominant += t;
            }
        }
        c->stats.ms_total = ms;
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
    return render_impl(c, views, nviews, pixel_out, z_out, nullptr, nullptr, hash_out, device_ms);
}

int gelcu_render_rgb8(gelcu_ctx* c, const gelcu_view* views, int nviews, uint8_t* rgb_out, uint64_t* hash_out, float* device_ms)
{
    if(!rgb_out && nviews > 0) return fail(GELCU_E_INVALID, "null rgb_out");
    return render_impl(c, views, nviews, nullptr, nullptr, rgb_out, nullptr, hash_out, device_ms);
}

int gelcu_render_region(gelcu_ctx* c, const gelcu_view* views, int nviews, void* pixel_io, float* z_io,
                        gelcu_rect* rect_io, int rgb8, uint64_t* hash_out, float* device_ms)
{
    if(nviews > 0 && (!pixel_io || !rect_io)) return fail(GELCU_E_INVALID, "gelcu_render_region needs pixel_io and rect_io");
    if(rgb8 && z_io) return fail(GELCU_E_INVALID, "z_io must be NULL with rgb8 output");
    if(rgb8) return render_impl(c, views, nviews, nullptr, nullptr, (uint8_t*) pixel_io, rect_io, hash_out, device_ms);
    return render_impl(c, views, nviews, (uint32_t*) pixel_io, z_io, nullptr, rect_io, hash_out, device_ms);
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
    sync_ctx(c);
    free_work(c);
    dfree(c->d_vpos); dfree(c->d_vnrm); dfree(c->d_i0); dfree(c->d_i1); dfree(c->d_i2); dfree(c->d_uv); dfree(c->d_trec);
    dfree(c->d_tex); dfree(c->d_views);
    if(c->h_cursors) cudaFreeHost(c->h_cursors);
    if(c->h_flags) cudaFreeHost(c->h_flags);
    if(c->h_vstat) cudaFreeHost(c->h_vstat);
    for(cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for(int k = 0; k < 2; k++) { if(c->render_done[k]) cudaEventDestroy(c->render_done[k]); if(c->copy_done[k]) cudaEventDestroy(c->copy_done[k]); if(c->stats_ready[k]) cudaEventDestroy(c->stats_ready[k]); }
    if(c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if(c->h_views) cudaFreeHost(c->h_views);
    if(c->h_hash) cudaFreeHost(c->h_hash);
    if(c->hi_stream) cudaStreamDestroy(c->hi_stream);
    if(c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if(c->stats_go) cudaEventDestroy(c->stats_go);
    if(c->stream) cudaStreamDestroy(c->stream);
    if(c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if(c->side_stream) cudaStreamDestroy(c->side_stream);
    if(c->side_go) cudaEventDestroy(c->side_go);
    if(c->side_done) cudaEventDestroy(c->side_done);
    delete c;
}

} /* extern "C" */
