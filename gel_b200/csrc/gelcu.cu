/* gelcu.cu -- sm_100a kernels + C ABI (include/gelcu.h) for gel's per-frame render path.
 *
 * Replaces /root/reference main.c:505-522 (reset, per-triangle transform, tdraw) for batches of views.
 * Pipeline per batch of B views (all launches on one stream, no host sync inside a batch):
 *
 *   K1 transform_kernel   one thread per (view, unique corner): tviewnrm/tviewtri/tperspective/tviewport
 *                         (main.c:372-390, 302-314, 288-300) -> float4 (screen x, y, z, shade)
 *   K2 bin_count_kernel   one thread per (view, triangle): bbox (main.c:344-347) -> screen-tile rect,
 *                         per-tile counts (warp-aggregated atomics)
 *      bin_scan_kernel    per view: exclusive scan of the tile counts -> list offsets
 *      bin_fill_kernel    writes (i0,i1,i2,tri) entries into each tile's list
 *   K3 raster_kernel      persistent CTAs pull (view, tile) items; a tile's depth+winner lives in shared
 *                         memory as one 64-bit key per pixel; the barycentric loop (main.c:348-356) runs
 *                         one triangle per lane (small) or one triangle per warp (large); the winning
 *                         fragment is shaded once (main.c:358-366) and the tile is written back in one
 *                         coalesced pass.
 *
 * Draw-order semantics (main.c:356, strict `z > zbuff`, first-drawn wins ties) are kept EXACTLY by the key
 *   key = zkey(z) << 32 | (0xFFFFFFFF - triangle_index),   resolved with a 64-bit max:
 * "first triangle in submission order to reach a strictly greater z" == "greatest z, ties to the lowest
 * index", so the result is independent of the order in which a tile's triangles are processed.
 */
#include "../../include/gelcu.h"
#include "gel_math.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int TW = 32;            /* tile width  (screen x, the framebuffer's SLOW axis)              */
constexpr int TH = 32;            /* tile height (screen y, contiguous in memory: index y + x*yres)   */
constexpr int RASTER_THREADS = 256;
constexpr int SMALL_MAX = 48;     /* bbox-in-tile pixels up to which a triangle is walked by one lane */
constexpr unsigned long long CLEAR_KEY = (0x00800000ull << 32) | 0xFFFFFFFFull;   /* zkey(-FLT_MAX), no winner */
constexpr uint32_t FLAG_CLIPPED = 1u, FLAG_TEXCLAMP = 2u, FLAG_OVERFLOW = 0x80000000u;

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

/* ------------------------------------------------------------------------------------------------ */
/* K1: vertex transform                                                                             */
/* ------------------------------------------------------------------------------------------------ */

__global__ void __launch_bounds__(256)
transform_kernel(const gelcu_view* __restrict__ views, const float4* __restrict__ vpos,
                 const float4* __restrict__ vnrm, float4* __restrict__ xf, int nuniq, int xres, int yres)
{
    const int view = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= nuniq) return;
    const gel::ViewConst c = gel::view_const(reinterpret_cast<const float*>(views + view), xres, yres);
    const float4 p = __ldg(vpos + i);
    const float4 n = __ldg(vnrm + i);
    float4 o;
    gel::transform_corner(c, p.x, p.y, p.z, n.x, n.y, n.z, o.x, o.y, o.z, o.w);
    xf[(size_t) view * nuniq + i] = o;
}

/* ------------------------------------------------------------------------------------------------ */
/* K2: triangle bbox -> tile rect, per-tile lists                                                   */
/* ------------------------------------------------------------------------------------------------ */

struct BinParams
{
    const float4* xf; const uint32_t *i0, *i1, *i2;
    ushort4* rect; int* tile_count; int* tile_off; int* tile_cursor; uint4* entries;
    uint32_t* flags; int* totals;
    int ntri, nuniq, xres, yres, tiles_x, tiles_y, ntiles, cap;
};

/* Warp-aggregated increment: lanes that target the same counter are grouped with match.any; the group
 * leader adds the group size once and every lane gets base + its rank (lane order = submission order
 * within the warp). */
__device__ __forceinline__ int warp_agg_add(int* counter_base, int tile, bool active)
{
    const unsigned act = __ballot_sync(0xFFFFFFFFu, active);
    if(!active) return 0;
    const unsigned peers = __match_any_sync(act, tile);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int base = 0;
    if(lane == leader) base = atomicAdd(counter_base + tile, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + rank;
}

__global__ void __launch_bounds__(256)
bin_count_kernel(BinParams p)
{
    const int view = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < p.ntri;
    int tx0 = 1, tx1 = 0, ty0 = 1, ty1 = 0;
    if(live)
    {
        const float4* xf = p.xf + (size_t) view * p.nuniq;
        const float4 a = __ldg(xf + __ldg(p.i0 + t));
        const float4 b = __ldg(xf + __ldg(p.i1 + t));
        const float4 c = __ldg(xf + __ldg(p.i2 + t));
        int x0 = gel::trunc_i(fminf(a.x, fminf(b.x, c.x)));      /* main.c:344-347 */
        int y0 = gel::trunc_i(fminf(a.y, fminf(b.y, c.y)));
        int x1 = gel::trunc_i(fmaxf(a.x, fmaxf(b.x, c.x)));
        int y1 = gel::trunc_i(fmaxf(a.y, fmaxf(b.y, c.y)));
        if(x0 < 0 || y0 < 0 || x1 > p.xres - 1 || y1 > p.yres - 1)
        {
            atomicOr(p.flags + view, FLAG_CLIPPED);              /* the reference has UB here (Q3) */
            x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, p.xres - 1); y1 = min(y1, p.yres - 1);
        }
        if(x0 <= x1 && y0 <= y1) { tx0 = x0 / TW; tx1 = x1 / TW; ty0 = y0 / TH; ty1 = y1 / TH; }
        p.rect[(size_t) view * p.ntri + t] = make_ushort4((unsigned short) tx0, (unsigned short) ty0,
                                                          (unsigned short) tx1, (unsigned short) ty1);
    }
    int* count = p.tile_count + (size_t) view * p.ntiles;
    const bool any = live && tx0 <= tx1 && ty0 <= ty1;
    /* first tile through the aggregated path (neighbouring triangles mostly share it) ... */
    warp_agg_add(count, any ? tx0 * p.tiles_y + ty0 : 0, any);
    /* ... the rest of the rect, if any, one atomic per tile */
    if(any)
        for(int tx = tx0; tx <= tx1; tx++)
            for(int ty = ty0; ty <= ty1; ty++)
                if(tx != tx0 || ty != ty0) atomicAdd(count + tx * p.tiles_y + ty, 1);
}

__global__ void __launch_bounds__(1024)
bin_scan_kernel(BinParams p)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int view = blockIdx.x;
    const int* count = p.tile_count + (size_t) view * p.ntiles;
    int* off = p.tile_off + (size_t) view * p.ntiles;
    int* cursor = p.tile_cursor + (size_t) view * p.ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(threadIdx.x == 0) carry = 0;
    __syncthreads();
    for(int base = 0; base < p.ntiles; base += 1024)
    {
        const int i = base + threadIdx.x;
        const int v = i < p.ntiles ? count[i] : 0;
        int incl = v;
        for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if(lane >= d) incl += n; }
        if(lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if(warp == 0)
        {
            int s = warp_sums[lane];
            for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, s, d); if(lane >= d) s += n; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int before = carry + (warp ? warp_sums[warp - 1] : 0) + incl - v;
        if(i < p.ntiles) { off[i] = before; cursor[i] = 0; }
        __syncthreads();
        if(threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if(threadIdx.x == 0)
    {
        p.totals[view] = carry;
        if(carry > p.cap) atomicOr(p.flags + view, FLAG_OVERFLOW);
    }
}

__global__ void __launch_bounds__(256)
bin_fill_kernel(BinParams p)
{
    const int view = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < p.ntri;
    ushort4 r = make_ushort4(1, 1, 0, 0);
    uint4 ent = make_uint4(0, 0, 0, 0);
    if(live)
    {
        r = p.rect[(size_t) view * p.ntri + t];
        ent = make_uint4(__ldg(p.i0 + t), __ldg(p.i1 + t), __ldg(p.i2 + t), (uint32_t) t);
    }
    const bool any = live && r.x <= r.z && r.y <= r.w;
    const int* off = p.tile_off + (size_t) view * p.ntiles;
    int* cursor = p.tile_cursor + (size_t) view * p.ntiles;
    uint4* entries = p.entries + (size_t) view * p.cap;
    const int first = any ? r.x * p.tiles_y + r.y : 0;
    const int slot0 = warp_agg_add(cursor, first, any);
    if(any)
    {
        const int s0 = off[first] + slot0;
        if(s0 < p.cap) entries[s0] = ent;
        for(int tx = r.x; tx <= r.z; tx++)
            for(int ty = r.y; ty <= r.w; ty++)
                if(tx != r.x || ty != r.y)
                {
                    const int tile = tx * p.tiles_y + ty;
                    const int s = off[tile] + atomicAdd(cursor + tile, 1);
                    if(s < p.cap) entries[s] = ent;
                }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* K3: tile rasteriser                                                                              */
/* ------------------------------------------------------------------------------------------------ */

struct RasterParams
{
    const float4* xf; const uint32_t *i0, *i1, *i2; const float2* uv;
    const int* tile_count; const int* tile_off; const uint4* entries;
    const uint32_t* tex; int tw, th;
    uint32_t* pixel; float* zbuf; unsigned long long* hash; uint32_t* flags; int* work_counter;
    int ntri, nuniq, xres, yres, tiles_x, tiles_y, ntiles, cap, nviews;
};

/* depth test + winner update on the tile's shared-memory key (main.c:356,365) */
__device__ __forceinline__ void key_update(unsigned long long* keys, int local, float z, uint32_t inv_tri)
{
    const unsigned long long key = ((unsigned long long) gel::zkey(z) << 32) | inv_tri;
    if(key > *reinterpret_cast<volatile unsigned long long*>(keys + local)) atomicMax(keys + local, key);
}

/* one pixel of the reference's inner loop (main.c:351-356) */
__device__ __forceinline__ void test_pixel(const gel::TriSetup& s, float sden, int x, int y, int local,
                                           unsigned long long* keys, uint32_t inv_tri)
{
    float nv, nw;
    gel::bary_numerators(s, gel::i2f(x), gel::i2f(y), nv, nw);
    if(gel::surely_negative(nv, sden) || gel::surely_negative(nw, sden)) return;
    float v, w, u, z;
    if(gel::bary_inside(s, nv, nw, v, w, u, z)) key_update(keys, local, z, inv_tri);
}

template<bool HASH>
__global__ void __launch_bounds__(RASTER_THREADS)
raster_kernel(RasterParams p)
{
    __shared__ unsigned long long keys[TW * TH];
    __shared__ int s_item;
    __shared__ unsigned long long s_hash[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nitems = p.nviews * p.ntiles;
    for(;;)
    {
        if(tid == 0) { s_item = atomicAdd(p.work_counter, 1); s_hash[0] = 0; s_hash[1] = 0; }
        __syncthreads();
        const int item = s_item;
        if(item >= nitems) break;
        const int view = item / p.ntiles, tile = item - view * p.ntiles;
        const int tx = tile / p.tiles_y, ty = tile - tx * p.tiles_y;
        const int px0 = tx * TW, py0 = ty * TH;
        const int px1 = min(px0 + TW, p.xres) - 1, py1 = min(py0 + TH, p.yres) - 1;
        const int count = min(p.tile_count[(size_t) view * p.ntiles + tile], p.cap);
        uint32_t* pixel = p.pixel + (size_t) view * p.xres * p.yres;
        float* zbuf = p.zbuf + (size_t) view * p.xres * p.yres;
        const float4* xf = p.xf + (size_t) view * p.nuniq;
        unsigned long long hp = 0, hz = 0;

        if(count == 0)
        {
            /* reset (main.c:413-417) for a tile nothing touches: straight to HBM */
            for(int i = tid; i < TW * TH; i += RASTER_THREADS)
            {
                const int x = px0 + (i >> 5), y = py0 + (i & 31);
                if(x <= px1 && y <= py1)
                {
                    const int idx = y + x * p.yres;
                    pixel[idx] = 0u;
                    zbuf[idx] = -FLT_MAX;
                    if(HASH) { hp += gel::salt_mix(0u, (uint32_t) idx); hz += gel::salt_mix(0xFF7FFFFFu, (uint32_t) idx); }
                }
            }
        }
        else
        {
            for(int i = tid; i < TW * TH; i += RASTER_THREADS) keys[i] = CLEAR_KEY;
            __syncthreads();

            /* ---- visibility: every (triangle, pixel) of main.c:348-356 that falls in this tile ---- */
            const uint4* entries = p.entries + (size_t) view * p.cap + p.tile_off[(size_t) view * p.ntiles + tile];
            for(int base = 0; base < count; base += RASTER_THREADS)
            {
                const int e = base + tid;
                const bool have = e < count;
                gel::TriSetup s;
                float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
                uint32_t inv_tri = 0;
                int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;
                if(have)
                {
                    const uint4 ent = __ldg(entries + e);
                    a = __ldg(xf + ent.x); b = __ldg(xf + ent.y); c = __ldg(xf + ent.z);
                    inv_tri = 0xFFFFFFFFu - ent.w;
                    s = gel::tri_setup(a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z);
                    bx0 = max(s.x0, px0); bx1 = min(s.x1, px1);
                    by0 = max(s.y0, py0); by1 = min(s.y1, py1);
                }
                const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
                const int npx = (bw > 0 && bh > 0) ? bw * bh : 0;
                const bool big = npx > SMALL_MAX;
                if(npx > 0 && !big)
                {
                    /* one lane walks its triangle's bbox, x outer / y inner like main.c:348-349 */
                    const float sden = gel::sign_guard(s.den);
                    int x = bx0, y = by0;
                    for(int i = 0; i < npx; i++)
                    {
                        test_pixel(s, sden, x, y, (x - px0) * TH + (y - py0), keys, inv_tri);
                        if(++y > by1) { y = by0; x++; }
                    }
                }
                /* large triangles: the whole warp sweeps one triangle, lanes along y */
                unsigned bigmask = __ballot_sync(0xFFFFFFFFu, big);
                while(bigmask)
                {
                    const int src = __ffs(bigmask) - 1;
                    bigmask &= bigmask - 1;
                    const float ax = __shfl_sync(0xFFFFFFFFu, a.x, src), ay = __shfl_sync(0xFFFFFFFFu, a.y, src), az = __shfl_sync(0xFFFFFFFFu, a.z, src);
                    const float bx = __shfl_sync(0xFFFFFFFFu, b.x, src), by = __shfl_sync(0xFFFFFFFFu, b.y, src), bz = __shfl_sync(0xFFFFFFFFu, b.z, src);
                    const float cx = __shfl_sync(0xFFFFFFFFu, c.x, src), cy = __shfl_sync(0xFFFFFFFFu, c.y, src), cz = __shfl_sync(0xFFFFFFFFu, c.z, src);
                    const uint32_t it = __shfl_sync(0xFFFFFFFFu, inv_tri, src);
                    const gel::TriSetup g = gel::tri_setup(ax, ay, az, bx, by, bz, cx, cy, cz);
                    const float sden = gel::sign_guard(g.den);
                    const int gx0 = max(g.x0, px0), gx1 = min(g.x1, px1);
                    const int gy0 = max(g.y0, py0), gy1 = min(g.y1, py1);
                    const int y = py0 + lane;
                    if(y >= gy0 && y <= gy1)
                        for(int x = gx0; x <= gx1; x++)
                            test_pixel(g, sden, x, y, (x - px0) * TH + lane, keys, it);
                }
            }
            __syncthreads();

            /* ---- shade the winner of every pixel once (main.c:358-366) and write the tile back ---- */
            for(int i = tid; i < TW * TH; i += RASTER_THREADS)
            {
                const int x = px0 + (i >> 5), y = py0 + (i & 31);
                if(x > px1 || y > py1) continue;
                const unsigned long long key = keys[i];
                uint32_t colour = 0u;
                float z = -FLT_MAX;
                if(key != CLEAR_KEY)
                {
                    const uint32_t tri = 0xFFFFFFFFu - (uint32_t) key;
                    z = gel::zkey_inv((uint32_t) (key >> 32));
                    const float4 a = __ldg(xf + __ldg(p.i0 + tri));
                    const float4 b = __ldg(xf + __ldg(p.i1 + tri));
                    const float4 c = __ldg(xf + __ldg(p.i2 + tri));
                    const gel::TriSetup s = gel::tri_setup(a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z);
                    float nv, nw, v, w, u, zz;
                    gel::bary_numerators(s, gel::i2f(x), gel::i2f(y), nv, nw);
                    gel::bary_inside(s, nv, nw, v, w, u, zz);
                    const float2 ta = __ldg(p.uv + 3 * (size_t) tri), tb = __ldg(p.uv + 3 * (size_t) tri + 1), tc = __ldg(p.uv + 3 * (size_t) tri + 2);
                    const float uv[6] = { ta.x, ta.y, tb.x, tb.y, tc.x, tc.y };
                    int xx, yy, shading;
                    gel::fragment_shade(v, w, u, uv, a.w, b.w, c.w, p.tw, p.th, xx, yy, shading);
                    if(xx < 0 || xx > p.tw - 1 || yy < 0 || yy > p.th - 1)
                    {
                        atomicOr(p.flags + view, FLAG_TEXCLAMP);   /* the reference reads out of bounds here (R) */
                        xx = min(max(xx, 0), p.tw - 1); yy = min(max(yy, 0), p.th - 1);
                    }
                    colour = gel::pshade(__ldg(p.tex + xx + yy * p.tw), shading);
                }
                const int idx = y + x * p.yres;
                pixel[idx] = colour;
                zbuf[idx] = z;
                if(HASH) { hp += gel::salt_mix(colour, (uint32_t) idx); hz += gel::salt_mix(__float_as_uint(z), (uint32_t) idx); }
            }
        }
        if(HASH)
        {
            for(int d = 16; d; d >>= 1) { hp += __shfl_xor_sync(0xFFFFFFFFu, hp, d); hz += __shfl_xor_sync(0xFFFFFFFFu, hz, d); }
            if(lane == 0) { atomicAdd(&s_hash[0], hp); atomicAdd(&s_hash[1], hz); }
            __syncthreads();
            if(tid == 0) { atomicAdd(p.hash + 2 * view, s_hash[0]); atomicAdd(p.hash + 2 * view + 1, s_hash[1]); }
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* host side                                                                                        */
/* ------------------------------------------------------------------------------------------------ */

template<typename T> void dfree(T*& p) { if(p) cudaFree(p); p = nullptr; }

} /* namespace */

struct gelcu_ctx
{
    int device = 0, xres = 0, yres = 0, tiles_x = 0, tiles_y = 0, ntiles = 0, num_sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    /* mesh */
    int ntri = 0, nuniq = 0; bool have_mesh = false;
    float4 *d_vpos = nullptr, *d_vnrm = nullptr; uint32_t *d_i0 = nullptr, *d_i1 = nullptr, *d_i2 = nullptr; float2* d_uv = nullptr;
    /* texture */
    uint32_t* d_tex = nullptr; int tw = 0, th = 0;
    /* per-batch work buffers */
    int batch_opt = 0, batch = 0, cap = 0, ctas_per_sm = 4, stage_timing = 1;
    float4* d_xf = nullptr; ushort4* d_rect = nullptr; int *d_tile_count = nullptr, *d_tile_off = nullptr, *d_tile_cursor = nullptr;
    uint4* d_entries = nullptr; uint32_t* d_flags = nullptr; int* d_totals = nullptr; unsigned long long* d_hash = nullptr; int* d_work = nullptr;
    uint32_t* d_pixel[2] = { nullptr, nullptr }; float* d_z[2] = { nullptr, nullptr };
    gelcu_view* d_views = nullptr; int views_cap = 0;
    int* h_totals = nullptr; uint32_t* h_flags = nullptr; int hcap = 0;
    std::vector<cudaEvent_t> ev;   /* 4 per batch */
    cudaEvent_t render_done[2] = { nullptr, nullptr }, copy_done[2] = { nullptr, nullptr };
    int last_batch_views = 0, last_buf = 0;
    gelcu_stats stats = {};
};

namespace {

void free_work(gelcu_ctx* c)
{
    dfree(c->d_xf); dfree(c->d_rect); dfree(c->d_tile_count); dfree(c->d_tile_off); dfree(c->d_tile_cursor);
    dfree(c->d_entries); dfree(c->d_flags); dfree(c->d_totals); dfree(c->d_hash); dfree(c->d_work);
    dfree(c->d_pixel[0]); dfree(c->d_pixel[1]); dfree(c->d_z[0]); dfree(c->d_z[1]);
    c->batch = 0; c->cap = 0;
}

size_t per_view_bytes(const gelcu_ctx* c, int cap)
{
    const size_t frame = (size_t) c->xres * c->yres;
    return 2 * frame * 8 + (size_t) c->nuniq * 16 + (size_t) c->ntri * 8 + (size_t) c->ntiles * 12 + (size_t) cap * 16 + 64;
}

int ensure_work(gelcu_ctx* c, int want_batch, int want_cap)
{
    if(c->batch >= want_batch && c->cap >= want_cap && c->d_xf) return GELCU_OK;
    free_work(c);
    const int B = want_batch, cap = want_cap;
    const size_t frame = (size_t) c->xres * c->yres;
    CU(cudaMalloc(&c->d_xf, sizeof(float4) * std::max<size_t>(1, (size_t) B * c->nuniq)));
    CU(cudaMalloc(&c->d_rect, sizeof(ushort4) * std::max<size_t>(1, (size_t) B * c->ntri)));
    CU(cudaMalloc(&c->d_tile_count, sizeof(int) * (size_t) B * c->ntiles));
    CU(cudaMalloc(&c->d_tile_off, sizeof(int) * (size_t) B * c->ntiles));
    CU(cudaMalloc(&c->d_tile_cursor, sizeof(int) * (size_t) B * c->ntiles));
    CU(cudaMalloc(&c->d_entries, sizeof(uint4) * std::max<size_t>(1, (size_t) B * cap)));
    CU(cudaMalloc(&c->d_flags, sizeof(uint32_t) * B));
    CU(cudaMalloc(&c->d_totals, sizeof(int) * B));
    CU(cudaMalloc(&c->d_hash, sizeof(unsigned long long) * 2 * B));
    CU(cudaMalloc(&c->d_work, sizeof(int)));
    for(int k = 0; k < 2; k++)
    {
        CU(cudaMalloc(&c->d_pixel[k], sizeof(uint32_t) * B * frame));
        CU(cudaMalloc(&c->d_z[k], sizeof(float) * B * frame));
    }
    c->batch = B; c->cap = cap;
    return GELCU_OK;
}

int default_batch(const gelcu_ctx* c, int cap)
{
    if(c->batch_opt > 0) return c->batch_opt;
    const size_t budget = (size_t) 24 << 30;
    const size_t pv = per_view_bytes(c, cap);
    return (int) std::min<size_t>(256, std::max<size_t>(1, budget / pv));
}

/* Enqueues K1..K3 for `n` views starting at d_views + first into frame buffer `buf`. */
int enqueue_batch(gelcu_ctx* c, int first, int n, int buf, bool want_hash, cudaEvent_t* ev4)
{
    cudaStream_t s = c->stream;
    CU(cudaMemsetAsync(c->d_tile_count, 0, sizeof(int) * (size_t) n * c->ntiles, s));
    CU(cudaMemsetAsync(c->d_flags, 0, sizeof(uint32_t) * n, s));
    CU(cudaMemsetAsync(c->d_work, 0, sizeof(int), s));
    if(want_hash) CU(cudaMemsetAsync(c->d_hash, 0, sizeof(unsigned long long) * 2 * n, s));
    if(ev4) CU(cudaEventRecord(ev4[0], s));
    if(c->nuniq > 0)
    {
        transform_kernel<<<dim3((c->nuniq + 255) / 256, n), 256, 0, s>>>(c->d_views + first, c->d_vpos, c->d_vnrm, c->d_xf, c->nuniq, c->xres, c->yres);
        c->stats.kernels_launched++;
    }
    if(ev4) CU(cudaEventRecord(ev4[1], s));
    BinParams bp = { c->d_xf, c->d_i0, c->d_i1, c->d_i2, c->d_rect, c->d_tile_count, c->d_tile_off, c->d_tile_cursor,
                     c->d_entries, c->d_flags, c->d_totals, c->ntri, c->nuniq, c->xres, c->yres, c->tiles_x, c->tiles_y, c->ntiles, c->cap };
    if(c->ntri > 0)
    {
        bin_count_kernel<<<dim3((c->ntri + 255) / 256, n), 256, 0, s>>>(bp);
        c->stats.kernels_launched++;
    }
    bin_scan_kernel<<<n, 1024, 0, s>>>(bp);
    c->stats.kernels_launched++;
    if(c->ntri > 0)
    {
        bin_fill_kernel<<<dim3((c->ntri + 255) / 256, n), 256, 0, s>>>(bp);
        c->stats.kernels_launched++;
    }
    if(ev4) CU(cudaEventRecord(ev4[2], s));
    RasterParams rp = { c->d_xf, c->d_i0, c->d_i1, c->d_i2, c->d_uv, c->d_tile_count, c->d_tile_off, c->d_entries,
                        c->d_tex, c->tw, c->th, c->d_pixel[buf], c->d_z[buf], c->d_hash, c->d_flags, c->d_work,
                        c->ntri, c->nuniq, c->xres, c->yres, c->tiles_x, c->tiles_y, c->ntiles, c->cap, n };
    const int items = n * c->ntiles;
    const int grid = std::max(1, std::min(items, c->num_sms * c->ctas_per_sm));
    if(want_hash) raster_kernel<true><<<grid, RASTER_THREADS, 0, s>>>(rp);
    else raster_kernel<false><<<grid, RASTER_THREADS, 0, s>>>(rp);
    c->stats.kernels_launched++;
    if(ev4) CU(cudaEventRecord(ev4[3], s));
    CU(cudaGetLastError());
    return GELCU_OK;
}

int ensure_events(gelcu_ctx* c, int nbatches)
{
    while((int) c->ev.size() < 4 * nbatches)
    {
        cudaEvent_t e; CU(cudaEventCreate(&e)); c->ev.push_back(e);
    }
    return GELCU_OK;
}

int ensure_host(gelcu_ctx* c, int n)
{
    if(c->hcap >= n) return GELCU_OK;
    if(c->h_totals) cudaFreeHost(c->h_totals);
    if(c->h_flags) cudaFreeHost(c->h_flags);
    CU(cudaMallocHost(&c->h_totals, sizeof(int) * n));
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
    if(xres <= 0 || yres <= 0 || xres > 16384 || yres > 16384) return fail(GELCU_E_INVALID, "resolution %dx%d out of range", xres, yres);
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess || n <= 0)
        return fail(GELCU_E_NOGPU, "no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count is 0");
    if(device < 0 || device >= n) return fail(GELCU_E_INVALID, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    gelcu_ctx* c = new gelcu_ctx();
    c->device = device; c->xres = xres; c->yres = yres;
    c->tiles_x = (xres + TW - 1) / TW; c->tiles_y = (yres + TH - 1) / TH; c->ntiles = c->tiles_x * c->tiles_y;
    cudaDeviceProp prop;
    if(cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    cudaError_t s1 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    cudaError_t s2 = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for(int k = 0; k < 2 && s1 == cudaSuccess && s2 == cudaSuccess; k++)
    {
        s1 = cudaEventCreateWithFlags(&c->render_done[k], cudaEventDisableTiming);
        s2 = cudaEventCreateWithFlags(&c->copy_done[k], cudaEventDisableTiming);
    }
    if(s1 != cudaSuccess || s2 != cudaSuccess) { delete c; return fail(GELCU_E_CUDA, "stream/event creation failed"); }
    *out = c;
    return GELCU_OK;
}

int gelcu_tile_grid(gelcu_ctx* c, int* tile_w, int* tile_h, int* tiles_x, int* tiles_y)
{
    if(!c) return fail(GELCU_E_INVALID, "null context");
    if(tile_w) *tile_w = TW; if(tile_h) *tile_h = TH;
    if(tiles_x) *tiles_x = c->tiles_x; if(tiles_y) *tiles_y = c->tiles_y;
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
    std::vector<float2> uv(ncorner);
    for(size_t cidx = 0; cidx < ncorner; cidx++) uv[cidx] = make_float2(tt[3 * cidx], tt[3 * cidx + 1]);   /* tt.z is never read, main.c:360-361 */

    free_work(c);
    dfree(c->d_vpos); dfree(c->d_vnrm); dfree(c->d_i0); dfree(c->d_i1); dfree(c->d_i2); dfree(c->d_uv);
    c->ntri = ntri; c->nuniq = (int) vpos.size(); c->have_mesh = true;
    const size_t nu = std::max<size_t>(1, vpos.size()), nt = std::max<size_t>(1, (size_t) ntri);
    CU(cudaMalloc(&c->d_vpos, sizeof(float4) * nu)); CU(cudaMalloc(&c->d_vnrm, sizeof(float4) * nu));
    CU(cudaMalloc(&c->d_i0, 4 * nt)); CU(cudaMalloc(&c->d_i1, 4 * nt)); CU(cudaMalloc(&c->d_i2, 4 * nt));
    CU(cudaMalloc(&c->d_uv, sizeof(float2) * 3 * nt));
    if(ntri > 0)
    {
        CU(cudaMemcpy(c->d_vpos, vpos.data(), sizeof(float4) * vpos.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_vnrm, vnrm.data(), sizeof(float4) * vnrm.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_i0, idx[0].data(), 4 * (size_t) ntri, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_i1, idx[1].data(), 4 * (size_t) ntri, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_i2, idx[2].data(), 4 * (size_t) ntri, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_uv, uv.data(), sizeof(float2) * uv.size(), cudaMemcpyHostToDevice));
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
    else if(!strcmp(name, "raster_ctas_per_sm")) { if(value < 1 || value > 32) return fail(GELCU_E_INVALID, "raster_ctas_per_sm out of [1,32]"); c->ctas_per_sm = value; }
    else if(!strcmp(name, "stage_timing")) c->stage_timing = value != 0;
    else return fail(GELCU_E_INVALID, "unknown option '%s'", name);
    return GELCU_OK;
}

int gelcu_get_stats(gelcu_ctx* c, gelcu_stats* out)
{
    if(!c || !out) return fail(GELCU_E_INVALID, "null argument");
    *out = c->stats;
    return GELCU_OK;
}

int gelcu_render(gelcu_ctx* c, const gelcu_view* views, int nviews,
                 uint32_t* pixel_out, float* z_out, uint64_t* hash_out, float* device_ms)
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
    int cap = c->cap > 0 ? c->cap : (int) std::min<size_t>((size_t) 1 << 30, (size_t) 2 * c->ntri + 4096);

    for(int attempt = 0; attempt < 3; attempt++)
    {
        const int B = std::min(nviews, std::max(c->batch, default_batch(c, cap)));
        rc = ensure_work(c, B, cap); if(rc) return rc;
        const int nb = (nviews + c->batch - 1) / c->batch;
        rc = ensure_events(c, nb); if(rc) return rc;
        c->stats.kernels_launched = 0; c->stats.h2d_bytes = 0; c->stats.d2h_bytes = 0; c->stats.batches = nb; c->stats.views = nviews;
        CU(cudaMemcpyAsync(c->d_views, views, sizeof(gelcu_view) * nviews, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += sizeof(gelcu_view) * (size_t) nviews;

        auto issue_copies = [&](int b) -> int {
            const int buf = b & 1, first = b * c->batch, n = std::min(c->batch, nviews - first);
            CU(cudaStreamWaitEvent(c->copy_stream, c->render_done[buf], 0));
            if(pixel_out) { CU(cudaMemcpyAsync(pixel_out + frame * first, c->d_pixel[buf], 4 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * frame * n; }
            if(z_out) { CU(cudaMemcpyAsync(z_out + frame * first, c->d_z[buf], 4 * frame * n, cudaMemcpyDeviceToHost, c->copy_stream)); c->stats.d2h_bytes += 4 * frame * n; }
            CU(cudaEventRecord(c->copy_done[buf], c->copy_stream));
            return GELCU_OK;
        };

        for(int b = 0; b < nb; b++)
        {
            const int buf = b & 1, first = b * c->batch, n = std::min(c->batch, nviews - first);
            if(b >= 2) CU(cudaStreamWaitEvent(c->stream, c->copy_done[buf], 0));
            rc = enqueue_batch(c, first, n, buf, hash_out != nullptr, &c->ev[4 * b]); if(rc) return rc;
            /* small per-batch results ride the render stream (they are overwritten by the next batch) */
            CU(cudaMemcpyAsync(c->h_totals + first, c->d_totals, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaMemcpyAsync(c->h_flags + first, c->d_flags, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
            if(hash_out) { CU(cudaMemcpyAsync(hash_out + 2 * (size_t) first, c->d_hash, 16 * (size_t) n, cudaMemcpyDeviceToHost, c->stream)); c->stats.d2h_bytes += 16 * (size_t) n; }
            CU(cudaEventRecord(c->render_done[buf], c->stream));
            if(b >= 1 && (pixel_out || z_out)) { rc = issue_copies(b - 1); if(rc) return rc; }
            c->last_batch_views = n; c->last_buf = buf;
        }
        if(pixel_out || z_out) { rc = issue_copies(nb - 1); if(rc) return rc; }
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaStreamSynchronize(c->copy_stream));

        uint32_t flags = 0; int max_total = 0; uint64_t entries = 0;
        for(int v = 0; v < nviews; v++) { flags |= c->h_flags[v]; max_total = std::max(max_total, c->h_totals[v]); entries += c->h_totals[v]; }
        if(flags & FLAG_OVERFLOW)
        {
            /* a tile list outgrew its pool: grow to the measured need and render the call again */
            cap = (int) std::min<size_t>((size_t) 1 << 30, (size_t) max_total + max_total / 8 + 1024);
            if(per_view_bytes(c, cap) > ((size_t) 150 << 30))
                return fail(GELCU_E_NOMEM, "bin lists need %d entries per view, beyond device memory", max_total);
            free_work(c);
            continue;
        }
        float ms = 0.0f, t = 0.0f;
        CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[4 * (nb - 1) + 3]));
        c->stats.ms_total = ms;
        if(c->stage_timing)
            for(int b = 0; b < nb; b++)
            {
                CU(cudaEventElapsedTime(&t, c->ev[4 * b], c->ev[4 * b + 1])); c->stats.ms_transform += t;
                CU(cudaEventElapsedTime(&t, c->ev[4 * b + 1], c->ev[4 * b + 2])); c->stats.ms_bin += t;
                CU(cudaEventElapsedTime(&t, c->ev[4 * b + 2], c->ev[4 * b + 3])); c->stats.ms_raster += t;
            }
        if(device_ms) *device_ms = ms;
        c->stats.bin_entries = entries;
        c->stats.flags = flags & ~FLAG_OVERFLOW;
        if(c->stats.flags) return fail(GELCU_W_CLIPPED, "input left the reference's defined domain (flags 0x%x: 1 = bbox off screen, 2 = texel out of range); clipped", c->stats.flags);
        return GELCU_OK;
    }
    return fail(GELCU_E_NOMEM, "bin list capacity did not converge");
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
    rc = gelcu_render(c, view, 1, nullptr, nullptr, nullptr, nullptr);
    if(rc < 0) return rc;
    std::vector<int> cnt(c->ntiles), off(c->ntiles);
    CU(cudaMemcpy(cnt.data(), c->d_tile_count, sizeof(int) * c->ntiles, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(off.data(), c->d_tile_off, sizeof(int) * c->ntiles, cudaMemcpyDeviceToHost));
    const int tot = c->h_totals[0];
    std::vector<uint4> ent(std::max(1, tot));
    CU(cudaMemcpy(ent.data(), c->d_entries, sizeof(uint4) * tot, cudaMemcpyDeviceToHost));
    if(total) *total = tot;
    int w = 0;
    for(int t = 0; t < c->ntiles; t++)
    {
        if(counts) counts[t] = cnt[t];
        std::vector<int> tris(cnt[t]);
        for(int k = 0; k < cnt[t]; k++) tris[k] = (int) ent[off[t] + k].w;
        std::sort(tris.begin(), tris.end());
        for(int k = 0; k < cnt[t] && entries && w < cap; k++) entries[w++] = tris[k];
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
    dfree(c->d_vpos); dfree(c->d_vnrm); dfree(c->d_i0); dfree(c->d_i1); dfree(c->d_i2); dfree(c->d_uv);
    dfree(c->d_tex); dfree(c->d_views);
    if(c->h_totals) cudaFreeHost(c->h_totals);
    if(c->h_flags) cudaFreeHost(c->h_flags);
    for(cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for(int k = 0; k < 2; k++) { if(c->render_done[k]) cudaEventDestroy(c->render_done[k]); if(c->copy_done[k]) cudaEventDestroy(c->copy_done[k]); }
    if(c->stream) cudaStreamDestroy(c->stream);
    if(c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

} /* extern "C" */
