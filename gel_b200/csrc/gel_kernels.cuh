/* gel_kernels.cuh -- the three sm_100a kernels of gel's per-frame render path (device code only).
 *
 *   K1 transform_kernel   one thread per (view, distinct corner): tviewnrm / tviewtri / tperspective / tviewport
 *                         (main.c:372-390, 302-314, 288-300); float4 in, float4 (screen x, y, z, shade) out
 *   K2 bin_kernel         triangle setup + screen-tile binning in ONE pass: a CTA takes 1024 consecutive
 *                         triangles, computes each bbox (main.c:344-347) and the per-triangle part of tbarycenter
 *                         (main.c:319-324) into one 128-byte record per (view, triangle), groups the (triangle, tile)
 *                         pairs by tile in shared memory (count -> scan -> place) and publishes one SEGMENT per touched
 *                         tile: a contiguous run of entries in a per-view pool, pushed on the tile's chain; lit tiles
 *                         go on one batch-wide work list
 *   K3 raster_kernel      persistent CTAs pull (view, tile) items; the tile's depth + winner live in shared
 *                         memory as one 64-bit key per pixel.  Small triangles are expanded into FRAGMENTS
 *                         (one lane per bbox pixel, so lanes stay busy whatever the triangle sizes); large ones
 *                         are swept by the whole CTA with every pixel owned by one thread.  The winning fragment
 *                         of each pixel is shaded once (main.c:358-366) and the tile goes back to HBM in one
 *                         coalesced pass.
 *
 * Draw-order semantics (main.c:356: strict `z > zbuff`, so the FIRST submitted triangle wins a tie) are kept
 * exactly by resolving  key = zkey(z) << 32 | (0xFFFFFFFF - triangle_index)  with a 64-bit max:
 * "first triangle in submission order to reach a strictly greater z" == "greatest z, ties to the lowest index".
 * The result therefore does not depend on the order in which a tile's triangles are visited.
 */
#ifndef GEL_KERNELS_CUH
#define GEL_KERNELS_CUH

#include "../../include/gelcu.h"
#include "gel_math.h"

#include <cuda_runtime.h>
#include <cuda.h>          /* CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link) */
#include <climits>

namespace gelk {

#ifndef GEL_TW
#define GEL_TW 32
#endif
constexpr int TW = GEL_TW;         /* tile width  (screen x, the framebuffer's SLOW axis)                */
constexpr int TH = 32;             /* tile height (screen y, contiguous in memory: index y + x*yres)     */
#ifndef GEL_RASTER_THREADS
#define GEL_RASTER_THREADS 128
#endif
constexpr int RASTER_THREADS = GEL_RASTER_THREADS;
constexpr int RASTER_WARPS = RASTER_THREADS / 32;
#ifndef GEL_FRAG_MAX
#define GEL_FRAG_MAX 256
#endif
#ifndef GEL_ZSPLIT_TILE
#define GEL_ZSPLIT_TILE 0.4f          /* near / far split of a view as a fraction of its depth range: a speed heuristic, any value is exact */
#endif
#ifndef GEL_UNIT_WINDOW
#define GEL_UNIT_WINDOW 128
#endif
#ifndef GEL_DEFER_MAX
#define GEL_DEFER_MAX 48
#endif
#ifndef GEL_RASTER_MINB
#define GEL_RASTER_MINB 8          /* resident rasteriser CTAs per SM the kernel is compiled for (registers) and its shared memory is sized for */
#endif
#ifndef GEL_TWO_PHASE_MIN
#define GEL_TWO_PHASE_MIN 8
#endif
constexpr int FRAG_MAX = GEL_FRAG_MAX;      /* bbox-in-tile pixels up to which a triangle goes through the per-warp unit path */
constexpr int UNIT_WINDOW = GEL_UNIT_WINDOW;   /* column units (one bbox column of one triangle) staged per warp per pass */
constexpr int QCAP = 64;           /* survivor stack per warp: < 32 left over + one row of 32 lanes       */
constexpr int FAR_CAP = 4096;       /* far triangles a CTA can park per tile (16 B each, global scratch)   */
constexpr int TWO_PHASE_MIN = GEL_TWO_PHASE_MIN;  /* tiles with fewer entries are rasterised in one phase                */
constexpr int MAX_BATCH = 256;     /* views per launch set (the work list packs the view in 8 bits)         */
constexpr int DEFER_MAX = GEL_DEFER_MAX;      /* large triangles per round left to the CTA-wide sweep (their setup records wait in shared memory) */
#ifndef GEL_RESET_BOX_COLS
#define GEL_RESET_BOX_COLS 16
#endif
static_assert(GEL_TW % GEL_RESET_BOX_COLS == 0, "a tile is a whole number of reset boxes");
constexpr int RESET_BOX_COLS = GEL_RESET_BOX_COLS;  /* tile columns one TMA tensor store resets (box = 32 rows x 16 columns = 2 KB of pattern): the copy unit takes its
                                     * operands from uniform registers, so every store is issued by ONE lane at a time (ptxas loops over the active
                                     * lanes, ~15 instructions per store) -- few large boxes, not many small ones */
constexpr int CLEAR_CHUNK = 8;     /* tiles a CTA checks (and resets when untouched) per work item        */
constexpr int SEG_SLOTS = RASTER_THREADS;   /* segments staged per round (one per thread)                 */
constexpr int NCHAIN = 8;          /* parallel segment chains per tile (chunk % NCHAIN)                   */
constexpr int BIN_THREADS = 256;
constexpr int BIN_TPT = 4;         /* triangles per thread in K2                                          */
constexpr int BIN_CHUNK = BIN_THREADS * BIN_TPT;
constexpr int LOCAL_MAX = 1024;    /* tile slots a K2 CTA can group locally                               */
constexpr int HUGE_TILES = 16;     /* triangles covering more tiles than this are published tile by tile  */
constexpr unsigned long long CLEAR_KEY = (0x00800000ull << 32) | 0xFFFFFFFFull;   /* zkey(-FLT_MAX), no winner */
constexpr uint32_t FLAG_CLIPPED = 1u, FLAG_TEXCLAMP = 2u, FLAG_OVERFLOW = 0x80000000u;
constexpr float GUARD_EPS = 1e-20f, GUARD_DEN_MAX = 1e18f, U_SLACK = 1.00001f;

/* 32 bytes in one request (sm_100: ld.global.nc.v8, SASS LDG.E.ENL2.256): a 32-byte-aligned pair of quads costs the L1 one sector
 * lookup instead of two.  `p` must be 32-byte aligned.  Used by the direct pipeline's resolve pass, which is bound by L1 sector
 * lookups (its 32-byte triangle record: l1tex__t_sectors -14 %, the kernel -3.6 %); in the band rasteriser, which is not, the same
 * loads measured -0.4 % and stayed 128-bit. */
#ifndef GEL_LD256
#define GEL_LD256 1
#endif
__device__ __forceinline__ void ldg256(const void* p, float4& lo, float4& hi)
{
#if GEL_LD256
    uint32_t a, b, c, d, e, f, g, h;
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    lo = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
    hi = make_float4(__uint_as_float(e), __uint_as_float(f), __uint_as_float(g), __uint_as_float(h));
#else
    lo = __ldg(reinterpret_cast<const float4*>(p)); hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
#endif
}

/* ------------------------------------------------------------------------------------------------ */
/* K1: vertex transform                                                                             */
/* ------------------------------------------------------------------------------------------------ */

/* per-view statistics K1 leaves for the rasterisers: words 0-1 range of zkey(screen z), 2-5 screen bbox of the
 * vertices (int min/max of the truncated x and y), 6 parked-triangle counter of the direct pipeline */
constexpr int VIEW_STAT_WORDS = 8;

/* A view's REGION: the screen bounding box of its transformed vertices (words 2-5), clipped to the frame and widened to
 * multiples of 8.  Every triangle bbox of main.c:344-347 lies inside it, so everything outside holds reset values.  Host
 * and device use this one definition (gelcu_render_region reports it; the direct pipeline resolves / fills by it). */
__host__ __device__ __forceinline__ bool region_from_stats(const uint32_t* s, int xres, int yres, int& x0, int& x1, int& y0, int& y1)
{
    const int sx0 = (int) s[2], sx1 = (int) s[3], sy0 = (int) s[4], sy1 = (int) s[5];
    x0 = (sx0 > 0 ? sx0 : 0) & ~7; y0 = (sy0 > 0 ? sy0 : 0) & ~7;
    x1 = (sx1 < xres - 1 ? sx1 : xres - 1) | 7; if(x1 > xres - 1) x1 = xres - 1;
    y1 = (sy1 < yres - 1 ? sy1 : yres - 1) | 7; if(y1 > yres - 1) y1 = yres - 1;
    return x0 <= x1 && y0 <= y1;
}

constexpr int XF_PER_THREAD = 4;     /* vertices per thread: amortises the per-warp reductions of the view statistics */

__global__ void __launch_bounds__(256)
transform_kernel(const float4* __restrict__ vconst, const float4* __restrict__ vpos,
                 const float4* __restrict__ vnrm, float4* __restrict__ xf, uint32_t* __restrict__ vstat, int nuniq, int xres, int yres)
{
    __shared__ uint32_t s_lo, s_hi;
    __shared__ int s_box[4];
    const int view = blockIdx.y;
    if(threadIdx.x == 0) { s_lo = 0xFFFFFFFFu; s_hi = 0u; s_box[0] = INT_MAX; s_box[1] = INT_MIN; s_box[2] = INT_MAX; s_box[3] = INT_MIN; }
    __syncthreads();
    /* the view's constants (basis rows, vdot(row, eye), viewport scale / offset: main.c:375-377, 290-293) were computed once per view by
     * batch_init_kernel -- view_const() per thread was a sixth of this kernel's instructions (four div.rn + three dot products) */
    gel::ViewConst c;
    {
        const float4 c0 = __ldg(vconst + 4 * view), c1 = __ldg(vconst + 4 * view + 1), c2 = __ldg(vconst + 4 * view + 2), c3 = __ldg(vconst + 4 * view + 3);
        c.xx = c0.x; c.xy = c0.y; c.xz = c0.z; c.yx = c0.w; c.yy = c1.x; c.yz = c1.y; c.zx = c1.z; c.zy = c1.w;
        c.zz = c2.x; c.xe = c2.y; c.ye = c2.z; c.ze = c2.w; c.w = c3.x; c.h = c3.y; c.x0 = c3.z; c.y0 = c3.w;
    }
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    int x0 = INT_MAX, x1 = INT_MIN, y0 = INT_MAX, y1 = INT_MIN;
    #pragma unroll
    for(int k = 0; k < XF_PER_THREAD; k++)
    {
        const int i = (blockIdx.x * XF_PER_THREAD + k) * blockDim.x + threadIdx.x;      /* coalesced float4 loads and stores */
        if(i < nuniq)
        {
            const float4 p = __ldg(vpos + i);
            const float4 n = __ldg(vnrm + i);
            float4 o;
            gel::transform_corner(c, p.x, p.y, p.z, n.x, n.y, n.z, o.x, o.y, o.z, o.w);
            xf[(size_t) view * nuniq + i] = o;
            if(o.z == o.z) { const uint32_t zk = gel::zkey(o.z); lo = min(lo, zk); hi = max(hi, zk); }
            const int ix = gel::trunc_i(o.x), iy = gel::trunc_i(o.y);
            x0 = min(x0, ix); x1 = max(x1, ix); y0 = min(y0, iy); y1 = max(y1, iy);
        }
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    x0 = __reduce_min_sync(0xFFFFFFFFu, x0); x1 = __reduce_max_sync(0xFFFFFFFFu, x1);
    y0 = __reduce_min_sync(0xFFFFFFFFu, y0); y1 = __reduce_max_sync(0xFFFFFFFFu, y1);
    if((threadIdx.x & 31) == 0)
    {
        atomicMin(&s_lo, lo); atomicMax(&s_hi, hi);
        atomicMin(&s_box[0], x0); atomicMax(&s_box[1], x1); atomicMin(&s_box[2], y0); atomicMax(&s_box[3], y1);
    }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        uint32_t* s = vstat + (size_t) view * VIEW_STAT_WORDS;
        atomicMin(s, s_lo); atomicMax(s + 1, s_hi);
        atomicMin(reinterpret_cast<int*>(s + 2), s_box[0]); atomicMax(reinterpret_cast<int*>(s + 3), s_box[1]);
        atomicMin(reinterpret_cast<int*>(s + 4), s_box[2]); atomicMax(reinterpret_cast<int*>(s + 5), s_box[3]);
    }
}

/* Batch initialisation in ONE launch (it used to be an H2D copy of the initial statistics and up to six memsets, each a
 * launch of its own -- a third of the device time of a single-view call): per-view statistics, pool cursors, flags, checksums
 * and, for the tile pipeline, chain heads (-1), lit flags and the two work queues. */
__global__ void __launch_bounds__(256)
batch_init_kernel(uint32_t* __restrict__ vstat, int* __restrict__ cursors, uint32_t* __restrict__ flags, unsigned long long* __restrict__ hash,
                  int* __restrict__ heads, int* __restrict__ tile_lit, int* __restrict__ work, int nviews, int ntiles, int want_hash,
                  const gelcu_view* __restrict__ views, float4* __restrict__ vconst, int xres, int yres)
{
    const size_t i0 = (size_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t) gridDim.x * blockDim.x;
    if(i0 < (size_t) nviews)
    {
        uint32_t* w = vstat + VIEW_STAT_WORDS * i0;
        w[0] = 0xFFFFFFFFu; w[1] = 0u;                                     /* depth range (min of keys, max of keys) */
        w[2] = 0x7FFFFFFFu; w[3] = 0x80000000u; w[4] = 0x7FFFFFFFu; w[5] = 0x80000000u;   /* screen bbox (int min / max) */
        w[6] = 0u; w[7] = 0u;
        cursors[4 * i0] = 0; cursors[4 * i0 + 1] = 0; cursors[4 * i0 + 2] = 0; cursors[4 * i0 + 3] = 0;
        flags[i0] = 0u;
        if(want_hash) { hash[2 * i0] = 0ull; hash[2 * i0 + 1] = 0ull; }
        /* the view's constants for K1 (same operations as ever: gel::view_const) */
        const gel::ViewConst c = gel::view_const(reinterpret_cast<const float*>(views + i0), xres, yres);
        vconst[4 * i0] = make_float4(c.xx, c.xy, c.xz, c.yx); vconst[4 * i0 + 1] = make_float4(c.yy, c.yz, c.zx, c.zy);
        vconst[4 * i0 + 2] = make_float4(c.zz, c.xe, c.ye, c.ze); vconst[4 * i0 + 3] = make_float4(c.w, c.h, c.x0, c.y0);
    }
    if(heads)
    {
        if(i0 < 3) work[i0] = 0;
        const size_t nlit = (size_t) nviews * ntiles, nheads = nlit * NCHAIN;
        for(size_t i = i0; i < nheads; i += stride) heads[i] = -1;
        for(size_t i = i0; i < nlit; i += stride) tile_lit[i] = 0;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* K2: setup + binning                                                                              */
/* ------------------------------------------------------------------------------------------------ */

/* Per-(view, triangle) record K2 leaves for the rasteriser: everything of tbarycenter / tdraw that depends on the triangle
 * and the view but not on the pixel (main.c:319-324, 344-347), computed ONCE here -- where the three corners are loaded anyway
 * for the bbox -- instead of once per (triangle, tile) in the visibility pass and once per lit PIXEL in the shade pass.
 * 8 x 16 bytes = one 128-byte line per triangle, so whatever part a consumer needs arrives with one L2 request:
 *   0  ax, ay, v0x, v0y            1  v1x, v1y, k0, k1           2  d00, d01, d11, den (sign-normalised, den > 0)
 *   3  az, bz, cz, ~tri            4  bbox x0 | x1 << 16, y0 | y1 << 16 (clipped to the frame), flags (1 drawable, 2 guard), zmax
 *   5  shade a, b, c, -            6  ta.x, ta.y, tb.x, tb.y     7  tc.x, tc.y, e_v, e_w (slack terms of the band rasteriser's row trimming: gel_math.h, trim_slack)
 * Same operations on the same operands as before (gel::tri_setup), so the frames do not change by a bit. */
constexpr int VREC_QUADS = 8;

struct BinParams
{
    const float4* xf; const uint32_t *i0, *i1, *i2;
    const float2* uv;    /* [3 * ntri] texture coordinates (copied into the records)                       */
    float4* vrec;        /* [view][ntri][VREC_QUADS]  per-view triangle records (null: not wanted)         */
    uint32_t* entries;   /* [view][cap_e]   triangle index                                                */
    uint4* descs;        /* [view][cap_d]   (next desc or -1, first entry, entry count, chunk)            */
    int* heads;          /* [view][ntiles][NCHAIN]  top of each chain, -1 = empty                         */
    int* cursors;        /* [view][4]       entries used, descs used (both keep counting past the capacity), lit tiles, - */
    int* tile_lit;       /* [view][ntiles]  1 once a segment was published for the tile                           */
    uint32_t* lit_list;  /* [nviews * ntiles]  the batch's lit tiles (view << 24 | tile) in publication order: K3's work list */
    int* work;           /* [2] = number of lit tiles published                                                          */
    uint32_t* flags;     /* [view]                                                                        */
    int ntri, nuniq, xres, yres, tiles_x, tiles_y, ntiles, cap_e, cap_d;
};

__device__ __forceinline__ void publish_segment(const BinParams& p, int view, int tile, int chunk, int id, int first, int count)
{
    int* head = p.heads + ((size_t) view * p.ntiles + tile) * NCHAIN + (chunk % NCHAIN);
    const int prev = atomicExch(head, id);
    p.descs[(size_t) view * p.cap_d + id] = make_uint4((uint32_t) prev, (uint32_t) first, (uint32_t) count, (uint32_t) chunk);
    if(prev < 0 && atomicExch(p.tile_lit + (size_t) view * p.ntiles + tile, 1) == 0)
        p.lit_list[atomicAdd(p.work + 2, 1)] = (uint32_t) view << 24 | (uint32_t) tile;
}

__global__ void __launch_bounds__(BIN_THREADS)
bin_kernel(BinParams p)
{
    /* one buffer, two lives: first the warps' record staging slabs (4.5 KB each), then -- after the barrier that follows the
     * triangle loop -- the three grouping arrays */
    __shared__ float4 s_rec[BIN_THREADS / 32][32 * (VREC_QUADS + 1)];
    static_assert(sizeof(float4) * (BIN_THREADS / 32) * 32 * (VREC_QUADS + 1) >= 3 * sizeof(int) * LOCAL_MAX, "the grouping arrays alias the staging slabs");
    int* const s_cnt = reinterpret_cast<int*>(&s_rec[0][0]);   /* [LOCAL_MAX] pairs per local tile slot                  */
    int* const s_pre = s_cnt + LOCAL_MAX;                      /* [LOCAL_MAX] exclusive prefix: entries | segments << 20 */
    int* const s_fil = s_pre + LOCAL_MAX;                      /* [LOCAL_MAX] placement cursor per slot                  */
    __shared__ int s_rect[4];            /* CTA bounding tile rect of its normal triangles    */
    __shared__ int s_warp[BIN_THREADS / 32];
    __shared__ int s_ebase, s_dbase, s_ok;

    const int view = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4* xf = p.xf + (size_t) view * p.nuniq;
    if(tid == 0) { s_rect[0] = 1 << 30; s_rect[1] = 1 << 30; s_rect[2] = -1; s_rect[3] = -1; }
    __syncthreads();

    /* ---- per triangle: bbox (main.c:344-347) -> tile rect ---- */
    uint32_t rect[BIN_TPT];                /* tx0 | ty0 << 8 | tx1 << 16 | ty1 << 24 (tile coordinates < 256: gelcu_create) */
    unsigned valid = 0;                    /* bit k: triangle k has something to bin */
    int minx = 1 << 30, miny = 1 << 30, maxx = -1, maxy = -1;
    bool clipped = false;
    #pragma unroll
    for(int k = 0; k < BIN_TPT; k++)
    {
        const int t = chunk * BIN_CHUNK + k * BIN_THREADS + tid;
        rect[k] = 0;
        if(t < p.ntri)
        {
            const float4 a = __ldg(xf + __ldg(p.i0 + t)), b = __ldg(xf + __ldg(p.i1 + t)), c = __ldg(xf + __ldg(p.i2 + t));
            int x0 = gel::trunc_i(fminf(a.x, fminf(b.x, c.x)));
            int y0 = gel::trunc_i(fminf(a.y, fminf(b.y, c.y)));
            int x1 = gel::trunc_i(fmaxf(a.x, fmaxf(b.x, c.x)));
            int y1 = gel::trunc_i(fmaxf(a.y, fmaxf(b.y, c.y)));
            if(x0 < 0 || y0 < 0 || x1 > p.xres - 1 || y1 > p.yres - 1)
            {
                clipped = true;            /* the reference writes out of bounds here */
                x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, p.xres - 1); y1 = min(y1, p.yres - 1);
            }
            if(p.vrec)
            {
                const gel::TriSetup s = gel::tri_setup(a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z);
                const float ad = fabsf(s.den);
                const bool drawable = ad > 0.0f && x0 <= x1 && y0 <= y1;   /* den == 0 or NaN can never pass main.c:352 */
                const float sg = s.den < 0.0f ? -1.0f : 1.0f;              /* exact sign flips */
                const float2 ta = __ldg(p.uv + 3 * (size_t) t), tb = __ldg(p.uv + 3 * (size_t) t + 1), tc = __ldg(p.uv + 3 * (size_t) t + 2);
                const uint32_t bx = (uint32_t) (x0 & 0xFFFF) | (uint32_t) (x1 & 0xFFFF) << 16, by = (uint32_t) (y0 & 0xFFFF) | (uint32_t) (y1 & 0xFFFF) << 16;
                /* the record goes through the warp's staging slab (stride 9 quads: conflict-free) so that the warp writes its 32
                 * consecutive records as eight fully coalesced 512-byte stores instead of 256 scattered 16-byte ones */
                float4* r = s_rec[warp] + lane * (VREC_QUADS + 1);
                r[0] = make_float4(s.ax, s.ay, s.v0x, s.v0y);
                r[1] = make_float4(s.v1x, s.v1y, s.k0, s.k1);
                r[2] = make_float4(s.d00 * sg, s.d01 * sg, s.d11 * sg, s.den * sg);
                r[3] = make_float4(s.az, s.bz, s.cz, __uint_as_float(0xFFFFFFFFu - (uint32_t) t));
                r[4] = make_float4(__uint_as_float(bx), __uint_as_float(by), __uint_as_float((drawable ? 1u : 0u) | (ad <= GUARD_DEN_MAX ? 2u : 0u)),
                                   fmaxf(a.z, fmaxf(b.z, c.z)));
                r[5] = make_float4(a.w, b.w, c.w, 0.0f);
                r[6] = make_float4(ta.x, ta.y, tb.x, tb.y);
                /* slack terms of the band rasteriser's row trimming (gel_math.h: row_trim) over the whole clipped bbox */
                float ev = INFINITY, ew = INFINITY;
                if(drawable) gel::trim_slack(s.ax, s.ay, s.v0x, s.v0y, s.v1x, s.v1y, s.k0, s.k1, s.d00 * sg, s.d01 * sg, s.d11 * sg, s.den * sg, x0, y0, x1, y1, ev, ew);
                r[7] = make_float4(tc.x, tc.y, ev, ew);
            }
            if(x0 <= x1 && y0 <= y1)
            {
                const int tx0 = x0 / TW, ty0 = y0 / TH, tx1 = x1 / TW, ty1 = y1 / TH;
                rect[k] = (uint32_t) tx0 | (uint32_t) ty0 << 8 | (uint32_t) tx1 << 16 | (uint32_t) ty1 << 24;
                valid |= 1u << k;
                if((tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= HUGE_TILES)
                { minx = min(minx, tx0); miny = min(miny, ty0); maxx = max(maxx, tx1); maxy = max(maxy, ty1); }
            }
        }
        if(p.vrec)
        {
            /* the warp's records of this round: triangles [t0, t0 + 32) below ntri */
            __syncwarp();
            const int t0 = chunk * BIN_CHUNK + k * BIN_THREADS + warp * 32;
            const int nrec = min(32, p.ntri - t0);
            float4* dst = p.vrec + ((size_t) view * p.ntri + t0) * VREC_QUADS;
            #pragma unroll
            for(int j = 0; j < VREC_QUADS; j++)
            {
                const int q = j * 32 + lane;
                if((q >> 3) < nrec) dst[q] = s_rec[warp][(q >> 3) * (VREC_QUADS + 1) + (q & 7)];
            }
            __syncwarp();
        }
    }
    if(__any_sync(0xFFFFFFFFu, clipped) && lane == 0) atomicOr(p.flags + view, FLAG_CLIPPED);
    for(int d = 16; d; d >>= 1)
    {
        minx = min(minx, __shfl_xor_sync(0xFFFFFFFFu, minx, d)); miny = min(miny, __shfl_xor_sync(0xFFFFFFFFu, miny, d));
        maxx = max(maxx, __shfl_xor_sync(0xFFFFFFFFu, maxx, d)); maxy = max(maxy, __shfl_xor_sync(0xFFFFFFFFu, maxy, d));
    }
    if(lane == 0 && maxx >= 0) { atomicMin(&s_rect[0], minx); atomicMin(&s_rect[1], miny); atomicMax(&s_rect[2], maxx); atomicMax(&s_rect[3], maxy); }
    __syncthreads();
    const int rx0 = s_rect[0], ry0 = s_rect[1];
    const int RW = s_rect[2] - rx0 + 1, RH = s_rect[3] - ry0 + 1;
    const int nslots = s_rect[2] >= 0 ? RW * RH : 0;
    const bool local = nslots > 0 && nslots <= LOCAL_MAX;
    __syncthreads();

    if(local)
    {
        for(int s = tid; s < nslots; s += BIN_THREADS) { s_cnt[s] = 0; s_fil[s] = 0; }
        __syncthreads();
        /* count */
        #pragma unroll
        for(int k = 0; k < BIN_TPT; k++)
        {
            if(!(valid >> k & 1)) continue;
            const int tx0 = rect[k] & 255, ty0 = (rect[k] >> 8) & 255, tx1 = (rect[k] >> 16) & 255, ty1 = (rect[k] >> 24) & 255;
            if((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > HUGE_TILES) continue;
            for(int tx = tx0; tx <= tx1; tx++)
                for(int ty = ty0; ty <= ty1; ty++) atomicAdd(&s_cnt[(tx - rx0) * RH + (ty - ry0)], 1);
        }
        __syncthreads();
        /* exclusive scan of (count | nonzero << 20) over the slots, 4 consecutive slots per thread */
        int v[4], sum = 0;
        #pragma unroll
        for(int j = 0; j < 4; j++)
        {
            const int s = tid * 4 + j;
            const int cn = s < nslots ? s_cnt[s] : 0;
            v[j] = cn | (cn > 0 ? 1 << 20 : 0);
            sum += v[j];
        }
        int incl = sum;
        for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if(lane >= d) incl += n; }
        if(lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
        #pragma unroll
        for(int w = 0; w < BIN_THREADS / 32; w++) { const int x = s_warp[w]; if(w < warp) wbase += x; total += x; }
        int run = wbase + incl - sum;
        #pragma unroll
        for(int j = 0; j < 4; j++) { const int s = tid * 4 + j; if(s < nslots) s_pre[s] = run; run += v[j]; }
        if(tid == 0)
        {
            const int E = total & 0xFFFFF, S = total >> 20;
            int* cur = p.cursors + 4 * view;
            const int eb = atomicAdd(cur, E), db = atomicAdd(cur + 1, S);
            s_ebase = eb; s_dbase = db;
            s_ok = (eb + E <= p.cap_e && db + S <= p.cap_d) ? 1 : 0;
            if(!s_ok) atomicOr(p.flags + view, FLAG_OVERFLOW);
        }
        __syncthreads();
        if(s_ok)
        {
            uint32_t* entries = p.entries + (size_t) view * p.cap_e + s_ebase;
            /* place */
            #pragma unroll
            for(int k = 0; k < BIN_TPT; k++)
            {
                if(!(valid >> k & 1)) continue;
                const int tx0 = rect[k] & 255, ty0 = (rect[k] >> 8) & 255, tx1 = (rect[k] >> 16) & 255, ty1 = (rect[k] >> 24) & 255;
                if((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > HUGE_TILES) continue;
                const uint32_t ent = (uint32_t) (chunk * BIN_CHUNK + k * BIN_THREADS + tid);
                for(int tx = tx0; tx <= tx1; tx++)
                    for(int ty = ty0; ty <= ty1; ty++)
                    {
                        const int s = (tx - rx0) * RH + (ty - ry0);
                        entries[(s_pre[s] & 0xFFFFF) + atomicAdd(&s_fil[s], 1)] = ent;
                    }
            }
            /* publish one segment per touched tile */
            for(int s = tid; s < nslots; s += BIN_THREADS)
            {
                const int cn = s_cnt[s];
                if(cn > 0)
                {
                    const int tile = (rx0 + s / RH) * p.tiles_y + (ry0 + s % RH);
                    publish_segment(p, view, tile, chunk, s_dbase + (s_pre[s] >> 20), s_ebase + (s_pre[s] & 0xFFFFF), cn);
                }
            }
        }
    }

    /* triangles that are huge, or all of them when the CTA's footprint does not fit the local grouping:
     * one single-entry segment per (triangle, tile) */
    #pragma unroll
    for(int k = 0; k < BIN_TPT; k++)
    {
        if(!(valid >> k & 1)) continue;
        const int tx0 = rect[k] & 255, ty0 = (rect[k] >> 8) & 255, tx1 = (rect[k] >> 16) & 255, ty1 = (rect[k] >> 24) & 255;
        if(local && (tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= HUGE_TILES) continue;
        const uint32_t ent = (uint32_t) (chunk * BIN_CHUNK + k * BIN_THREADS + tid);
        int* cur = p.cursors + 4 * view;
        for(int tx = tx0; tx <= tx1; tx++)
            for(int ty = ty0; ty <= ty1; ty++)
            {
                const int e = atomicAdd(cur, 1), id = atomicAdd(cur + 1, 1);
                if(e < p.cap_e && id < p.cap_d)
                {
                    p.entries[(size_t) view * p.cap_e + e] = ent;
                    publish_segment(p, view, tx * p.tiles_y + ty, chunk, id, e, 1);
                }
                else atomicOr(p.flags + view, FLAG_OVERFLOW);
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* K3: tile rasteriser                                                                              */
/* ------------------------------------------------------------------------------------------------ */

struct RasterParams
{
    const float4* vrec;   /* [view][ntri][VREC_QUADS] from K2 */
    const uint32_t* entries; const uint4* descs; const int* heads; const int* cursors; const uint32_t* lit_list; const int* tile_lit;
    const uint32_t* vstat; uint4* far_scratch;
    const uint32_t* tex; int tw, th;
    uint32_t* pixel; float* zbuf; unsigned long long* hash; uint32_t* flags; int* work_counter;   /* [0] lit-tile queue, [1] reset queue, [2] lit tiles published by K2 */
    int ntri, nuniq, xres, yres, tiles_x, tiles_y, ntiles, cap_e, cap_d, nviews;
    int tma_reset;        /* 1: untouched tiles are reset by TMA tensor stores through the two maps below */
    alignas(64) CUtensorMap tm_pixel, tm_z;   /* the batch's frames as 3-D tensors (y, x, view) of 32-bit words, box 32 x RESET_BOX_COLS x 1 */
};

/* per-warp scratch of the small-triangle path */
struct WarpScratch
{
    float4 slab[4][32];                 /* 2 KB   per-triangle constants of the warp's current 32 entries            */
    unsigned short unit[UNIT_WINDOW];   /* 1 KB   column unit -> (lane << 5 | x_local)                               */
    float2 q_n[QCAP];                   /* survivors of the cheap tests, waiting for the division stage: (nv, nw)     */
    uint32_t q_id[QCAP];                /*                                   triangle slot << 10 | x_local << 5 | y_local */
    uint32_t bbox[32];                  /* clipped tile-local bbox: x0 | x1 << 5 | y0 << 10 | y1 << 15 | guard << 20  */
    float den_hi[32];                   /* den * (1 + 1e-5): above it nv + nw means u < 0 for certain                 */
};

struct RasterSmem
{
    alignas(128) uint32_t pat_pixel[TH * RESET_BOX_COLS];   /* 2 KB of 0x00000000: source of the TMA stores that reset pixels */
    alignas(128) uint32_t pat_z[TH * RESET_BOX_COLS];       /* 2 KB of 0xFF7FFFFF (-FLT_MAX): the same for z                   */
    unsigned long long keys[TW * TH];   /* 8 KB  depth+winner per pixel, index key_slot(x_local, y_local) */
    WarpScratch ws[RASTER_WARPS];
    int seg_first[SEG_SLOTS];
    int seg_pre[SEG_SLOTS];             /* exclusive prefix of the staged segment sizes */
    float4 dslab[4][DEFER_MAX];         /* setup records (q0..q3 of the slab layout) of the triangles left to the CTA-wide sweep */
    uint32_t dbbox[DEFER_MAX];
    int chain[NCHAIN];
    int warp_sums[RASTER_WARPS];
    unsigned long long hash[2];
    uint32_t hiz[16];                   /* per 8x8 block: min over its pixels of the depth key (after the near phase) */
    int it_view, it_tile, it_tx, it_ty, it_clear;
    int next_entry, ndefer, nfar;
    float zthr;
};

/* slab layout (den-sign normalised: if den < 0 the four Gram terms are negated, which negates both
 * numerators and the denominator exactly, so the quotients are unchanged):
 *   q0 = ax, ay, v0x, v0y      q1 = v1x, v1y, k0, k1      q2 = d00, d01, d11, den (> 0)      q3 = az, bz, cz, ~tri */
struct TriRecord { float4 q0, q1, q2, q3; uint32_t bbox; int npx; };

/* triangle record of the view (from K2) with its bbox clipped to the tile; r4 = quad 4 of the record (already loaded by the
 * caller for the near / far decision) */
__device__ __forceinline__ TriRecord load_record(const float4* __restrict__ rec, const float4& r4, int px0, int py0, int px1, int py1)
{
    const uint32_t bx = __float_as_uint(r4.x), by = __float_as_uint(r4.y), fl = __float_as_uint(r4.z);
    const int bx0 = max((int) (bx & 0xFFFF), px0) - px0, bx1 = min((int) (bx >> 16), px1) - px0;
    const int by0 = max((int) (by & 0xFFFF), py0) - py0, by1 = min((int) (by >> 16), py1) - py0;
    TriRecord r;
    r.npx = (bx0 <= bx1 && by0 <= by1 && (fl & 1u)) ? (bx1 - bx0 + 1) * (by1 - by0 + 1) : 0;
    r.q0 = __ldg(rec); r.q1 = __ldg(rec + 1); r.q2 = __ldg(rec + 2); r.q3 = __ldg(rec + 3);
    r.bbox = (uint32_t) (bx0 & 31) | (uint32_t) (bx1 & 31) << 5 | (uint32_t) (by0 & 31) << 10 | (uint32_t) (by1 & 31) << 15 | ((fl & 2u) ? 1u << 20 : 0u);
    return r;
}

/* Two-phase depth culling (exact): a tile's triangles are split at zthr = midpoint of the range of their max
 * vertex z (from K2).  The near ones are rasterised first; the far ones are parked as 16-byte records
 * (triangle, tile-local bbox, depth bound) in the CTA's scratch.  Then hiz[] = per 8x8 block the minimum depth
 * key over its pixels, and a far triangle whose every fragment is provably below that -- z <= zmax*(1+7ulp)
 * < bound, and bound < the minimum of every block its bbox touches, strictly -- can never pass main.c:356
 * and is dropped without being set up. */
__device__ __forceinline__ uint32_t clipped_bbox(const float4& r4, int px0, int py0, int px1, int py1, bool& any)
{
    const uint32_t bx = __float_as_uint(r4.x), by = __float_as_uint(r4.y), fl = __float_as_uint(r4.z);
    const int bx0 = max((int) (bx & 0xFFFF), px0) - px0, bx1 = min((int) (bx >> 16), px1) - px0;
    const int by0 = max((int) (by & 0xFFFF), py0) - py0, by1 = min((int) (by >> 16), py1) - py0;
    any = bx0 <= bx1 && by0 <= by1 && (fl & 1u);           /* a triangle that cannot draw (den == 0 / NaN, empty bbox) is not worth parking */
    return (uint32_t) (bx0 & 31) | (uint32_t) (bx1 & 31) << 5 | (uint32_t) (by0 & 31) << 10 | (uint32_t) (by1 & 31) << 15;
}

__device__ __forceinline__ uint32_t depth_bound_key(float zmax) { return gel::zkey(zmax + fabsf(zmax) * 1e-5f + 1e-37f); }

/* true when the parked triangle cannot be culled */
__device__ __forceinline__ bool survives_hiz(const RasterSmem& sm, uint32_t bbox, uint32_t bound)
{
    const int gx0 = (bbox & 31) >> 3, gx1 = ((bbox >> 5) & 31) >> 3, gy0 = ((bbox >> 10) & 31) >> 3, gy1 = ((bbox >> 15) & 31) >> 3;
    /* every 8x8 block the bbox touches inside this tile (at most 4 x 4): large pieces are the expensive ones to
     * rasterise, so they are worth up to 16 shared-memory reads */
    uint32_t lowest = 0xFFFFFFFFu;
    for(int gx = gx0; gx <= gx1; gx++)
        for(int gy = gy0; gy <= gy1; gy++) lowest = min(lowest, sm.hiz[gx * 4 + gy]);
    return !(bound < lowest);
}

/* Cheap tests that are EXACT rejections of main.c:352 (den > 0 after normalisation):
 *   nv < -1e-20 (guard: den <= 1e18)  =>  v = nv/den is a negative non-zero float  =>  `v >= 0` fails; same for w
 *   nv + nw > den*(1+1e-5)            =>  v + w > 1 + 9e-6 after rounding           =>  u = (1-v)-w < 0
 * Everything else (including every NaN) goes on to the real divisions. */
__device__ __forceinline__ bool may_be_inside(float nv, float nw, float eps, float den_hi)
{
    return !(nv < eps) && !(nw < eps) && !(gel::add(nv, nw) > den_hi);
}

/* division, inside test and depth of main.c:327-329, 352, 355; returns the key, or 0 when outside */
__device__ __forceinline__ unsigned long long fragment_key(float nv, float nw, float den, const float4& q3)
{
    const float v = gel::dvd(nv, den), w = gel::dvd(nw, den);
    const float u = gel::sub(gel::sub(1.0f, v), w);
    if(!(v >= 0.0f && w >= 0.0f && u >= 0.0f)) return 0ull;
    const float z = gel::add(gel::add(gel::mul(v, q3.y), gel::mul(w, q3.z)), gel::mul(u, q3.x));
    if(!(z == z)) return 0ull;                          /* a NaN depth never passes `z > zbuff` (main.c:356); its key would be the largest */
    return ((unsigned long long) gel::zkey(z) << 32) | __float_as_uint(q3.w);
}

/* Where pixel (x_local, y_local) of the tile keeps its key.  Column x owns the 32 slots [x*TH, x*TH + 32) -- a warp walking a
 * column (shade pass, hi-Z) touches every bank once -- but inside the column the rows are ROTATED by 8*(x&1) + (x>>1).  A 64-bit
 * slot covers a PAIR of banks, so there are 16 bank pairs and a warp's 32 keys need two wavefronts at best.  The survivors a warp
 * resolves together mostly sit on the same row of neighbouring columns (the unit path walks the columns of a triangle in lock
 * step): with the plain x*TH + y layout they all fall on one bank pair (an 11-way conflict for an 11-column triangle; 32 % of
 * the kernel's shared-memory wavefronts were conflicts).  With the rotation up to 16 neighbouring columns of one row land on 16
 * different pairs (32 columns: each pair exactly twice), and the 4 columns x 8 rows of a sweep patch on every pair exactly twice. */
#ifndef GEL_KEY_SWIZZLE
#define GEL_KEY_SWIZZLE 2
#endif
__device__ __forceinline__ int key_slot(int xl, int yl)
{
    return GEL_KEY_SWIZZLE == 2 ? xl * TH + ((yl + 8 * (xl & 1) + (xl >> 1)) & 31)
         : GEL_KEY_SWIZZLE == 1 ? xl * TH + ((yl + 8 * (xl & 3) + (xl >> 2)) & 31) : xl * TH + yl;
}
__device__ __forceinline__ int key_slot_id(uint32_t id) { return key_slot((int) ((id >> 5) & 31), (int) (id & 31)); }

/* stage 2 of the small path: one survivor per lane */
__device__ __forceinline__ void resolve_survivor(RasterSmem& sm, WarpScratch& ws, int i)
{
    const uint32_t id = ws.q_id[i];
    const float2 n = ws.q_n[i];
    const int src = id >> 10;
    const unsigned long long key = fragment_key(n.x, n.y, ws.slab[2][src].w, ws.slab[3][src]);
    unsigned long long* k = sm.keys + key_slot_id(id);
    if(key > *reinterpret_cast<volatile unsigned long long*>(k)) atomicMax(k, key);
}

/* reset (main.c:413-417) of a tile no triangle touches: pure HBM stores, one warp per tile.  The rasteriser CTAs
 * issue these fire-and-forget stores between their work items, so they overlap the instruction-bound raster work. */
/* stage 2 of the CTA-wide sweep: the survivor's triangle record is one of the deferred records */
__device__ __forceinline__ void resolve_swept(RasterSmem& sm, WarpScratch& ws, int i)
{
    const uint32_t id = ws.q_id[i];
    const float2 n = ws.q_n[i];
    const int src = id >> 10;
    const unsigned long long key = fragment_key(n.x, n.y, sm.dslab[2][src].w, sm.dslab[3][src]);
    unsigned long long* k = sm.keys + key_slot_id(id);
    if(key > *reinterpret_cast<volatile unsigned long long*>(k)) atomicMax(k, key);
}

/* tile g of the batch (view-major), when no triangle touches it: geometry for the reset */
__device__ __forceinline__ bool untouched_tile(const RasterParams& p, int g, int& view, int& px0, int& py0, int& px1, int& py1)
{
    if(g >= p.ntiles * p.nviews || __ldg(p.tile_lit + g)) return false;
    view = g / p.ntiles;
    const int tile = g - view * p.ntiles;
    const int tx = tile / p.tiles_y, ty = tile - tx * p.tiles_y;
    px0 = tx * TW; py0 = ty * TH;
    px1 = min(px0 + TW, p.xres) - 1; py1 = min(py0 + TH, p.yres) - 1;
    return true;
}

template<bool HASH>
__device__ __forceinline__ void reset_untouched_tile(const RasterParams& p, int g, int lane, uint32_t pat_pixel, uint32_t pat_z)
{
    int view, px0, py0, px1, py1;
    if(!untouched_tile(p, g, view, px0, py0, px1, py1)) return;
    uint32_t* pixel = p.pixel + (size_t) view * p.xres * p.yres;
    float* zbuf = p.zbuf + (size_t) view * p.xres * p.yres;
    if(!HASH && (p.yres & 3) == 0 && py1 - py0 + 1 == TH)
    {
        /* 8 lanes x 16 B = one 128-byte column of the tile; 4 columns per store instruction */
        const int y = py0 + (lane & 7) * 4;
        #pragma unroll
        for(int k = 0; k < TW / 4; k++)
        {
            const int x = px0 + k * 4 + (lane >> 3);
            if(x <= px1)
            {
                const size_t idx = (size_t) y + (size_t) x * p.yres;
                *reinterpret_cast<uint4*>(pixel + idx) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<float4*>(zbuf + idx) = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            }
        }
        return;
    }
    unsigned long long hp = 0, hz = 0;
    const int y = py0 + lane;
    if(y <= py1)
        for(int x = px0; x <= px1; x++)
        {
            const int idx = y + x * p.yres;
            pixel[idx] = 0u;
            zbuf[idx] = -FLT_MAX;
            if(HASH) { hp += gel::salt_mix(0u, (uint32_t) idx); hz += gel::salt_mix(0xFF7FFFFFu, (uint32_t) idx); }
        }
    if(HASH)
    {
        for(int d = 16; d; d >>= 1) { hp += __shfl_xor_sync(0xFFFFFFFFu, hp, d); hz += __shfl_xor_sync(0xFFFFFFFFu, hz, d); }
        if(lane == 0) { atomicAdd(p.hash + 2 * view, hp); atomicAdd(p.hash + 2 * view + 1, hz); }
    }
}

/* The `count` tiles first, first + step, ... : reset those no triangle touches.
 * TMA form (cp.async.bulk.tensor, SASS UTMASTG): LANE j checks tile j -- flag load and the two divisions of the tile's position
 * run once for all the tiles -- and issues the tile's 2 x (TW / RESET_BOX_COLS) tensor stores itself: boxes of 32 rows x 16 columns
 * (pixels, then z) from two constant 2 KB patterns in shared memory; the copy unit generates the addresses and clips boxes at the
 * frame's edges.  A tensor store takes its operands from uniform registers, so ptxas serialises the lanes that issue one (~15
 * instructions per store): 4 stores per tile cost the rasteriser ~60 instructions.  (Round 2 used 16 boxes of 4 columns issued by 16
 * lanes "in one instruction": the capture showed 240 instructions per tile, 5 % of the kernel.)
 * Otherwise (checksums wanted, no tensor map): the warp-wide store loops, tile after tile. */
template<bool HASH>
__device__ __forceinline__ void reset_untouched_tiles(const RasterParams& p, int first, int step, int count, int lane, uint32_t pat_pixel, uint32_t pat_z)
{
    if(!HASH && p.tma_reset)
    {
        int view, px0, py0, px1, py1;
        if(lane < count && untouched_tile(p, first + lane * step, view, px0, py0, px1, py1))
        {
            const unsigned long long tmp = reinterpret_cast<unsigned long long>(&p.tm_pixel), tmz = reinterpret_cast<unsigned long long>(&p.tm_z);
            #pragma unroll
            for(int b = 0; b < TW / RESET_BOX_COLS; b++)
            {
                const int col = px0 + b * RESET_BOX_COLS;
                if(col > px1) break;
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                             :: "l"(tmp), "r"(py0), "r"(col), "r"(view), "r"(pat_pixel) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                             :: "l"(tmz), "r"(py0), "r"(col), "r"(view), "r"(pat_z) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");     /* bounds the groups a thread has in flight */
        }
        return;
    }
    for(int j = 0; j < count; j++) reset_untouched_tile<HASH>(p, first + j * step, lane, pat_pixel, pat_z);
}

template<bool HASH>
__global__ void __launch_bounds__(RASTER_THREADS, GEL_RASTER_MINB)
raster_kernel(const __grid_constant__ RasterParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RasterSmem& sm = *reinterpret_cast<RasterSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    WarpScratch& ws = sm.ws[warp];
    /* the reset patterns: written once through the generic proxy, read by the TMA unit (async proxy) from then on -- the fence
     * orders the two; the first barrier of the work loop publishes them to the whole CTA */
    for(int i = tid; i < TH * RESET_BOX_COLS; i += RASTER_THREADS) { sm.pat_pixel[i] = 0u; sm.pat_z[i] = 0xFF7FFFFFu; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t pat_pixel = (uint32_t) __cvta_generic_to_shared(sm.pat_pixel), pat_z = (uint32_t) __cvta_generic_to_shared(sm.pat_z);

    /* work list = the lit tiles of the whole batch as K2 published them (view << 24 | tile) */
    const int nitems = __ldg(p.work_counter + 2);

    /* thread 0 owns the work queue: the atomic for the NEXT item is issued when the current tile starts and its
     * result is only consumed after the tile's visibility pass, so the round trip hides behind real work */
    int nx_view = -1, nx_tile = 0, g_next = 0, c_next = 0;
    const int nclear = p.ntiles * p.nviews;
    auto locate = [&](int g) {
        nx_view = -1;
        if(g < nitems)
        {
            const uint32_t item = __ldg(p.lit_list + g);
            nx_view = (int) (item >> 24); nx_tile = (int) (item & 0xFFFFFFu);
        }
    };
    if(tid == 0) { locate(atomicAdd(p.work_counter, 1)); c_next = atomicAdd(p.work_counter + 1, CLEAR_CHUNK); }

    for(;;)
    {
        if(tid == 0)
        {
            sm.it_view = nx_view; sm.it_tile = nx_tile; sm.it_clear = c_next;
            const int tx = nx_tile / p.tiles_y;
            sm.it_tx = tx; sm.it_ty = nx_tile - tx * p.tiles_y;
            sm.hash[0] = 0; sm.hash[1] = 0;
        }
        __syncthreads();
        const int view = sm.it_view;
        if(view < 0) break;
        if(tid == 0) { g_next = atomicAdd(p.work_counter, 1); c_next = atomicAdd(p.work_counter + 1, CLEAR_CHUNK); }
        reset_untouched_tiles<HASH>(p, sm.it_clear + warp, RASTER_WARPS, (CLEAR_CHUNK - warp + RASTER_WARPS - 1) / RASTER_WARPS, lane, pat_pixel, pat_z);
        const int tile = sm.it_tile;
        const int px0 = sm.it_tx * TW, py0 = sm.it_ty * TH;
        const int px1 = min(px0 + TW, p.xres) - 1, py1 = min(py0 + TH, p.yres) - 1;
        uint32_t* pixel = p.pixel + (size_t) view * p.xres * p.yres;
        float* zbuf = p.zbuf + (size_t) view * p.xres * p.yres;
        const float4* __restrict__ vrec = p.vrec + (size_t) view * p.ntri * VREC_QUADS;
        const uint4* descs = p.descs + (size_t) view * p.cap_d;
        const uint32_t* entries = p.entries + (size_t) view * p.cap_e;
        unsigned long long hp = 0, hz = 0;
        const float twm1 = gel::i2f(p.tw - 1), thm1 = gel::i2f(p.th - 1);   /* (float) (w - 1), (float) (h - 1) of main.c:360-361 */

        if(tid == 0)
        {
            const float lo = gel::zkey_inv(__ldg(p.vstat + VIEW_STAT_WORDS * view)), hi = gel::zkey_inv(__ldg(p.vstat + VIEW_STAT_WORDS * view + 1));
            sm.zthr = lo + GEL_ZSPLIT_TILE * (hi - lo);
            sm.nfar = 0;
        }
        if(tid < 16) sm.hiz[tid] = 0xFFFFFFFFu;
        for(int i = tid; i < TW * TH; i += RASTER_THREADS) sm.keys[i] = CLEAR_KEY;

        /* ================= visibility: every (triangle, pixel) of main.c:348-356 inside this tile ================= */
        uint4* far_rec = p.far_scratch + (size_t) blockIdx.x * FAR_CAP;

        /* One warp, 32 candidate triangles (have / tri / a,b,c per lane) -> column units -> rows -> survivors -> keys.
         * Triangles too large for the unit path are left in sm.defer for the CTA-wide sweep. */
        int qn = 0;                                                       /* survivors on the warp's stack (warp-uniform) */
        auto rasterise_batch = [&](bool have, uint32_t tri, const float4& r4)
        {
            int nun = 0, x = 0;
            if(have)
            {
                const TriRecord r = load_record(vrec + (size_t) tri * VREC_QUADS, r4, px0, py0, px1, py1);
                bool unitised = r.npx > 0;
                if(r.npx > FRAG_MAX)
                {
                    /* too large for the unit path: its record waits in shared memory for the CTA-wide sweep (no second
                     * gather, no second setup); when the list is full the unit path takes it after all */
                    const int slot = atomicAdd(&sm.ndefer, 1);
                    if(slot < DEFER_MAX)
                    {
                        sm.dslab[0][slot] = r.q0; sm.dslab[1][slot] = r.q1; sm.dslab[2][slot] = r.q2; sm.dslab[3][slot] = r.q3;
                        sm.dbbox[slot] = r.bbox;
                        unitised = false;
                    }
                }
                if(unitised)
                {
                    x = r.bbox & 31;
                    nun = (int) ((r.bbox >> 5) & 31) - x + 1;                 /* one unit per bbox column */
                    ws.slab[0][lane] = r.q0; ws.slab[1][lane] = r.q1; ws.slab[2][lane] = r.q2; ws.slab[3][lane] = r.q3;
                    ws.bbox[lane] = r.bbox;
                    ws.den_hi[lane] = r.q2.w * U_SLACK;
                }
            }
            int uincl = nun;
            for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, uincl, d); if(lane >= d) uincl += n; }
            const int ustart = uincl - nun;
            const int utotal = __shfl_sync(0xFFFFFFFFu, uincl, 31);
            int emitted = 0;
            __syncwarp();
            for(int w0 = 0; w0 < utotal; w0 += UNIT_WINDOW)
            {
                /* expansion: one 16-bit word per column unit */
                while(emitted < nun && ustart + emitted < w0 + UNIT_WINDOW)
                {
                    ws.unit[ustart + emitted - w0] = (unsigned short) (lane << 5 | (x + emitted));
                    emitted++;
                }
                __syncwarp();
                const int n = min(UNIT_WINDOW, utotal - w0);
                for(int u0 = 0; u0 < n; u0 += 32)
                {
                    /* stage 1: numerators of v and w (main.c:325-328) down the column; exact cheap rejections */
                    const bool act = u0 + lane < n;
                    const uint32_t o = act ? ws.unit[u0 + lane] : 0u;
                    const int src = o >> 5, xl = o & 31;
                    const float4 q0 = ws.slab[0][src], q1 = ws.slab[1][src], q2 = ws.slab[2][src];
                    const uint32_t bb = ws.bbox[src];
                    const float den_hi = ws.den_hi[src];
                    const int y0l = (bb >> 10) & 31;
                    const int rows = act ? (int) ((bb >> 15) & 31) - y0l + 1 : 0;
                    const int maxrows = __reduce_max_sync(0xFFFFFFFFu, rows);
                    const float eps = (bb >> 20) & 1 ? -GUARD_EPS : -INFINITY;
                    const float v2x = gel::sub(gel::i2f(px0 + xl), q0.x);
                    const float cx0 = gel::mul(v2x, q0.z), cx1 = gel::mul(v2x, q1.x);
                    float fy = gel::i2f(py0 + y0l);
                    uint32_t id = (uint32_t) src << 10 | (uint32_t) xl << 5 | (uint32_t) y0l;
                    for(int r = 0; r < maxrows; r++, id++)
                    {
                        const float v2y = gel::sub(fy, q0.y);
                        fy = gel::add(fy, 1.0f);                              /* exact: small integers */
                        const float d20 = gel::add(gel::add(cx0, gel::mul(v2y, q0.w)), q1.z);
                        const float d21 = gel::add(gel::add(cx1, gel::mul(v2y, q1.y)), q1.w);
                        const float nv = gel::sub(gel::mul(q2.z, d20), gel::mul(q2.y, d21));
                        const float nw = gel::sub(gel::mul(q2.x, d21), gel::mul(q2.y, d20));
                        const bool pass = r < rows && may_be_inside(nv, nw, eps, den_hi);
                        const unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
                        if(pass)
                        {
                            const int slot = qn + __popc(m & lt_mask);
                            ws.q_id[slot] = id;
                            ws.q_n[slot] = make_float2(nv, nw);
                        }
                        qn += __popc(m);
                        if(qn >= 32)
                        {
                            /* stage 2: divisions, inside test, depth, key -- a full warp off the top of the stack */
                            __syncwarp();
                            qn -= 32;
                            resolve_survivor(sm, ws, qn + lane);
                            __syncwarp();
                        }
                    }
                }
                __syncwarp();
            }
            __syncwarp();
            if(lane < qn) resolve_survivor(sm, ws, lane);
            qn = 0;
            __syncwarp();
        };

        /* large triangles: the whole CTA sweeps one triangle at a time in 4x8-pixel patches dealt to the warps;
         * survivors of the cheap tests are compacted so the divisions run in full warps.  Called after a barrier that
         * makes the deferred records visible; ends with one. */
        auto sweep_deferred = [&]()
        {
            const int ndefer = min(sm.ndefer, DEFER_MAX);
            int sq = 0;                                                   /* survivors on this warp's stack */
            for(int li = 0; li < ndefer; li++)
            {
                const uint32_t bb = sm.dbbox[li];
                const int gx0 = bb & 31, gx1 = (bb >> 5) & 31, gy0 = (bb >> 10) & 31, gy1 = (bb >> 15) & 31;
                const float eps = (bb >> 20) & 1 ? -GUARD_EPS : -INFINITY;
                const float4 q0 = sm.dslab[0][li], q1 = sm.dslab[1][li], q2 = sm.dslab[2][li];
                const float den_hi = q2.w * U_SLACK;
                /* the bbox is covered by patches of 4 columns x 8 rows (lane = 8*column + row): clipped bboxes are
                 * rarely 32 rows tall, so this keeps far more lanes busy than one 32-row column per warp */
                const int pcols = (gx1 - gx0 + 4) >> 2, prows = (gy1 - gy0 + 8) >> 3;
                for(int pr = 0; pr < prows; pr++)
                {
                    const int yl = gy0 + pr * 8 + (lane & 7);
                    const bool rowok = yl <= gy1;
                    const float v2y = gel::sub(gel::i2f(py0 + yl), q0.y);
                    const float cy0 = gel::mul(v2y, q0.w), cy1 = gel::mul(v2y, q1.y);
                    for(int pc = warp; pc < pcols; pc += RASTER_WARPS)
                    {
                        const int xl = gx0 + pc * 4 + (lane >> 3);
                        const float v2x = gel::sub(gel::i2f(px0 + xl), q0.x);
                        const float d20 = gel::add(gel::add(gel::mul(v2x, q0.z), cy0), q1.z);
                        const float d21 = gel::add(gel::add(gel::mul(v2x, q1.x), cy1), q1.w);
                        const float nv = gel::sub(gel::mul(q2.z, d20), gel::mul(q2.y, d21));
                        const float nw = gel::sub(gel::mul(q2.x, d21), gel::mul(q2.y, d20));
                        const bool pass = rowok && xl <= gx1 && may_be_inside(nv, nw, eps, den_hi);
                        const unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
                        if(pass)
                        {
                            const int slot = sq + __popc(m & lt_mask);
                            ws.q_id[slot] = (uint32_t) li << 10 | (uint32_t) xl << 5 | (uint32_t) yl;
                            ws.q_n[slot] = make_float2(nv, nw);
                        }
                        sq += __popc(m);
                        if(sq >= 32)
                        {
                            /* divisions, inside test, depth, key for a full warp of survivors */
                            __syncwarp();
                            sq -= 32;
                            resolve_swept(sm, ws, sq + lane);
                            __syncwarp();
                        }
                    }
                }
            }
            __syncwarp();
            if(lane < sq) resolve_swept(sm, ws, lane);
            __syncthreads();
        };

        /* ---------------- phase 0: walk the tile's segments; near triangles are rasterised, far ones parked ---------------- */
        if(tid < NCHAIN) sm.chain[tid] = __ldg(p.heads + ((size_t) view * p.ntiles + tile) * NCHAIN + tid);
        bool first_round = true;
        float zthr = 0.0f;
        for(;;)
        {
            /* stage up to SEG_SLOTS segments: chain c fills slots [c*32, c*32+32) */
            if(tid == 0) { sm.next_entry = 0; sm.ndefer = 0; }
            __syncthreads();
            if(tid < NCHAIN)
            {
                int cur = sm.chain[tid], k = 0;
                while(cur >= 0 && k < SEG_SLOTS / NCHAIN)
                {
                    const uint4 d = __ldg(descs + cur);
                    sm.seg_first[tid * (SEG_SLOTS / NCHAIN) + k] = (int) d.y;
                    sm.seg_pre[tid * (SEG_SLOTS / NCHAIN) + k] = (int) d.z;      /* size for now */
                    cur = (int) d.x; k++;
                }
                for(; k < SEG_SLOTS / NCHAIN; k++) sm.seg_pre[tid * (SEG_SLOTS / NCHAIN) + k] = 0;
                sm.chain[tid] = cur;
            }
            const int more = __syncthreads_or(tid < NCHAIN && sm.chain[tid] >= 0);
            const int my_count = sm.seg_pre[tid];
            int incl = my_count;
            for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if(lane >= d) incl += n; }
            if(lane == 31) sm.warp_sums[warp] = incl;
            __syncthreads();
            int wbase = 0, round_entries = 0;
            #pragma unroll
            for(int w = 0; w < RASTER_WARPS; w++) { const int x = sm.warp_sums[w]; if(w < warp) wbase += x; round_entries += x; }
            sm.seg_pre[tid] = wbase + incl - my_count;
            if(first_round)
            {
                /* short lists are not worth a second phase: everything is "near" */
                zthr = (!more && round_entries < TWO_PHASE_MIN) ? -INFINITY : sm.zthr;
                first_round = false;
            }
            __syncthreads();

            /* every warp pulls `grab` entries at a time: 32 for long lists; fewer for short ones so that all the
             * warps of the CTA get a share (large triangles are split into column units afterwards anyway) */
            const int grab = max(4, min(32, (round_entries + 2 * RASTER_WARPS - 1) / (2 * RASTER_WARPS)));
            for(;;)
            {
                int e0 = 0;
                if(lane == 0) e0 = atomicAdd(&sm.next_entry, grab);
                e0 = __shfl_sync(0xFFFFFFFFu, e0, 0);
                if(e0 >= round_entries) break;
                {
                    const int e = e0 + lane;
                    bool have = false, park = false;
                    uint32_t tri = 0, bbox = 0, bound = 0;
                    float4 r4 = make_float4(0, 0, 0, 0);
                    if(lane < grab && e < round_entries)
                    {
                        /* staged segment holding entry e: last slot with seg_pre <= e */
                        int lo = 0;
                        #pragma unroll
                        for(int step = SEG_SLOTS / 2; step; step >>= 1) if(sm.seg_pre[lo + step] <= e) lo += step;
                        tri = __ldg(entries + sm.seg_first[lo] + (e - sm.seg_pre[lo]));
                        r4 = __ldg(vrec + (size_t) tri * VREC_QUADS + 4);
                        const float zmax = r4.w;
                        have = true;
                        if(zmax < zthr)                                   /* NaN compares false: near */
                        {
                            bool any;
                            bbox = clipped_bbox(r4, px0, py0, px1, py1, any);
                            bound = depth_bound_key(zmax);
                            park = any; have = false;
                        }
                    }
                    const unsigned pm = __ballot_sync(0xFFFFFFFFu, park);
                    if(pm)
                    {
                        int base = 0;
                        if(lane == 0) base = atomicAdd(&sm.nfar, __popc(pm));
                        base = __shfl_sync(0xFFFFFFFFu, base, 0);
                        if(park)
                        {
                            const int slot = base + __popc(pm & lt_mask);
                            if(slot < FAR_CAP) far_rec[slot] = make_uint4(tri, bbox, bound, 0u);
                            else have = true;                             /* scratch full: rasterise it now */
                        }
                    }
                    rasterise_batch(have, tri, r4);
                }
            }
            __syncthreads();
            sweep_deferred();
            if(!more) break;
        }
        __syncthreads();

        /* ---------------- phase 1: the parked triangles against the hierarchical depth of what is already drawn ---------------- */
        const int nfar = min(sm.nfar, FAR_CAP);
        if(nfar > 0)
        {
            /* hiz: a warp reads column x (lane = row); block = (x/8)*4 + lane/8 */
            /* block column gx = the 8 pixel columns [8 gx, 8 gx + 8): a lane first takes the minimum of its row over them, then
             * one reduction per group of 8 lanes (8 rows) gives the block's value -- no atomics, each block has one writer */
            for(int gx = warp; gx < TW / 8; gx += RASTER_WARPS)
            {
                uint32_t zk = 0xFFFFFFFFu;
                #pragma unroll
                for(int dx = 0; dx < 8; dx++) zk = min(zk, (uint32_t) (sm.keys[key_slot(gx * 8 + dx, lane)] >> 32));
                zk = __reduce_min_sync(0xFFu << (lane & 24), zk);
                if((lane & 7) == 0) sm.hiz[gx * 4 + (lane >> 3)] = zk;
            }
            if(tid == 0) { sm.next_entry = 0; sm.ndefer = 0; }
            __syncthreads();                                              /* hiz complete, far_rec visible to the whole CTA */
            for(;;)
            {
                int e0 = 0;
                if(lane == 0) e0 = atomicAdd(&sm.next_entry, 32);
                e0 = __shfl_sync(0xFFFFFFFFu, e0, 0);
                if(e0 >= nfar) break;
                {
                    const int e = e0 + lane;
                    bool have = false;
                    uint32_t tri = 0;
                    float4 r4 = make_float4(0, 0, 0, 0);
                    if(e < nfar)
                    {
                        const uint4 rec = far_rec[e];
                        if(survives_hiz(sm, rec.y, rec.z))
                        {
                            tri = rec.x;
                            r4 = __ldg(vrec + (size_t) tri * VREC_QUADS + 4);
                            have = true;
                        }
                    }
                    if(__any_sync(0xFFFFFFFFu, have)) rasterise_batch(have, tri, r4);
                }
            }
            __syncthreads();
            sweep_deferred();
            __syncthreads();
        }
        if(tid == 0) locate(g_next);

        /* ================= shade the winner of every pixel once (main.c:358-366), write the tile back ================= */
        for(int i = tid; i < TW * TH; i += RASTER_THREADS)
        {
            const int x = px0 + (i >> 5), y = py0 + (i & 31);
            if(x > px1 || y > py1) continue;
            const unsigned long long key = sm.keys[key_slot(i >> 5, i & 31)];
            uint32_t colour = 0u;
            float z = -FLT_MAX;
            if(key != CLEAR_KEY)
            {
                const uint32_t tri = 0xFFFFFFFFu - (uint32_t) key;
                z = gel::zkey_inv((uint32_t) (key >> 32));
                /* tbarycenter at this pixel (main.c:316-332) from the triangle's record: the operations and operands of the
                 * visibility pass, so v, w, u are the bits that passed the inside test there */
                const float4* __restrict__ rec = vrec + (size_t) tri * VREC_QUADS;
                const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), sh = __ldg(rec + 5), t0 = __ldg(rec + 6), t1 = __ldg(rec + 7);
                const float v2x = gel::sub(gel::i2f(x), q0.x), v2y = gel::sub(gel::i2f(y), q0.y);
                const float d20 = gel::add(gel::add(gel::mul(v2x, q0.z), gel::mul(v2y, q0.w)), q1.z);
                const float d21 = gel::add(gel::add(gel::mul(v2x, q1.x), gel::mul(v2y, q1.y)), q1.w);
                const float nv = gel::sub(gel::mul(q2.z, d20), gel::mul(q2.y, d21));
                const float nw = gel::sub(gel::mul(q2.x, d21), gel::mul(q2.y, d20));
                const float v = gel::dvd(nv, q2.w), w = gel::dvd(nw, q2.w);
                const float u = gel::sub(gel::sub(1.0f, v), w);
                const float uv[6] = { t0.x, t0.y, t0.z, t0.w, t1.x, t1.y };
                int xx, yy, shading;
                gel::fragment_shade_f(v, w, u, uv, sh.x, sh.y, sh.z, twm1, thm1, xx, yy, shading);
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
        if(HASH)
        {
            for(int d = 16; d; d >>= 1) { hp += __shfl_xor_sync(0xFFFFFFFFu, hp, d); hz += __shfl_xor_sync(0xFFFFFFFFu, hz, d); }
            if(lane == 0) { atomicAdd(&sm.hash[0], hp); atomicAdd(&sm.hash[1], hz); }
            __syncthreads();
            if(tid == 0) { atomicAdd(p.hash + 2 * view, sm.hash[0]); atomicAdd(p.hash + 2 * view + 1, sm.hash[1]); }
        }
        __syncthreads();
    }
    /* no lit tile left: finish the chunk already reserved, then drain the reset queue */
    for(int base = sm.it_clear; base < nclear; )
    {
        reset_untouched_tiles<HASH>(p, base + warp, RASTER_WARPS, (CLEAR_CHUNK - warp + RASTER_WARPS - 1) / RASTER_WARPS, lane, pat_pixel, pat_z);
        __syncthreads();
        if(tid == 0) sm.it_clear = atomicAdd(p.work_counter + 1, CLEAR_CHUNK);
        __syncthreads();
        base = sm.it_clear;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");               /* every tensor store this thread issued has landed */
}

} /* namespace gelk */
#endif /* GEL_KERNELS_CUH */
