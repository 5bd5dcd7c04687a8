/* gel_direct.cuh -- the DIRECT pipeline: gel's render path for meshes whose triangles cover a few pixels each
 * (cfg 3: 1 M triangles of ~3 px^2 at 4K).  For such meshes screen-tile binning costs more than the raster work
 * itself, so warps stream the triangles in submission order and merge fragments straight into a per-view key
 * buffer in global memory with 64-bit atomic max (resolved in L2), using the same key as the tile pipeline:
 *     key = zkey(z) << 32 | (0xFFFFFFFF - triangle)         (main.c:356 semantics, see gel_kernels.cuh)
 *
 *   D0 direct_clear_kernel    publishes the view's screen region (bbox from K1) and depth threshold, hi-Z := 0
 *                             (the key buffer is "no winner" everywhere between batches: allocated that way, and the
 *                             resolve pass puts every key it consumed back to that value)
 *   D1 direct_raster_kernel<0>  every triangle: near ones (max vertex z >= view midpoint) are rasterised, far ones parked
 *   D2 direct_hiz_kernel      per 8x8 pixel block: minimum depth key over its pixels
 *   D3 direct_raster_kernel<1>  parked triangles: dropped when provably behind every block they touch, else rasterised
 *   D5 direct_resolve_kernel  every pixel of the frame: winner shaded once (main.c:358-366) or reset (main.c:413-417),
 *                             colour + z written once, coalesced
 *
 * The per-pixel arithmetic (numerators, exact cheap rejections, divisions, depth, shading) is shared with the tile
 * pipeline; only the staging differs.  D1/D3 run one warp per CTA (32 resident per SM): a warp that parks most of
 * its triangles retires at once instead of idling beside slower warps of the same CTA.
 */
#ifndef GEL_DIRECT_CUH
#define GEL_DIRECT_CUH

#include "gel_kernels.cuh"
#include <type_traits>

namespace gelk {

#ifndef GEL_DIRECT_THREADS
#define GEL_DIRECT_THREADS 32
#endif
#ifndef GEL_DIRECT_MINB
#define GEL_DIRECT_MINB (1024 / GEL_DIRECT_THREADS)     /* resident CTAs per SM the raster kernels are compiled for */
#endif
constexpr int DIRECT_THREADS = GEL_DIRECT_THREADS;
constexpr int DIRECT_WARPS = DIRECT_THREADS / 32;
#ifndef GEL_HIZ_WIDE
#define GEL_HIZ_WIDE 1
#endif
#ifndef GEL_ZSPLIT
#define GEL_ZSPLIT 0.4f           /* near / far split of a view, as a fraction of its depth range (any value is exact; this one is a speed heuristic) */
#endif
#ifndef GEL_RESOLVE_WCOLS
#define GEL_RESOLVE_WCOLS 4        /* columns of a resolve warp's pixel footprint (x 32/WCOLS rows); 1, 2, 4 or 8 */
#endif
#ifndef GEL_RESOLVE_MINB
#define GEL_RESOLVE_MINB 8        /* resident CTAs per SM the resolve pass is compiled for (32 registers): it is latency bound */
#endif
#ifndef GEL_DIRECT_TPW
#define GEL_DIRECT_TPW 1024
#endif
#ifndef GEL_TRIM_ROUNDS
#define GEL_TRIM_ROUNDS 1          /* rounds of exact bbox trimming per rasterised triangle (0 = off) */
#endif
constexpr int RESOLVE_WCOLS = GEL_RESOLVE_WCOLS;
constexpr int TREC_QUADS = 4;                    /* wide static record per triangle for the resolve pass: 64 bytes */
constexpr int TREC_COMPACT_BITS = 21;            /* meshes with < 2^21 distinct vertices: 32-byte record, three 21-bit indices */
constexpr int DIRECT_TRIS_PER_WARP = GEL_DIRECT_TPW;          /* consecutive triangles a warp streams through (large batches) */
constexpr int DIRECT_TRIS_PER_WARP_MIN = 128;                  /* ... down to this many when a batch would otherwise leave SMs without warps */
constexpr int REGION_WORDS = 8;                    /* per view: x0, x1, y0, y1 (block aligned, -1.. when empty), zthr bits */
constexpr int DIRECT_UNIT_WINDOW = 256;
constexpr int DIRECT_CAND = 256;                 /* ring of hi-Z survivors waiting for a full batch (phase 1) */
constexpr int DIRECT_HIZ_SPAN = 4;               /* parked bboxes of up to this many 8x8 blocks per side are tested against the hi-Z map */
constexpr int DIRECT_TEST_UNROLL = 4;            /* parked records tested per lane per refill round */
#ifndef GEL_DIRECT_MAX_ROWS
#define GEL_DIRECT_MAX_ROWS 32
#endif
constexpr int DIRECT_MAX_ROWS = GEL_DIRECT_MAX_ROWS;                /* taller (or > FRAG_MAX px) bboxes are swept by the whole warp */
constexpr int VSTAT = VIEW_STAT_WORDS;             /* per-view words: zlo, zhi, xmin, xmax, ymin, ymax, far count, - */

struct DirectParams
{
    const float4* xf; const uint32_t *i0, *i1, *i2; const uint4* trec;   /* per triangle, wide: 4 x 16 bytes {i0,i1,i2,-} {u0,v0,u1,v1} {u2,v2,-,-} -;  compact: 2 x 16 bytes {i0 | i1 << 21 | i2 << 42 (64 bits), u0, v0} {u1,v1,u2,v2} */
    const uint32_t* tex; int tw, th;
    unsigned long long* keys;      /* [view][xres*yres]  index y + x*yres                                   */
    uint32_t* hiz;                 /* [view][hbx*hby]    min depth key per 8x8 block, index bx*hby + by     */
    uint4* far;                    /* [view][ntri]       parked: tri, x0 | x1 << 16, y0 | y1 << 16, bound; warp w of D1 owns
                                    *                     records [w*tpw, ...) and writes their count to far_count */
    int* far_count;                /* [view][warps]                                                         */
    int* region;                   /* [view][REGION_WORDS]  written by D0                                   */
    uint32_t* vstat;               /* [view][VSTAT]                                                         */
    uint32_t* pixel; float* zbuf; unsigned long long* hash; uint32_t* flags;
    int ntri, nuniq, xres, yres, hbx, hby, nviews;
    int tpw;                       /* triangles per warp of D1 / D3 for this batch (power of two) */
};

/* per warp.  The parked pass (PHASE 1) also keeps a ring of hi-Z survivors; the near pass has no use for it, and without it a one-warp
 * CTA needs 4 224 + 1 024 bytes of shared memory: 32 CTAs per SM fit the 164 KB carve-out exactly (the common layout needed the 196 KB
 * one; a carve-out beyond the kernel's need costs the step 2 %: profiles/README.md, session 3). */
struct DirectCand { uint32_t cand[DIRECT_CAND]; };   /* triangle ids */
struct DirectNoCand {};
template<int PHASE>
struct DirectScratchT : std::conditional<PHASE == 1, DirectCand, DirectNoCand>::type
{
    float4 slab[4][32];                  /* q0 .. q3 of the current 32 triangles (gel_kernels.cuh, slab layout) */
    uint32_t unit[DIRECT_UNIT_WINDOW];   /* lane << 13 | x */
    float2 q_n[QCAP];
    uint32_t q_id[QCAP];                 /* lane << 26 | x << 13 | y */
    uint32_t bx[32], by[32];             /* x0 | x1 << 16 ;  y0 | y1 << 13 | guard << 26 */
    float den_hi[32];
};
static_assert(sizeof(DirectScratchT<0>) == 4224 && sizeof(DirectScratchT<1>) == 5248, "shared-memory budget of the direct raster kernels (carve-out fits)");

/* L2 cache-policy hints (sm_80+ createpolicy; on sm_100a the policy rides in the memory descriptor, no extra
 * instruction per access).  The fill's 51 MB per frame of write-once stores are marked evict-first and the key buffer's
 * reductions evict-last, so that the background fill can run UNDER the instruction-bound near pass without pushing the
 * L2-resident keys out to DRAM (profiles/README.md, round 2). */
__device__ __forceinline__ uint64_t l2_policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void st_v4_hint(void* ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d), "l"(pol) : "memory");
}
template<bool HINT> __device__ __forceinline__ void st_b32(void* ptr, uint32_t v, uint64_t pol)
{
    if(HINT) asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(ptr), "r"(v), "l"(pol) : "memory");
    else *reinterpret_cast<uint32_t*>(ptr) = v;
}
template<bool HINT> __device__ __forceinline__ void st_b64(void* ptr, unsigned long long v, uint64_t pol)
{
    if(HINT) asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" :: "l"(ptr), "l"(v), "l"(pol) : "memory");
    else *reinterpret_cast<unsigned long long*>(ptr) = v;
}
template<bool HINT>
__device__ __forceinline__ void key_max(unsigned long long* ptr, unsigned long long key, uint64_t pol)
{
    if(HINT) asm volatile("red.global.max.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(ptr), "l"(key), "l"(pol) : "memory");
    else atomicMax(ptr, key);
}

/* screen bbox of a view's vertices, clipped to the frame and widened to whole 8x8 blocks; false when empty */
__device__ __forceinline__ bool view_region(const DirectParams& p, int view, int& x0, int& x1, int& y0, int& y1)
{
    return region_from_stats(p.vstat + (size_t) view * VSTAT, p.xres, p.yres, x0, x1, y0, y1);
}

/* the region D0 published for the view; false when empty */
__device__ __forceinline__ bool load_region(const DirectParams& p, int view, int& x0, int& x1, int& y0, int& y1)
{
    const int4 r = __ldg(reinterpret_cast<const int4*>(p.region + (size_t) view * REGION_WORDS));
    x0 = r.x; x1 = r.y; y0 = r.z; y1 = r.w;
    return x0 <= x1 && y0 <= y1;
}

/* D0 ------------------------------------------------------------------------------------------------------------ */
/* grid (G, nviews): CTA g of a view takes the region's columns x0+g, x0+g+G, ... */
__global__ void __launch_bounds__(256)
direct_clear_kernel(DirectParams p)
{
    const int view = blockIdx.y;
    int x0, x1, y0, y1;
    /* hi-Z defaults to 0 (= nothing can be culled) everywhere */
    for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.hbx * p.hby; i += gridDim.x * blockDim.x)
        p.hiz[(size_t) view * p.hbx * p.hby + i] = 0u;
    const bool any = view_region(p, view, x0, x1, y0, y1);
    if(blockIdx.x == 0 && threadIdx.x == 0)
    {
        int* r = p.region + (size_t) view * REGION_WORDS;
        const uint32_t* s = p.vstat + (size_t) view * VSTAT;
        const float lo = gel::zkey_inv(s[0]), hi = gel::zkey_inv(s[1]);
        r[0] = any ? x0 : 0; r[1] = any ? x1 : -1; r[2] = any ? y0 : 0; r[3] = any ? y1 : -1;
        r[4] = __float_as_int(lo + GEL_ZSPLIT * (hi - lo));          /* near / far split of the view */
    }
}

/* whole key buffer := "no winner" (after allocation, or after a call that did not reach its resolve pass) */
__global__ void __launch_bounds__(256)
direct_keys_init_kernel(unsigned long long* keys, size_t n)
{
    for(size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) keys[i] = CLEAR_KEY;
}

/* D2 ------------------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256)
direct_hiz_kernel(DirectParams p)
{
    const int view = blockIdx.y;
    int x0, x1, y0, y1;
    if(!load_region(p, view, x0, x1, y0, y1)) return;
    /* a warp takes 8 columns x 32 rows (4 blocks): lanes along y so every load is one contiguous 256-byte run */
    const int nbx = (x1 >> 3) - (x0 >> 3) + 1, nby = (y1 >> 3) - (y0 >> 3) + 1, nby4 = (nby + 3) >> 2;
    const unsigned long long* keys = p.keys + (size_t) view * p.xres * p.yres;
    const int lane = threadIdx.x & 31, nwarps = gridDim.x * (blockDim.x >> 5);
    for(int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); task < nbx * nby4; task += nwarps)
    {
        const int bx = (x0 >> 3) + task / nby4, by = (y0 >> 3) + (task % nby4) * 4 + (lane >> 3);
        const int y = by * 8 + (lane & 7);
        const bool live = y < p.yres && by <= (y1 >> 3);
        uint32_t lowest = 0xFFFFFFFFu;
        if(live)
        {
            const unsigned long long* col = keys + (size_t) y + (size_t) bx * 8 * p.yres;
#pragma unroll
            for(int dx = 0; dx < 8; dx++)
                if(bx * 8 + dx < p.xres) lowest = min(lowest, (uint32_t) (col[(size_t) dx * p.yres] >> 32));
        }
        for(int d = 1; d < 8; d <<= 1) lowest = min(lowest, __shfl_xor_sync(0xFFFFFFFFu, lowest, d));
        if(live && (lane & 7) == 0) p.hiz[(size_t) view * p.hbx * p.hby + (size_t) bx * p.hby + by] = lowest;
    }
}

/* D1 / D3 ------------------------------------------------------------------------------------------------------- */

template<bool HINT, class Scratch>
__device__ __forceinline__ void direct_resolve(const DirectParams& p, unsigned long long* keys, Scratch& ws, int i, uint64_t pol)
{
    const uint32_t id = ws.q_id[i];
    const float2 n = ws.q_n[i];
    const uint32_t src = id >> 26, x = (id >> 13) & 8191u, y = id & 8191u;
    const unsigned long long key = fragment_key(n.x, n.y, ws.slab[2][src].w, ws.slab[3][src]);
    if(key) key_max<HINT>(keys + (x * (uint32_t) p.yres + y), key, pol);  /* 32-bit unsigned offset inside the view's frame */
}

template<int PHASE, bool HINT>
__global__ void __launch_bounds__(DIRECT_THREADS, GEL_DIRECT_MINB)
direct_raster_kernel(DirectParams p)
{
    __shared__ DirectScratchT<PHASE> scratch[DIRECT_WARPS];
    const uint64_t pol = HINT ? l2_policy_evict_last() : 0ull;
    const int view = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    DirectScratchT<PHASE>& ws = scratch[warp];
    const float4* xf = p.xf + (size_t) view * p.nuniq;
    unsigned long long* keys = p.keys + (size_t) view * p.xres * p.yres;
    uint4* far = p.far + (size_t) view * p.ntri;
    const uint32_t* hiz = p.hiz + (size_t) view * p.hbx * p.hby;
    const float zthr = PHASE == 0 ? __int_as_float(__ldg(p.region + (size_t) view * REGION_WORDS + 4)) : 0.0f;
    const int gwarp = blockIdx.x * DIRECT_WARPS + warp, nwarps = gridDim.x * DIRECT_WARPS;
    const int first = gwarp * p.tpw;
    if(first >= p.ntri) return;
    int* my_far_count = p.far_count + (size_t) view * nwarps + gwarp;
    const int last = PHASE == 0 ? min(first + p.tpw, p.ntri) : first + *my_far_count;
    int qn = 0, parked = 0;
    bool clipped = false;

    int tnext = first, chead = 0, ncand = 0;
    for(;;)
    {
        bool have = false, park = false;
        uint32_t tri = 0, pbx = 0, pby = 0, bound = 0;
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
        if constexpr(PHASE == 0)
        {
            if(tnext >= last) break;
            const int t = tnext + lane;
            tnext += 32;
            if(t < last)
            {
                tri = (uint32_t) t;
                a = __ldg(xf + __ldg(p.i0 + tri)); b = __ldg(xf + __ldg(p.i1 + tri)); c = __ldg(xf + __ldg(p.i2 + tri));
                const float zmax = fmaxf(a.z, fmaxf(b.z, c.z));
                have = true;
                if(zmax < zthr)                                           /* NaN compares false: near */
                {
                    int x0 = gel::trunc_i(fminf(a.x, fminf(b.x, c.x))), x1 = gel::trunc_i(fmaxf(a.x, fmaxf(b.x, c.x)));
                    int y0 = gel::trunc_i(fminf(a.y, fminf(b.y, c.y))), y1 = gel::trunc_i(fmaxf(a.y, fmaxf(b.y, c.y)));
                    if(x0 < 0 || y0 < 0 || x1 > p.xres - 1 || y1 > p.yres - 1) { clipped = true; x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, p.xres - 1); y1 = min(y1, p.yres - 1); }
                    park = x0 <= x1 && y0 <= y1; have = false;
                    pbx = (uint32_t) x0 | (uint32_t) x1 << 16; pby = (uint32_t) y0 | (uint32_t) y1 << 16;
                    bound = depth_bound_key(zmax);
                }
            }
        }
        else
        {
            /* the parked records are tested against the hi-Z map several per lane (independent loads in flight);
             * the few survivors wait in a ring until a full batch of 32 can be rasterised */
            while(ncand < 32 && tnext < last)
            {
                uint4 rec[DIRECT_TEST_UNROLL];
                bool survive[DIRECT_TEST_UNROLL];
#pragma unroll
                for(int k = 0; k < DIRECT_TEST_UNROLL; k++)
                {
                    const int t = tnext + k * 32 + lane;
                    rec[k] = t < last ? far[t] : make_uint4(0u, 0u, 0u, 0u);
                    survive[k] = t < last;
                }
#pragma unroll
                for(int k = 0; k < DIRECT_TEST_UNROLL; k++)
                {
                    const int gx0 = (rec[k].y & 0xFFFF) >> 3, gx1 = (rec[k].y >> 16) >> 3, gy0 = (rec[k].z & 0xFFFF) >> 3, gy1 = (rec[k].z >> 16) >> 3;
                    if(survive[k] && gx1 - gx0 <= 1 && gy1 - gy0 <= 1)
                    {
                        const uint32_t lowest = min(min(__ldg(hiz + gx0 * p.hby + gy0), __ldg(hiz + gx0 * p.hby + gy1)),
                                                    min(__ldg(hiz + gx1 * p.hby + gy0), __ldg(hiz + gx1 * p.hby + gy1)));
                        survive[k] = !(rec[k].w < lowest);
                    }
                    else if(GEL_HIZ_WIDE && survive[k] && gx1 - gx0 < DIRECT_HIZ_SPAN && gy1 - gy0 < DIRECT_HIZ_SPAN)
                    {
                        /* larger bbox: every block it touches (the bigger the triangle, the more a cull saves) */
                        uint32_t lowest = 0xFFFFFFFFu;
                        for(int gx = gx0; gx <= gx1; gx++)
                            for(int gy = gy0; gy <= gy1; gy++) lowest = min(lowest, __ldg(hiz + gx * p.hby + gy));
                        survive[k] = !(rec[k].w < lowest);
                    }
                }
#pragma unroll
                for(int k = 0; k < DIRECT_TEST_UNROLL; k++)
                {
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, survive[k]);
                    if(survive[k]) ws.cand[(chead + ncand + __popc(m & lt_mask)) & (DIRECT_CAND - 1)] = rec[k].x;
                    ncand += __popc(m);
                }
                tnext += 32 * DIRECT_TEST_UNROLL;
            }
            if(ncand == 0) break;
            __syncwarp();
            const int take = min(ncand, 32);
            if(lane < take)
            {
                tri = ws.cand[(chead + lane) & (DIRECT_CAND - 1)];
                a = __ldg(xf + __ldg(p.i0 + tri)); b = __ldg(xf + __ldg(p.i1 + tri)); c = __ldg(xf + __ldg(p.i2 + tri));
                have = true;
            }
            chead += take; ncand -= take;
            __syncwarp();
        }
        if(PHASE == 0)
        {
            /* parked triangles are compacted into this warp's own slice of far[] (ballot ranks, no atomics) */
            const unsigned pm = __ballot_sync(0xFFFFFFFFu, park);
            if(park) far[first + parked + __popc(pm & lt_mask)] = make_uint4(tri, pbx, pby, bound);
            parked += __popc(pm);
        }
        if(!__any_sync(0xFFFFFFFFu, have)) continue;

        /* ---- per-triangle setup (main.c:319-324, 344-347), bbox clipped to the frame ---- */
        int nun = 0, ux = 0;
        bool sweep = false;
        if(have)
        {
            const gel::TriSetup s = gel::tri_setup(a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z);
            int x0 = s.x0, y0 = s.y0, x1 = s.x1, y1 = s.y1;
            if(x0 < 0 || y0 < 0 || x1 > p.xres - 1 || y1 > p.yres - 1) { clipped = true; x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, p.xres - 1); y1 = min(y1, p.yres - 1); }
            const float ad = fabsf(s.den);
            const bool drawable = ad > 0.0f && x0 <= x1 && y0 <= y1;     /* den == 0 or NaN can never pass main.c:352 */
            if(drawable)
            {
                const bool guard = ad <= GUARD_DEN_MAX;
                const float sg = s.den < 0.0f ? -1.0f : 1.0f;
                const float4 q2 = make_float4(s.d00 * sg, s.d01 * sg, s.d11 * sg, s.den * sg);
                /* edges of the bbox on which no pixel can pass main.c:352 are dropped before any pixel is tested (exact:
                 * gel_math.h, bbox_trim); a tiny triangle's first column and first row almost always go */
                #pragma unroll
                for(int round = 0; round < GEL_TRIM_ROUNDS; round++)
                    gel::bbox_trim(s.ax, s.ay, s.v0x, s.v0y, s.v1x, s.v1y, s.k0, s.k1, q2.x, q2.y, q2.z, q2.w, x0, y0, x1, y1);
                if(x0 <= x1 && y0 <= y1)
                {
                    ws.slab[0][lane] = make_float4(s.ax, s.ay, s.v0x, s.v0y);
                    ws.slab[1][lane] = make_float4(s.v1x, s.v1y, s.k0, s.k1);
                    ws.slab[2][lane] = q2;
                    ws.slab[3][lane] = make_float4(s.az, s.bz, s.cz, __uint_as_float(0xFFFFFFFFu - tri));
                    ws.bx[lane] = (uint32_t) x0 | (uint32_t) x1 << 16;
                    ws.by[lane] = (uint32_t) y0 | (uint32_t) y1 << 13 | (guard ? 1u << 26 : 0u);
                    ws.den_hi[lane] = s.den * sg * U_SLACK;
                    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
                    sweep = bh > DIRECT_MAX_ROWS || bw * bh > FRAG_MAX;
                    if(!sweep) { nun = bw; ux = x0; }
                }
            }
        }
        __syncwarp();

        /* ---- column units: one lane per bbox column, rows walked with a warp-uniform trip count ---- */
        int uincl = nun;
        for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, uincl, d); if(lane >= d) uincl += n; }
        const int ustart = uincl - nun;
        const int utotal = __shfl_sync(0xFFFFFFFFu, uincl, 31);
        int emitted = 0;
        for(int w0 = 0; w0 < utotal; w0 += DIRECT_UNIT_WINDOW)
        {
            while(emitted < nun && ustart + emitted < w0 + DIRECT_UNIT_WINDOW)
            {
                ws.unit[ustart + emitted - w0] = (uint32_t) lane << 13 | (uint32_t) (ux + emitted);
                emitted++;
            }
            __syncwarp();
            const int n = min(DIRECT_UNIT_WINDOW, utotal - w0);
            for(int u0 = 0; u0 < n; u0 += 32)
            {
                const bool act = u0 + lane < n;
                const uint32_t o = act ? ws.unit[u0 + lane] : 0u;
                const int src = o >> 13, x = o & 8191;
                const float4 q0 = ws.slab[0][src], q1 = ws.slab[1][src], q2 = ws.slab[2][src];
                const uint32_t yy = ws.by[src];
                const float den_hi = ws.den_hi[src];
                const int y0 = yy & 8191;
                const int rows = act ? (int) ((yy >> 13) & 8191) - y0 + 1 : 0;
                const int maxrows = __reduce_max_sync(0xFFFFFFFFu, rows);
                const float eps = (yy >> 26) & 1 ? -GUARD_EPS : -INFINITY;
                const float v2x = gel::sub(gel::i2f(x), q0.x);
                const float cx0 = gel::mul(v2x, q0.z), cx1 = gel::mul(v2x, q1.x);
                float fy = gel::i2f(y0);
                uint32_t id = (uint32_t) src << 26 | (uint32_t) x << 13 | (uint32_t) y0;
                for(int r = 0; r < maxrows; r++, id++)
                {
                    const float v2y = gel::sub(fy, q0.y);
                    fy = gel::add(fy, 1.0f);
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
                        __syncwarp();
                        qn -= 32;
                        direct_resolve<HINT>(p, keys, ws, qn + lane, pol);
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
        if(lane < qn) direct_resolve<HINT>(p, keys, ws, lane, pol);
        qn = 0;
        __syncwarp();

        /* ---- triangles too large for the unit path: the whole warp sweeps the bbox, lanes along y ---- */
        unsigned sm = __ballot_sync(0xFFFFFFFFu, sweep);
        while(sm)
        {
            const int src = __ffs(sm) - 1;
            sm &= sm - 1;
            const float4 q0 = ws.slab[0][src], q1 = ws.slab[1][src], q2 = ws.slab[2][src], q3 = ws.slab[3][src];
            const uint32_t xx = ws.bx[src], yy = ws.by[src];
            const int x0 = xx & 0xFFFF, x1 = xx >> 16, y0 = yy & 8191, y1 = (yy >> 13) & 8191;
            const float eps = (yy >> 26) & 1 ? -GUARD_EPS : -INFINITY;
            const float den_hi = ws.den_hi[src];
            for(int yb = y0; yb <= y1; yb += 32)
            {
                const int y = yb + lane;
                if(y > y1) continue;
                const float v2y = gel::sub(gel::i2f(y), q0.y);
                const float cy0 = gel::mul(v2y, q0.w), cy1 = gel::mul(v2y, q1.y);
                for(int x = x0; x <= x1; x++)
                {
                    const float v2x = gel::sub(gel::i2f(x), q0.x);
                    const float d20 = gel::add(gel::add(gel::mul(v2x, q0.z), cy0), q1.z);
                    const float d21 = gel::add(gel::add(gel::mul(v2x, q1.x), cy1), q1.w);
                    const float nv = gel::sub(gel::mul(q2.z, d20), gel::mul(q2.y, d21));
                    const float nw = gel::sub(gel::mul(q2.x, d21), gel::mul(q2.y, d20));
                    if(!may_be_inside(nv, nw, eps, den_hi)) continue;
                    const unsigned long long key = fragment_key(nv, nw, q2.w, q3);
                    if(key) key_max<HINT>(keys + ((uint32_t) x * (uint32_t) p.yres + (uint32_t) y), key, pol);
                }
            }
        }
        __syncwarp();
    }
    if(PHASE == 0 && lane == 0) *my_far_count = parked;
    if(__any_sync(0xFFFFFFFFu, clipped) && lane == 0) atomicOr(p.flags + view, FLAG_CLIPPED);
}

/* D5 ------------------------------------------------------------------------------------------------------------ */
/* one pixel: the winner's barycentrics recomputed from the same operands (identical bits), shaded once */
template<bool COMPACT>
__device__ __forceinline__ void direct_shade(const DirectParams& p, const float4* __restrict__ xf, uint32_t vbase /* first vertex of the view in xf */, uint32_t* vflags, float twm1, float thm1,
                                             unsigned long long key, int x, int y, uint32_t& colour, float& z)
{
    colour = 0u; z = -FLT_MAX;
    if(key == CLEAR_KEY) return;
    const uint32_t tri = 0xFFFFFFFFu - (uint32_t) key;
    z = gel::zkey_inv((uint32_t) (key >> 32));
    uint32_t ia, ib, ic, uvw[6];
    if(COMPACT)
    {
        float4 f0, f1;
        ldg256(p.trec + 2 * (size_t) tri, f0, f1);                       /* the 32-byte record in one request */
        const uint4 r0 = make_uint4(__float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f0.z), __float_as_uint(f0.w));
        const uint4 r1 = make_uint4(__float_as_uint(f1.x), __float_as_uint(f1.y), __float_as_uint(f1.z), __float_as_uint(f1.w));
        const uint32_t m = (1u << TREC_COMPACT_BITS) - 1u;
        ia = r0.x & m; ib = (r0.x >> 21 | r0.y << 11) & m; ic = (r0.y >> 10) & m;
        uvw[0] = r0.z; uvw[1] = r0.w; uvw[2] = r1.x; uvw[3] = r1.y; uvw[4] = r1.z; uvw[5] = r1.w;
    }
    else
    {
        const uint4* rec = p.trec + (size_t) TREC_QUADS * tri;
        float4 f0, f1;
        ldg256(rec, f0, f1);
        const uint4 r0 = make_uint4(__float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f0.z), __float_as_uint(f0.w));
        const uint4 r1 = make_uint4(__float_as_uint(f1.x), __float_as_uint(f1.y), __float_as_uint(f1.z), __float_as_uint(f1.w));
        const uint4 r2 = __ldg(rec + 2);
        ia = r0.x; ib = r0.y; ic = r0.z;
        uvw[0] = r1.x; uvw[1] = r1.y; uvw[2] = r1.z; uvw[3] = r1.w; uvw[4] = r2.x; uvw[5] = r2.y;
    }
    const float4 a = __ldg(xf + (vbase + ia));
    const float4 b = __ldg(xf + (vbase + ib));
    const float4 c = __ldg(xf + (vbase + ic));
    /* tbarycenter at this pixel (main.c:316-332), same operations and operands as the visibility pass */
    const float v0x = gel::sub(b.x, a.x), v0y = gel::sub(b.y, a.y), v0z = gel::sub(b.z, a.z);
    const float v1x = gel::sub(c.x, a.x), v1y = gel::sub(c.y, a.y), v1z = gel::sub(c.z, a.z);
    const float d00 = gel::dot3(v0x, v0y, v0z, v0x, v0y, v0z), d01 = gel::dot3(v0x, v0y, v0z, v1x, v1y, v1z), d11 = gel::dot3(v1x, v1y, v1z, v1x, v1y, v1z);
    const float den = gel::sub(gel::mul(d00, d11), gel::mul(d01, d01));
    const float v2x = gel::sub(gel::i2f(x), a.x), v2y = gel::sub(gel::i2f(y), a.y), v2z = gel::sub(0.0f, a.z);
    const float d20 = gel::add(gel::add(gel::mul(v2x, v0x), gel::mul(v2y, v0y)), gel::mul(v2z, v0z));
    const float d21 = gel::add(gel::add(gel::mul(v2x, v1x), gel::mul(v2y, v1y)), gel::mul(v2z, v1z));
    const float v = gel::dvd(gel::sub(gel::mul(d11, d20), gel::mul(d01, d21)), den);
    const float w = gel::dvd(gel::sub(gel::mul(d00, d21), gel::mul(d01, d20)), den);
    const float u = gel::sub(gel::sub(1.0f, v), w);
    const float uv[6] = { __uint_as_float(uvw[0]), __uint_as_float(uvw[1]), __uint_as_float(uvw[2]), __uint_as_float(uvw[3]), __uint_as_float(uvw[4]), __uint_as_float(uvw[5]) };
    int xx, yy, shading;
    gel::fragment_shade_f(v, w, u, uv, a.w, b.w, c.w, twm1, thm1, xx, yy, shading);
    if((unsigned) xx > (unsigned) (p.tw - 1) || (unsigned) yy > (unsigned) (p.th - 1))
    {
        atomicOr(vflags, FLAG_TEXCLAMP);                   /* the reference reads out of bounds here (R) */
        xx = min(max(xx, 0), p.tw - 1); yy = min(max(yy, 0), p.th - 1);
    }
    colour = gel::pshade(__ldg(p.tex + (uint32_t) (xx + yy * p.tw)), shading);
}

/* D5a: reset (main.c:413-417) of everything outside the view's region -- pure stores.
 * grid (ceil(yres / 1024), xres, nviews): a thread owns 4 consecutive rows of one column (regions are 8-aligned) */
template<bool HASH, bool HINT>
__global__ void __launch_bounds__(256)
direct_fill_kernel(DirectParams p)
{
    const uint64_t pol = HINT ? l2_policy_evict_first() : 0ull;
    const int view = blockIdx.z, x = blockIdx.y;
    const int y4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int rx0, rx1, ry0, ry1;
    const bool region = load_region(p, view, rx0, rx1, ry0, ry1) && x >= rx0 && x <= rx1;
    unsigned long long hp = 0, hz = 0;
    if(y4 < p.yres && !(region && y4 >= ry0 && y4 <= ry1))
    {
        const size_t base = (size_t) view * p.xres * p.yres + (size_t) x * p.yres;
        if((p.yres & 3) == 0)
        {
            if(HINT) { st_v4_hint(p.pixel + base + y4, 0u, 0u, 0u, 0u, pol); st_v4_hint(p.zbuf + base + y4, 0xFF7FFFFFu, 0xFF7FFFFFu, 0xFF7FFFFFu, 0xFF7FFFFFu, pol); }
            else
            {
                *reinterpret_cast<uint4*>(p.pixel + base + y4) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<float4*>(p.zbuf + base + y4) = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            }
        }
        for(int k = 0; k < 4; k++)
        {
            const int y = y4 + k;
            if(y >= p.yres || (region && y >= ry0 && y <= ry1)) continue;
            if((p.yres & 3) != 0) { p.pixel[base + y] = 0u; p.zbuf[base + y] = -FLT_MAX; }
            if(HASH) { const uint32_t idx = (uint32_t) (y + x * p.yres); hp += gel::salt_mix(0u, idx); hz += gel::salt_mix(0xFF7FFFFFu, idx); }
        }
    }
    if(HASH)
    {
        for(int d = 16; d; d >>= 1) { hp += __shfl_xor_sync(0xFFFFFFFFu, hp, d); hz += __shfl_xor_sync(0xFFFFFFFFu, hz, d); }
        if((threadIdx.x & 31) == 0 && (hp | hz)) { atomicAdd(p.hash + 2 * view, hp); atomicAdd(p.hash + 2 * view + 1, hz); }
    }
}

/* D5a, overlapped form: the same reset as direct_fill_kernel, as a small persistent grid (a few CTAs per SM) that is
 * launched BEFORE the near pass on a higher-priority stream and streams its stores out underneath it.  Work item = one
 * column of one view, view-major like the near pass; 256 threads x 4 rows per store pair.  HINT marks the stores
 * evict-first in L2. */
template<bool HASH, bool HINT>
__global__ void __launch_bounds__(256)
direct_fill_persistent_kernel(DirectParams p, int sleep_ns)
{
    const uint64_t pol = HINT ? l2_policy_evict_first() : 0ull;
    const uint32_t zc = 0xFF7FFFFFu;                                   /* -FLT_MAX */
    unsigned long long hp = 0, hz = 0;
    int hview = -1;
    const int nitems = p.nviews * p.xres;
    auto flush_hash = [&]() {
        if constexpr(HASH)
        {
            if(hview < 0) return;
            for(int d = 16; d; d >>= 1) { hp += __shfl_xor_sync(0xFFFFFFFFu, hp, d); hz += __shfl_xor_sync(0xFFFFFFFFu, hz, d); }
            if((threadIdx.x & 31) == 0 && (hp | hz)) { atomicAdd(p.hash + 2 * hview, hp); atomicAdd(p.hash + 2 * hview + 1, hz); }
            hp = 0; hz = 0;
        }
    };
    for(int item = blockIdx.x; item < nitems; item += gridDim.x)
    {
        const int view = item / p.xres, x = item - view * p.xres;
        if(HASH && view != hview) { flush_hash(); hview = view; }
        int rx0, rx1, ry0, ry1;
        const bool region = load_region(p, view, rx0, rx1, ry0, ry1) && x >= rx0 && x <= rx1;
        const size_t base = (size_t) view * p.xres * p.yres + (size_t) x * p.yres;
        for(int y4 = threadIdx.x * 4; y4 < p.yres; y4 += 1024)
        {
            if(region && y4 >= ry0 && y4 <= ry1) continue;               /* regions are 8-aligned: a group of 4 rows is in or out as a whole */
            if((p.yres & 3) == 0)
            {
                if(HINT) { st_v4_hint(p.pixel + base + y4, 0u, 0u, 0u, 0u, pol); st_v4_hint(p.zbuf + base + y4, zc, zc, zc, zc, pol); }
                else
                {
                    *reinterpret_cast<uint4*>(p.pixel + base + y4) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<float4*>(p.zbuf + base + y4) = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
                }
            }
            for(int k = 0; k < 4; k++)
            {
                const int y = y4 + k;
                if(y >= p.yres || (region && y >= ry0 && y <= ry1)) continue;
                if((p.yres & 3) != 0) { p.pixel[base + y] = 0u; p.zbuf[base + y] = -FLT_MAX; }
                if(HASH) { const uint32_t idx = (uint32_t) (y + x * p.yres); hp += gel::salt_mix(0u, idx); hz += gel::salt_mix(zc, idx); }
            }
        }
        /* pacing: at full speed these stores saturate DRAM and every memory operation of the raster kernel beside them
         * queues behind them; spread over the near pass's duration they take a quarter of the bandwidth */
        if(sleep_ns > 0) __nanosleep((unsigned) sleep_ns);
    }
    flush_hash();
}

/* D5a, bulk form (sm_90+ / sm_100a): the same reset issued as TMA bulk stores -- `cp.async.bulk.global.shared::cta` (SASS
 * UBLKCP) copies of a constant pattern held in shared memory (zeros for the pixels, 0xFF7FFFFF = -FLT_MAX for z).  One elected
 * thread per CTA issues a copy per 8 KB of frame, so the 3.3 GB per cfg-3 step cost a few hundred thousand instructions instead of
 * the 245 M warp-instructions of the store loop: the fill stops competing with the instruction-bound near pass for issue slots and
 * LSU queues and can run underneath it.  Needs yres % 4 == 0 (16-byte granules); the caller falls back to the store loop otherwise,
 * and whenever checksums are requested (those are computed per reset pixel).
 *   per view: the frame as a flat word array; everything left of the region's first column and right of its last one is two
 *   contiguous ranges, cut into BULK_BYTES pieces; inside the region's columns the rows below / above the region are one strip each. */
constexpr int BULK_BYTES = 8192;
__device__ __forceinline__ void bulk_store(void* gptr, uint32_t smem_addr, uint32_t bytes, uint64_t pol, bool hint)
{
    if(hint) asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" :: "l"(gptr), "r"(smem_addr), "r"(bytes), "l"(pol) : "memory");
    else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gptr), "r"(smem_addr), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(32)
direct_fill_bulk_kernel(DirectParams p, int hint)
{
    __shared__ __align__(128) uint32_t pat_pixel[BULK_BYTES / 4];
    __shared__ __align__(128) uint32_t pat_z[BULK_BYTES / 4];
    for(int i = threadIdx.x; i < BULK_BYTES / 4; i += 32) { pat_pixel[i] = 0u; pat_z[i] = 0xFF7FFFFFu; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        /* the pattern is read by the async proxy */
    __syncwarp();
    if(threadIdx.x != 0) return;
    const uint64_t pol = hint ? l2_policy_evict_first() : 0ull;
    const uint32_t sp = (uint32_t) __cvta_generic_to_shared(pat_pixel), sz = (uint32_t) __cvta_generic_to_shared(pat_z);
    const size_t frame = (size_t) p.xres * p.yres, frame_bytes = frame * 4;
    const int pieces = (int) ((frame_bytes + BULK_BYTES - 1) / BULK_BYTES);
    const int per_view = pieces + p.xres;                               /* flat pieces, then one item per column (strips) */
    const long long nitems = (long long) p.nviews * per_view;
    for(long long item = blockIdx.x; item < nitems; item += gridDim.x)
    {
        const int view = (int) (item / per_view), k = (int) (item - (long long) view * per_view);
        int rx0, rx1, ry0, ry1;
        const bool any = load_region(p, view, rx0, rx1, ry0, ry1);
        char* pix = reinterpret_cast<char*>(p.pixel + (size_t) view * frame);
        char* zb = reinterpret_cast<char*>(p.zbuf + (size_t) view * frame);
        if(k < pieces)
        {
            /* flat piece [b0, b1) minus the region's column range [r0, r1) (bytes): at most one part on either side */
            const size_t b0 = (size_t) k * BULK_BYTES, b1 = min(b0 + (size_t) BULK_BYTES, frame_bytes);
            const size_t r0 = any ? (size_t) rx0 * p.yres * 4 : frame_bytes, r1 = any ? (size_t) (rx1 + 1) * p.yres * 4 : frame_bytes;
            const size_t a1 = min(b1, r0);                               /* part left of the region */
            if(a1 > b0) { bulk_store(pix + b0, sp, (uint32_t) (a1 - b0), pol, hint); bulk_store(zb + b0, sz, (uint32_t) (a1 - b0), pol, hint); }
            const size_t c0 = max(b0, r1);                               /* part right of it */
            if(b1 > c0) { bulk_store(pix + c0, sp, (uint32_t) (b1 - c0), pol, hint); bulk_store(zb + c0, sz, (uint32_t) (b1 - c0), pol, hint); }
        }
        else if(any)
        {
            const int x = k - pieces;
            if(x < rx0 || x > rx1) continue;
            const size_t col = (size_t) x * p.yres * 4;
            for(size_t o = 0; o < (size_t) ry0 * 4; o += BULK_BYTES)
            {
                const uint32_t n = (uint32_t) min((size_t) BULK_BYTES, (size_t) ry0 * 4 - o);
                bulk_store(pix + col + o, sp, n, pol, hint); bulk_store(zb + col + o, sz, n, pol, hint);
            }
            for(size_t o = (size_t) (ry1 + 1) * 4; o < (size_t) p.yres * 4; o += BULK_BYTES)
            {
                const uint32_t n = (uint32_t) min((size_t) BULK_BYTES, (size_t) p.yres * 4 - o);
                bulk_store(pix + col + o, sp, n, pol, hint); bulk_store(zb + col + o, sz, n, pol, hint);
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");    /* bound the copies in flight per CTA */
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");             /* every store has landed before the kernel ends */
}

/* D5b: every pixel inside the region: winner shaded once or reset.  grid (G, nviews): a CTA takes strips of 8
 * adjacent columns, 32 rows at a time; a warp covers RESOLVE_WCOLS columns x (32 / RESOLVE_WCOLS) rows of them.  The
 * compact footprint is what keeps the gathers cheap: a warp's 32 pixels then touch few distinct triangles, vertices
 * and -- above all -- texture rows (a 1 x 32 column touched ~30 texture lines per texel load, the L1 data pipe was
 * the kernel's limit), and neighbouring warps reuse the same lines out of L1. */
template<bool HASH, bool COMPACT, bool HINT>
__global__ void __launch_bounds__(256, GEL_RESOLVE_MINB)
direct_resolve_kernel(DirectParams p)
{
    const uint64_t pol = HINT ? l2_policy_evict_first() : 0ull;          /* frames are written once and not read again by this path */
    constexpr int WROWS = 32 / RESOLVE_WCOLS, WARPS_X = 8 / RESOLVE_WCOLS, WARPS_Y = 8 / WARPS_X, CTA_ROWS = WROWS * WARPS_Y;
    const int view = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = (warp % WARPS_X) * RESOLVE_WCOLS + lane / WROWS, py = (warp / WARPS_X) * WROWS + lane % WROWS;
    int rx0, rx1, ry0, ry1;
    if(!load_region(p, view, rx0, rx1, ry0, ry1)) return;
    unsigned long long hp = 0, hz = 0;
    uint32_t* vflags = p.flags + view;
    const float twm1 = gel::i2f(p.tw - 1), thm1 = gel::i2f(p.th - 1);     /* (float) (w - 1), (float) (h - 1) of main.c:360-361, converted once */
    const int nstrips = (rx1 - rx0 + 8) / 8;
    /* Addressing: a pixel is ONE 32-bit element index into the batch's buffers (view * frame + x * yres + y; the host keeps
     * views-per-batch * frame and views-per-batch * distinct vertices below 2^32), scaled onto the buffer bases.  With per-view 64-bit
     * base pointers the 32-register budget made ptxas rebuild them in every iteration (45 of the loop's 196 instructions). */
    const uint32_t frame = (uint32_t) p.xres * (uint32_t) p.yres;
    const uint32_t fbase = (uint32_t) view * frame, vbase = (uint32_t) view * (uint32_t) p.nuniq;
    unsigned long long* bkeys = p.keys; uint32_t* bpixel = p.pixel; float* bz = p.zbuf; const float4* bxf = p.xf;
    for(int strip = blockIdx.x; strip < nstrips; strip += gridDim.x)
    {
        const int x = rx0 + strip * 8 + px;
        if(x > rx1) continue;
        uint32_t off = (uint32_t) x * (uint32_t) p.yres + (uint32_t) (ry0 + py);     /* inside the frame: the checksum's position salt */
        unsigned long long next_key = ry0 + py <= ry1 ? bkeys[fbase + off] : CLEAR_KEY;
        for(int y = ry0 + py; y <= ry1; y += CTA_ROWS, off += CTA_ROWS)
        {
            const uint32_t g = fbase + off;
            const unsigned long long key = next_key;
            if(y + CTA_ROWS <= ry1) next_key = bkeys[g + CTA_ROWS];         /* one iteration ahead of its use */
            uint32_t colour; float z;
            direct_shade<COMPACT>(p, bxf, vbase, vflags, twm1, thm1, key, x, y, colour, z);
            st_b64<HINT>(bkeys + g, CLEAR_KEY, pol);                     /* the buffer is all "no winner" again for the next batch */
            st_b32<HINT>(bpixel + g, colour, pol);
            st_b32<HINT>(bz + g, __float_as_uint(z), pol);
            if(HASH) { hp += gel::salt_mix(colour, off); hz += gel::salt_mix(__float_as_uint(z), off); }
        }
    }
    if(HASH)
    {
        for(int d = 16; d; d >>= 1) { hp += __shfl_xor_sync(0xFFFFFFFFu, hp, d); hz += __shfl_xor_sync(0xFFFFFFFFu, hz, d); }
        if(lane == 0 && (hp | hz)) { atomicAdd(p.hash + 2 * view, hp); atomicAdd(p.hash + 2 * view + 1, hz); }
    }
}

} /* namespace gelk */
#endif /* GEL_DIRECT_CUH */
