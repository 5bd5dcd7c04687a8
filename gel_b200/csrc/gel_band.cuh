/* gel_band.cuh -- K3 of the tile pipeline, warp-per-band form (device code only).
 *
 * raster_kernel (gel_kernels.cuh) gives a 32x32 tile to a CTA of four warps: the warps share the tile's keys in shared memory
 * (64-bit atomicMax, which sm_100 executes as a compare-and-swap spin: ~45 shared-memory wavefronts per warp instruction), hand
 * large triangles to a CTA-wide sweep and meet at a dozen barriers per tile -- barrier waits were its largest stall.
 *
 * Here the unit of work is a BAND: 8 adjacent pixel columns x 32 rows of a lit tile (a quarter tile), and it belongs to ONE WARP
 * from its first triangle to its write-back.  The warp walks the tile's whole entry list itself -- with the per-(view, triangle)
 * records of K2 an entry costs one 16-byte load and a bbox comparison before it is known to miss the band -- and rasterises what
 * overlaps its 8 columns with the column-unit path (a unit = one bbox column, rows walked with a warp-uniform trip count, exact
 * cheap rejections, survivors compacted on a stack so the divisions run in full warps).  Consequences:
 *   - nobody else touches the band's keys: the depth merge is a plain load / compare / store, repeated while two survivors of the
 *     same flush hit one pixel (the value only grows, so the loop ends; almost always one round) -- no atomics at all;
 *   - no barrier anywhere after the CTA's start-up: warps pull (view, tile, band) items from the work queue independently;
 *   - a triangle of any size goes through the same path: inside a band a large triangle is 8 units of up to 32 rows -- full
 *     lanes, uniform trip counts -- so the CTA-wide sweep of large triangles and its deferral lists are gone;
 *   - the frame is still written as full 128-byte column segments (a band column is 32 rows = 128 bytes);
 *   - a triangle much larger than a band covers a chord of each of its bbox columns: before the row loop every unit trims its row
 *     range to the rows that can pass the cheap tests (gel_math.h: row_trim, exact), which takes a third of the row iterations away.
 * The arithmetic per (triangle, pixel) -- and therefore every bit of the frames -- is that of raster_kernel (main.c:316-370);
 * so are the exact two-phase depth culling (near triangles first, far ones parked and tested against the band's own hi-Z) and the
 * TMA reset of untouched tiles.
 */
#ifndef GEL_BAND_CUH
#define GEL_BAND_CUH

#include "gel_kernels.cuh"

namespace gelk {

#ifndef GEL_BAND_MERGE
#define GEL_BAND_MERGE 0                          /* depth merge of a flush: 0 = store + re-check (two survivors of one flush on one pixel race, the larger key stores again);
                                                   * 1 = MATCH.ANY on the slot, atomic only on such a conflict: race-free under compute-sanitizer, same frames, -4.5 % */
#endif
#ifndef GEL_BAND_ROW_TRIM
#define GEL_BAND_ROW_TRIM 1                       /* per-column exact row trimming in the unit prologue (0 = walk the bbox rows) */
#endif
constexpr int BAND_W = 8;                         /* pixel columns of a band */
constexpr int BANDS = TW / BAND_W;                /* bands per tile */
constexpr int BAND_UNITS = 32 * BAND_W;           /* most column units a batch of 32 triangles has inside a band */
constexpr int BAND_FAR_CAP = FAR_CAP / BANDS;     /* far triangles a warp can park per band (16 B each, global scratch) */
constexpr int BAND_SEGS = 32;                     /* segments staged per round: one per lane */
constexpr int BAND_CLEAR = 2;                     /* tiles a warp checks (and resets when untouched) per band item */
static_assert(TW % BAND_W == 0 && BAND_W == 8 && TH == 32, "a band is one 8x8-block column of a 32-row tile");

struct BandScratch
{
    unsigned long long keys[BAND_W * TH];         /* 2 KB  depth + winner per pixel: band_slot(x_local, y_local)                 */
    float4 slab[4][32];                           /* 2 KB  per-triangle constants of the current 32 entries                     */
    unsigned char unit[BAND_UNITS];               /* 256 B column unit -> (lane << 3 | x_local)                                 */
    float2 q_n[QCAP];                             /* survivors of the cheap tests, waiting for the division stage: (nv, nw)     */
    uint32_t q_id[QCAP];                          /*                          triangle slot << 10 | x_local << 5 | y_local      */
    uint32_t bbox[32];                            /* band-local bbox: x0 | x1 << 5 | y0 << 10 | y1 << 15 | guard << 20          */
    float2 etrim[32];                             /* slack terms of the row trimming (K2 record, quad 7: ev, ew)               */
    int seg_first[BAND_SEGS], seg_pre[BAND_SEGS]; /* staged segments: first entry, exclusive prefix of the sizes                */
    uint32_t hiz[4];                              /* per 8x8 block of the band: min depth key after the near phase             */
};

struct BandSmem
{
    alignas(128) uint32_t pat_pixel[TH * RESET_BOX_COLS];
    alignas(128) uint32_t pat_z[TH * RESET_BOX_COLS];
    BandScratch ws[RASTER_WARPS];
};

/* column x owns 32 slots; its rows are rotated by 2x so that one row of the band's 8 columns lands on 8 different bank pairs */
__device__ __forceinline__ int band_slot(int xl, int yl) { return xl * TH + ((yl + 2 * xl) & 31); }

/* Queue ticket for ONE lane: the counter's value before it is incremented by one.  atomicAdd() -- and equally a PTX `atom.add` -- on a
 * warp-uniform address is turned by ptxas into its warp-aggregated form (vote, leader atomic, SHFL of the result), and that shuffle waits
 * for the atomic's round trip on the spot: the capture showed 6 % of the kernel's stall samples on the two tickets that are meant to be
 * fetched one item AHEAD.  `atom.inc` is left alone, and its result is only waited for where it is used: at the top of the next item. */
__device__ __forceinline__ int queue_ticket(int* counter)
{
    unsigned old;
    asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(old) : "l"(counter) : "memory");
    return (int) old;
}

template<bool HASH>
__global__ void __launch_bounds__(RASTER_THREADS, GEL_RASTER_MINB)
raster_band_kernel(const __grid_constant__ RasterParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BandSmem& sm = *reinterpret_cast<BandSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    BandScratch& ws = sm.ws[warp];
    for(int i = tid; i < TH * RESET_BOX_COLS; i += RASTER_THREADS) { sm.pat_pixel[i] = 0u; sm.pat_z[i] = 0xFF7FFFFFu; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t pat_pixel = (uint32_t) __cvta_generic_to_shared(sm.pat_pixel), pat_z = (uint32_t) __cvta_generic_to_shared(sm.pat_z);
    __syncthreads();                                                      /* the only barrier: the patterns are in place */

    const int nitems = __ldg(p.work_counter + 2) * BANDS;                 /* (lit tile, band) pairs of the batch */
    const int nclear = p.ntiles * p.nviews;
    const float twm1 = gel::i2f(p.tw - 1), thm1 = gel::i2f(p.th - 1);     /* (float) (w - 1), (float) (h - 1) of main.c:360-361 */
    const size_t frame = (size_t) p.xres * p.yres;
    uint4* far_rec = p.far_scratch + ((size_t) blockIdx.x * RASTER_WARPS + warp) * BAND_FAR_CAP;

    /* lane 0 owns the warp's queue cursors: the atomics for the NEXT item are issued when the current one starts */
    int g_next = 0, c_next = 0;
    if(lane == 0) { g_next = queue_ticket(p.work_counter); c_next = queue_ticket(p.work_counter + 1) * BAND_CLEAR; }

    for(;;)
    {
        const int g = __shfl_sync(0xFFFFFFFFu, g_next, 0), c0 = __shfl_sync(0xFFFFFFFFu, c_next, 0);
        if(g >= nitems)
        {
            /* no band left: finish the chunk already reserved, then drain the reset queue */
            for(int base = c0; base < nclear; )
            {
                reset_untouched_tiles<HASH>(p, base, 1, BAND_CLEAR, lane, pat_pixel, pat_z);
                if(lane == 0) base = queue_ticket(p.work_counter + 1) * BAND_CLEAR;
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
            }
            break;
        }
        if(lane == 0) { g_next = queue_ticket(p.work_counter); c_next = queue_ticket(p.work_counter + 1) * BAND_CLEAR; }
        reset_untouched_tiles<HASH>(p, c0, 1, BAND_CLEAR, lane, pat_pixel, pat_z);

        const uint32_t item = __ldg(p.lit_list + g / BANDS);
        const int band = g % BANDS, view = (int) (item >> 24), tile = (int) (item & 0xFFFFFFu);
        const int tx = tile / p.tiles_y, ty = tile - tx * p.tiles_y;
        const int px0 = tx * TW + band * BAND_W, py0 = ty * TH;
        const int px1 = min(px0 + BAND_W, p.xres) - 1, py1 = min(py0 + TH, p.yres) - 1;
        if(px0 > px1) continue;                                           /* the band lies right of the frame */
        /* element index of the band's first pixel in the batch's frame buffers (32 bits: the host keeps views per batch x frame < 2^32) */
        const uint32_t gband = (uint32_t) view * (uint32_t) frame + (uint32_t) px0 * (uint32_t) p.yres + (uint32_t) (py0 + lane);
        const float4* __restrict__ vrec = p.vrec + (size_t) view * p.ntri * VREC_QUADS;
        const uint4* __restrict__ descs = p.descs + (size_t) view * p.cap_d;
        const uint32_t* __restrict__ entries = p.entries + (size_t) view * p.cap_e;

        #pragma unroll
        for(int i = 0; i < BAND_W * TH / 32; i++) ws.keys[i * 32 + lane] = CLEAR_KEY;
        float zthr_view;
        {
            const float lo = gel::zkey_inv(__ldg(p.vstat + VIEW_STAT_WORDS * view)), hi = gel::zkey_inv(__ldg(p.vstat + VIEW_STAT_WORDS * view + 1));
            zthr_view = lo + GEL_ZSPLIT_TILE * (hi - lo);
        }
        int nfar = 0;
        __syncwarp();

        /* ---- depth merge of one survivor per lane: division, inside test, depth (main.c:327-329, 352, 355-356) ---- */
        auto resolve = [&](int i, bool valid)
        {
            unsigned long long key = 0ull;
            unsigned long long* k = ws.keys;
            if(valid)
            {
                const uint32_t id = ws.q_id[i];
                const float2 n = ws.q_n[i];
                const int src = id >> 10;
                key = fragment_key(n.x, n.y, ws.slab[2][src].w, ws.slab[3][src]);
                k = ws.keys + band_slot((int) ((id >> 5) & 31), (int) (id & 31));
            }
#if GEL_BAND_MERGE == 1
            /* Only this warp writes these keys, so the depth merge needs no atomic: load, compare, store.  Two survivors of one flush
             * can still share a pixel (neighbouring triangles on their common edge, overdraw inside a batch): MATCH.ANY on the slot
             * finds that, and only then the warp goes through the atomic (a compare-and-swap loop on sm_100; rare). */
            const bool pending = key > *k;
            __syncwarp();                                                 /* every lane has read its slot before any lane writes one */
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, pending ? (uint32_t) (k - ws.keys) : 0x10000u + (uint32_t) lane);
            if(__any_sync(0xFFFFFFFFu, (peers & (peers - 1)) != 0u))
            { if(pending) atomicMax(k, key); }
            else if(pending) *k = key;
#else
            /* Only this warp writes these keys: load / compare / store, again while two survivors of this flush share a pixel.  That is
             * a deliberate intra-warp write-write race (compute-sanitizer's racecheck reports it): each lane's 64-bit store is one
             * aligned access, so one of the racing keys lands whole; every lane then re-reads its slot, and a lane whose key is still
             * larger stores once more -- the slot only ever grows, so the loop ends with the maximum in place (almost always after one
             * round).  GEL_BAND_MERGE=1 is the race-free spelling. */
            bool pending = key > *reinterpret_cast<volatile unsigned long long*>(k);
            while(__any_sync(0xFFFFFFFFu, pending))
            {
                if(pending) *reinterpret_cast<volatile unsigned long long*>(k) = key;
                __syncwarp();
                pending = pending && key > *reinterpret_cast<volatile unsigned long long*>(k);
                __syncwarp();
            }
#endif
            __syncwarp();                                                 /* the stack slots just read may be pushed on again */
        };

        /* ---- one batch: up to 32 triangles (have / tri / r4 per lane) -> column units -> rows -> survivors -> keys ---- */
        int qn = 0;                                                       /* survivors on the stack (warp-uniform) */
        auto rasterise_batch = [&](bool have, uint32_t tri, const float4& r4)
        {
            int nun = 0, x = 0;
            if(have)
            {
                const TriRecord r = load_record(vrec + (size_t) tri * VREC_QUADS, r4, px0, py0, px1, py1);
                if(r.npx > 0)
                {
                    x = r.bbox & 31;
                    nun = (int) ((r.bbox >> 5) & 31) - x + 1;             /* one unit per bbox column inside the band: <= 8 */
                    ws.slab[0][lane] = r.q0; ws.slab[1][lane] = r.q1; ws.slab[2][lane] = r.q2; ws.slab[3][lane] = r.q3;
                    ws.bbox[lane] = r.bbox;
                    const float4 r7 = __ldg(vrec + (size_t) tri * VREC_QUADS + 7);
                    ws.etrim[lane] = make_float2(r7.z, r7.w);
                }
            }
            int uincl = nun;
            for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, uincl, d); if(lane >= d) uincl += n; }
            const int ustart = uincl - nun;
            const int utotal = __shfl_sync(0xFFFFFFFFu, uincl, 31);       /* <= BAND_UNITS */
            for(int k = 0; k < nun; k++) ws.unit[ustart + k] = (unsigned char) (lane << 3 | (x + k));
            __syncwarp();
            for(int u0 = 0; u0 < utotal; u0 += 32)
            {
                /* stage 1: numerators of v and w (main.c:325-328) down the column; exact cheap rejections */
                const bool act = u0 + lane < utotal;
                const uint32_t o = act ? ws.unit[u0 + lane] : 0u;
                const int src = o >> 3, xl = o & 7;
                const float4 q0 = ws.slab[0][src], q1 = ws.slab[1][src], q2 = ws.slab[2][src];
                const uint32_t bb = ws.bbox[src];
                const float den_hi = q2.w * U_SLACK;
                int y0l = (bb >> 10) & 31;
                int rows = act ? (int) ((bb >> 15) & 31) - y0l + 1 : 0;
                const float eps = (bb >> 20) & 1 ? -GUARD_EPS : -INFINITY;
                const float v2x = gel::sub(gel::i2f(px0 + xl), q0.x);
                const float cx0 = gel::mul(v2x, q0.z), cx1 = gel::mul(v2x, q1.x);
#if GEL_BAND_ROW_TRIM
                {
                    /* Row trimming (gel_math.h: row_trim): the numerators at the column's first and last row -- the row loop's own
                     * operations -- say how many rows at either end cannot pass the cheap tests; a triangle much larger than a band
                     * covers a chord of each bbox column, and the loop below then walks the chord instead of the bbox. */
                    const float2 et = ws.etrim[src];
                    const int n = max(rows - 1, 0);
                    const float ya = gel::sub(gel::i2f(py0 + y0l), q0.y), yb = gel::sub(gel::i2f(py0 + y0l + n), q0.y);
                    const float a20 = gel::add(gel::add(cx0, gel::mul(ya, q0.w)), q1.z), a21 = gel::add(gel::add(cx1, gel::mul(ya, q1.y)), q1.w);
                    const float b20 = gel::add(gel::add(cx0, gel::mul(yb, q0.w)), q1.z), b21 = gel::add(gel::add(cx1, gel::mul(yb, q1.y)), q1.w);
                    const float nv0 = gel::sub(gel::mul(q2.z, a20), gel::mul(q2.y, a21)), nw0 = gel::sub(gel::mul(q2.x, a21), gel::mul(q2.y, a20));
                    const float nv1 = gel::sub(gel::mul(q2.z, b20), gel::mul(q2.y, b21)), nw1 = gel::sub(gel::mul(q2.x, b21), gel::mul(q2.y, b20));
                    int lo, hi;
                    gel::row_trim(nv0, nw0, nv1, nw1, eps, den_hi, et.x, et.y, n, lo, hi);
                    y0l += lo;
                    rows = max(rows - lo - hi, 0);
                }
#endif
                const int maxrows = __reduce_max_sync(0xFFFFFFFFu, rows);
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
                        /* stage 2: a full warp off the top of the stack */
                        __syncwarp();
                        qn -= 32;
                        resolve(qn + lane, true);
                    }
                }
            }
            __syncwarp();
            resolve(lane, lane < qn);
            qn = 0;
            __syncwarp();
        };

        /* ---------------- phase 0: walk the tile's segments; near triangles are rasterised, far ones parked ---------------- */
        int cur = lane < NCHAIN ? __ldg(p.heads + ((size_t) view * p.ntiles + tile) * NCHAIN + lane) : -1;   /* lane c walks chain c */
        bool first_round = true;
        float zthr = 0.0f;
        for(;;)
        {
            /* stage up to BAND_SEGS segments: chain c fills slots [c*4, c*4 + 4) */
            constexpr int PER_CHAIN = BAND_SEGS / NCHAIN;
            if(lane < NCHAIN)
            {
                int k = 0;
                while(cur >= 0 && k < PER_CHAIN)
                {
                    const uint4 d = __ldg(descs + cur);
                    ws.seg_first[lane * PER_CHAIN + k] = (int) d.y;
                    ws.seg_pre[lane * PER_CHAIN + k] = (int) d.z;         /* size for now */
                    cur = (int) d.x; k++;
                }
                for(; k < PER_CHAIN; k++) ws.seg_pre[lane * PER_CHAIN + k] = 0;
            }
            const bool more = __any_sync(0xFFFFFFFFu, cur >= 0);
            __syncwarp();
            const int my_count = ws.seg_pre[lane];
            int incl = my_count;
            for(int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if(lane >= d) incl += n; }
            const int round_entries = __shfl_sync(0xFFFFFFFFu, incl, 31);
            __syncwarp();
            ws.seg_pre[lane] = incl - my_count;
            if(first_round)
            {
                /* short lists are not worth a second phase: everything is "near" */
                zthr = (!more && round_entries < TWO_PHASE_MIN) ? -INFINITY : zthr_view;
                first_round = false;
            }
            __syncwarp();
            for(int e0 = 0; e0 < round_entries; e0 += 32)
            {
                const int e = e0 + lane;
                bool have = false, park = false;
                uint32_t tri = 0, bbox = 0, bound = 0;
                float4 r4 = make_float4(0, 0, 0, 0);
                if(e < round_entries)
                {
                    /* staged segment holding entry e: last slot with seg_pre <= e */
                    int lo = 0;
                    #pragma unroll
                    for(int step = BAND_SEGS / 2; step; step >>= 1) if(ws.seg_pre[lo + step] <= e) lo += step;
                    tri = __ldg(entries + ws.seg_first[lo] + (e - ws.seg_pre[lo]));
                    r4 = __ldg(vrec + (size_t) tri * VREC_QUADS + 4);
                    const uint32_t bx = __float_as_uint(r4.x);
                    if((int) (bx & 0xFFFF) <= px1 && (int) (bx >> 16) >= px0)   /* the triangle's columns reach this band */
                    {
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
                }
                const unsigned pm = __ballot_sync(0xFFFFFFFFu, park);
                if(pm)
                {
                    if(park)
                    {
                        const int slot = nfar + __popc(pm & lt_mask);
                        if(slot < BAND_FAR_CAP) far_rec[slot] = make_uint4(tri, bbox, bound, 0u);
                        else have = true;                                 /* scratch full: rasterise it now */
                    }
                    nfar += __popc(pm);
                }
                if(__any_sync(0xFFFFFFFFu, have)) rasterise_batch(have, tri, r4);
            }
            if(!more) break;
            __syncwarp();
        }

        /* ---------------- phase 1: the parked triangles against the hierarchical depth of what is already drawn ---------------- */
        nfar = min(nfar, BAND_FAR_CAP);
        if(nfar > 0)
        {
            __syncwarp();
            {
                /* the band is one column of four 8x8 blocks: a lane takes the minimum of its row over the 8 columns, one
                 * reduction per group of 8 lanes gives the block's value */
                uint32_t zk = 0xFFFFFFFFu;
                #pragma unroll
                for(int dx = 0; dx < BAND_W; dx++) zk = min(zk, (uint32_t) (ws.keys[band_slot(dx, lane)] >> 32));
                zk = __reduce_min_sync(0xFFu << (lane & 24), zk);
                if((lane & 7) == 0) ws.hiz[lane >> 3] = zk;
            }
            __syncwarp();                                                 /* hiz complete; far_rec was written by this warp */
            for(int e0 = 0; e0 < nfar; e0 += 32)
            {
                const int e = e0 + lane;
                bool have = false;
                uint32_t tri = 0;
                float4 r4 = make_float4(0, 0, 0, 0);
                if(e < nfar)
                {
                    const uint4 rec = far_rec[e];
                    const int gy0 = ((rec.y >> 10) & 31) >> 3, gy1 = ((rec.y >> 15) & 31) >> 3;
                    uint32_t lowest = 0xFFFFFFFFu;
                    for(int gy = gy0; gy <= gy1; gy++) lowest = min(lowest, ws.hiz[gy]);
                    if(!(rec.z < lowest))                                 /* cannot be culled */
                    {
                        tri = rec.x;
                        r4 = __ldg(vrec + (size_t) tri * VREC_QUADS + 4);
                        have = true;
                    }
                }
                if(__any_sync(0xFFFFFFFFu, have)) rasterise_batch(have, tri, r4);
            }
        }
        __syncwarp();

        /* ================= shade the winner of every pixel once (main.c:358-366), write the band back ================= */
        unsigned long long hp = 0, hz = 0;
        #pragma unroll 1
        for(int xl = 0; xl < BAND_W; xl++)
        {
            const int x = px0 + xl, y = py0 + lane;
            if(x > px1 || y > py1) continue;
            const unsigned long long key = ws.keys[band_slot(xl, lane)];
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
                colour = gel::pshade(__ldg(p.tex + (uint32_t) (xx + yy * p.tw)), shading);
            }
            const uint32_t gi = gband + (uint32_t) xl * (uint32_t) p.yres;
            p.pixel[gi] = colour;
            p.zbuf[gi] = z;
            if(HASH) { const uint32_t idx = (uint32_t) (y + x * p.yres); hp += gel::salt_mix(colour, idx); hz += gel::salt_mix(__float_as_uint(z), idx); }
        }
        if(HASH)
        {
            for(int d = 16; d; d >>= 1) { hp += __shfl_xor_sync(0xFFFFFFFFu, hp, d); hz += __shfl_xor_sync(0xFFFFFFFFu, hz, d); }
            if(lane == 0 && (hp | hz)) { atomicAdd(p.hash + 2 * view, hp); atomicAdd(p.hash + 2 * view + 1, hz); }
        }
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");               /* every tensor store this thread issued has landed */
}

} /* namespace gelk */
#endif /* GEL_BAND_CUH */
