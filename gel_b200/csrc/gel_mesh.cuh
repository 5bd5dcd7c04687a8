/* gel_mesh.cuh -- load-time soup generation on the device.
 *
 * The reference turns the indexed OBJ into three triangle soups on the host: vmaxlen + tvgen (positions scaled by
 * 1.0f / (int) max|v|), tngen, ttgen -- main.c:233-286 -- 108 bytes per triangle, every shared corner repeated.  Here the
 * indexed arrays go to the device as they are and four small kernels build the render path's own layout from them:
 *
 *   M1 mesh_maxlen_kernel   max over the `v` lines of vlen(v) = sqrtf((x*x + y*y) + z*z)          main.c:205-208,233-240
 *   M2 mesh_pairs_kernel    every corner's (position index, normal index) pair goes into a hash set; a pair that is new
 *                           takes the next local rank of its position (positions with several normals = hard edges)
 *   M3 mesh_scan_kernel     exclusive scan of the per-position pair counts -> first vertex id of every position
 *   M4 mesh_emit_kernel     per triangle: ids of its three corners (first id of the position + the pair's local rank),
 *                           the merged vertex arrays (position * inv scale -- one multiply per component, exactly tmul
 *                           of main.c:221-225,253 -- and the normal), the uv pairs (main.c:273-286; uv.z is never read) and
 *                           the resolve pass's packed per-triangle record
 *
 * Vertex ids follow the position order of the file, so consecutive triangles keep touching neighbouring vertices (the
 * raster kernels gather xf[i0], xf[i1], xf[i2] per triangle).  Frames do not depend on the numbering.
 */
#ifndef GEL_MESH_CUH
#define GEL_MESH_CUH

#include "gel_direct.cuh"

namespace gelk {

constexpr unsigned long long PAIR_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t MESH_BAD_INDEX = 1u;

struct MeshBuild
{
    const float *v, *vt, *vn; const int* faces; int nv, nvt, nvn, nfaces;
    unsigned long long* keys; uint32_t* ranks; uint32_t mask;          /* hash set of (va << 32 | na), local rank per entry */
    uint32_t* count;           /* [nv] distinct normals paired with the position; after M3: first vertex id */
    uint32_t* result;          /* [0] max|v| bits, [1] flags, [2] number of merged vertices */
    double* area;              /* sum over triangles of the model-space area (pipeline choice) */
};

__device__ __forceinline__ uint32_t pair_slot(unsigned long long key, uint32_t mask)
{
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    return (uint32_t) h & mask;
}

/* M1: vmaxlen (main.c:233-240).  `vlen(v) > max` ignores NaN lengths; non-negative floats order like their bit patterns. */
__global__ void __launch_bounds__(256)
mesh_maxlen_kernel(MeshBuild m)
{
    uint32_t best = 0u;
    for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.nv; i += gridDim.x * blockDim.x)
    {
        const float x = m.v[3 * (size_t) i], y = m.v[3 * (size_t) i + 1], z = m.v[3 * (size_t) i + 2];
        const float len = gel::root(gel::add(gel::add(gel::mul(x, x), gel::mul(y, y)), gel::mul(z, z)));
        if(len == len) best = max(best, __float_as_uint(len));
    }
    best = __reduce_max_sync(0xFFFFFFFFu, best);
    if((threadIdx.x & 31) == 0 && best) atomicMax(m.result, best);
}

/* M2 */
__global__ void __launch_bounds__(256)
mesh_pairs_kernel(MeshBuild m)
{
    const size_t ncorner = (size_t) m.nfaces * 3;
    for(size_t c = (size_t) blockIdx.x * blockDim.x + threadIdx.x; c < ncorner; c += (size_t) gridDim.x * blockDim.x)
    {
        const int* face = m.faces + 9 * (c / 3);
        const int k = (int) (c % 3);
        const int va = face[k], ta = face[3 + k], na = face[6 + k];
        if(va < 0 || va >= m.nv || ta < 0 || ta >= m.nvt || na < 0 || na >= m.nvn) { atomicOr(m.result + 1, MESH_BAD_INDEX); continue; }
        const unsigned long long key = (unsigned long long) (uint32_t) va << 32 | (uint32_t) na;
        for(uint32_t slot = pair_slot(key, m.mask);; slot = (slot + 1) & m.mask)
        {
            const unsigned long long seen = atomicCAS(m.keys + slot, PAIR_EMPTY, key);
            if(seen == PAIR_EMPTY) { m.ranks[slot] = atomicAdd(m.count + va, 1u); break; }
            if(seen == key) break;
        }
    }
}

/* M3: one CTA walks the array in chunks of 1024 with a running carry (load time; 0.5 M positions take ~0.3 ms) */
__global__ void __launch_bounds__(1024)
mesh_scan_kernel(MeshBuild m)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for(int base = 0; base < m.nv; base += 1024)
    {
        const int i = base + threadIdx.x;
        const uint32_t mine = i < m.nv ? m.count[i] : 0u;
        uint32_t incl = mine;
        for(int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if(lane >= d) incl += n; }
        if(lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if(warp == 0)
        {
            uint32_t w = warp_sum[lane];
            for(int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, w, d); if(lane >= d) w += n; }
            warp_sum[lane] = w;                                        /* inclusive over the warps */
        }
        __syncthreads();
        const uint32_t before = carry + (warp ? warp_sum[warp - 1] : 0u) + incl - mine;
        if(i < m.nv) m.count[i] = before;
        __syncthreads();
        if(threadIdx.x == 1023) carry = before + mine;
        __syncthreads();
    }
    if(threadIdx.x == 0) m.result[2] = carry;
}

struct MeshOut
{
    float4 *vpos, *vnrm; uint32_t *i0, *i1, *i2; float2* uv; uint4* trec; int compact; float inv;
};

/* M4: one thread per triangle */
__global__ void __launch_bounds__(256)
mesh_emit_kernel(MeshBuild m, MeshOut o)
{
    double area = 0.0;
    for(int t = blockIdx.x * blockDim.x + threadIdx.x; t < m.nfaces; t += gridDim.x * blockDim.x)
    {
        const int* face = m.faces + 9 * (size_t) t;
        uint32_t id[3];
        float2 tex[3];
        float3 pos[3];
        #pragma unroll
        for(int k = 0; k < 3; k++)
        {
            const int va = face[k], ta = face[3 + k], na = face[6 + k];
            const unsigned long long key = (unsigned long long) (uint32_t) va << 32 | (uint32_t) na;
            uint32_t slot = pair_slot(key, m.mask);
            while(m.keys[slot] != key) slot = (slot + 1) & m.mask;
            id[k] = m.count[va] + m.ranks[slot];
            /* tvgen: tmul(t, 1.0f / scale), main.c:253 -- every writer of a merged vertex stores the same bits */
            pos[k] = make_float3(gel::mul(m.v[3 * (size_t) va], o.inv), gel::mul(m.v[3 * (size_t) va + 1], o.inv), gel::mul(m.v[3 * (size_t) va + 2], o.inv));
            o.vpos[id[k]] = make_float4(pos[k].x, pos[k].y, pos[k].z, 0.0f);
            o.vnrm[id[k]] = make_float4(m.vn[3 * (size_t) na], m.vn[3 * (size_t) na + 1], m.vn[3 * (size_t) na + 2], 0.0f);
            tex[k] = make_float2(m.vt[3 * (size_t) ta], m.vt[3 * (size_t) ta + 1]);
            o.uv[3 * (size_t) t + k] = tex[k];
        }
        o.i0[t] = id[0]; o.i1[t] = id[1]; o.i2[t] = id[2];
        if(o.compact)
        {
            const unsigned long long packed = (unsigned long long) id[0] | (unsigned long long) id[1] << 21 | (unsigned long long) id[2] << 42;
            o.trec[2 * (size_t) t] = make_uint4((uint32_t) packed, (uint32_t) (packed >> 32), __float_as_uint(tex[0].x), __float_as_uint(tex[0].y));
            o.trec[2 * (size_t) t + 1] = make_uint4(__float_as_uint(tex[1].x), __float_as_uint(tex[1].y), __float_as_uint(tex[2].x), __float_as_uint(tex[2].y));
        }
        else
        {
            o.trec[(size_t) TREC_QUADS * t] = make_uint4(id[0], id[1], id[2], 0u);
            o.trec[(size_t) TREC_QUADS * t + 1] = make_uint4(__float_as_uint(tex[0].x), __float_as_uint(tex[0].y), __float_as_uint(tex[1].x), __float_as_uint(tex[1].y));
            o.trec[(size_t) TREC_QUADS * t + 2] = make_uint4(__float_as_uint(tex[2].x), __float_as_uint(tex[2].y), 0u, 0u);
            o.trec[(size_t) TREC_QUADS * t + 3] = make_uint4(0u, 0u, 0u, 0u);
        }
        /* model-space area, only for the tile / direct pipeline choice (a heuristic: double precision, any order) */
        const double ux = (double) pos[1].x - pos[0].x, uy = (double) pos[1].y - pos[0].y, uz = (double) pos[1].z - pos[0].z;
        const double wx = (double) pos[2].x - pos[0].x, wy = (double) pos[2].y - pos[0].y, wz = (double) pos[2].z - pos[0].z;
        const double cx = uy * wz - uz * wy, cy = uz * wx - ux * wz, cz = ux * wy - uy * wx;
        const double a2 = cx * cx + cy * cy + cz * cz;
        if(a2 == a2 && a2 < 1e30) area += 0.5 * sqrt(a2);
    }
    for(int d = 16; d; d >>= 1) area += __shfl_xor_sync(0xFFFFFFFFu, area, d);
    if((threadIdx.x & 31) == 0 && area != 0.0) atomicAdd(m.area, area);
}

} /* namespace gelk */
#endif /* GEL_MESH_CUH */
