/* gelcu.h -- C ABI of the B200 (sm_100a) implementation of gel's per-frame render path.
 *
 * The reference (yellingintothefan/gel, main.c) has no plugin or FFI interface: its hot path is the code
 * between slock() and sunlock() in the frame loop, main.c:504-523 --
 *     reset (main.c:413-417, called at :505), camera basis (:506-512), and per triangle
 *     tviewnrm/tviewtri/tperspective/tviewport/tdraw (:513-522).
 * This header is the seam a maintainer would cut there (INTEGRATION.md shows the patch): the host keeps
 * OBJ/BMP loading, soup expansion, input handling and present; one gelcu_render() call replaces :505-522.
 *
 * Plain C, no C++/torch types.  Every function returns GELCU_OK (0), a positive warning, or a negative
 * error; gelcu_last_error() returns a thread-local message for the most recent non-zero return.
 * There is NO CPU fallback: without a CUDA device gelcu_create() fails with GELCU_E_NOGPU.
 *
 * Threading: calls on one context must be serialised by the caller; different contexts (one per GPU)
 * may be driven concurrently from different host threads or processes.
 */
#ifndef GELCU_H
#define GELCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GELCU_OK            0
/* Frames were produced, but some input left the domain where the reference is defined:
 * a projected bbox left [0,xres-1]x[0,yres-1] (reference: out-of-bounds writes, main.c:344-349) or a
 * texel coordinate left the texture (reference: out-of-bounds read, main.c:360-361,366).  The device
 * clips / clamps instead of emulating undefined behaviour. */
#define GELCU_W_CLIPPED     1
#define GELCU_E_INVALID    (-1)   /* bad argument / call order                           */
#define GELCU_E_CUDA       (-2)   /* a CUDA runtime call failed (message has the detail)  */
#define GELCU_E_NOMEM      (-3)   /* host or device allocation failed                     */
#define GELCU_E_NOGPU      (-4)   /* no usable CUDA device                                */

typedef struct gelcu_ctx gelcu_ctx;   /* one per GPU */

/* One view = the camera basis the reference builds at main.c:509-512 from (xt, yt):
 *   eye = (sinf xt, sinf yt, cosf xt); z = unit(eye - 0); x = unit((0,1,0) x z); y = z x x.
 * It is computed ON THE HOST (libm sinf/cosf) -- gel_view_basis() in gel_host.h -- so the CPU reference
 * and the device consume identical floats. */
typedef struct { float x[3], y[3], z[3], eye[3]; } gelcu_view;

/* Counters of the most recent gelcu_render() call. */
typedef struct
{
    uint64_t kernels_launched;   /* CUDA kernels of this library launched by the call              */
    uint64_t views;              /* views rendered                                                 */
    uint64_t bin_entries;        /* (triangle, tile) pairs binned, summed over views               */
    uint64_t unique_vertices;    /* distinct (position, normal) corners the transform kernel sees  */
    uint64_t triangles;          /* triangles per view                                             */
    uint64_t h2d_bytes;          /* bytes copied host->device by the call (views)                  */
    uint64_t d2h_bytes;          /* bytes copied device->host by the call (frames, z, checksums)   */
    float    ms_transform;       /* CUDA-event time of the vertex-transform kernels                */
    float    ms_bin;             /* ... of triangle setup + tile binning (count, scan, fill)       */
    float    ms_raster;          /* ... of the raster stage (all its kernels)                      */
    float    ms_dominant;        /* ... of the dominant kernel alone (raster_kernel / direct_raster_kernel<0>) */
    float    ms_total;           /* first kernel start to last kernel end (no copies)              */
    uint32_t flags;              /* OR of per-view device flags: 1 = bbox clipped, 2 = texel clamped */
    uint32_t batches;            /* view batches the call was split into                           */
    uint32_t pipeline;           /* 1 = tile pipeline, 2 = direct pipeline (meshes of tiny triangles) */
    uint32_t reserved;
}
gelcu_stats;

/* Number of CUDA devices, or a negative error. */
int gelcu_device_count(void);

/* Creates a context bound to `device` rendering at xres x yres (the reference's Sdl.xres/yres, main.c:434-445). */
int gelcu_create(gelcu_ctx** ctx, int device, int xres, int yres);

/* Mesh = the three triangle soups of main.c:496-498 (tv: positions already scaled by tvgen, tn: normals,
 * tt: texture coordinates), 9 floats per triangle in the reference's Triangle layout (a.xyz b.xyz c.xyz,
 * main.c:8-12,45-49).  Copied; the caller may free its arrays.  ntri may be 0. */
int gelcu_set_mesh(gelcu_ctx* ctx, const float* tv, const float* tn, const float* tt, int ntri);

/* Mesh, indexed: the OBJ arrays as the reference holds them after oparse (main.c:129-180) --
 *   v / vt / vn   3 floats per `v`, `vt`, `vn` line (the reference's Vertices.vertex, main.c:8-20); v is UNSCALED
 *   faces         the reference's Face array (main.c:22-28): 9 ints { va,vb,vc, ta,tb,tc, na,nb,nc }, 0-based
 * and replaces vmaxlen + tvgen / ttgen / tngen (main.c:233-286): the scale 1.0f / (int) max|v| and the per-corner
 * gathers run on the device (one multiply per component, bit-identical to tvgen's tmul), corners are merged per
 * (position, normal) index pair, and the host never builds or re-hashes the 108-byte-per-triangle soups.
 * Frames are bit-identical to gelcu_set_mesh on the soups the reference would have generated.  Errors:
 * GELCU_E_INVALID for an index outside its array or max|v| < 1 (the reference divides by (int) 0, main.c:244,253). */
int gelcu_set_mesh_indexed(gelcu_ctx* ctx, const float* v, int nv, const float* vt, int nvt, const float* vn, int nvn,
                           const int* faces, int nfaces);

/* Texture = fdif->pixels / w / h (main.c:359-361,366): XRGB8888, top-down, pitch 4*w.  Copied. */
int gelcu_set_texture(gelcu_ctx* ctx, const uint32_t* xrgb, int w, int h);

/* Renders `nviews` frames; replaces main.c:505-522 for each of them.
 *   pixel_out  host, nviews*xres*yres uint32 in the reference's sideways order (index y + x*yres,
 *              main.c:356,365-366,441), values 0x00RRGGBB; or NULL to leave frames on the device
 *   z_out      host, nviews*xres*yres float (the reference's zbuff, main.c:500); or NULL
 *   hash_out   host, 2*nviews: [2k] = checksum of view k's pixel words, [2k+1] = of its z bit patterns,
 *              computed on the device as  sum_i mix32(word_i ^ i*0x9E3779B1) mod 2^64  with
 *              mix32(h): h*=0x85EBCA6B; h^=h>>13; h*=0xC2B2AE35; h^=h>>16  (order-independent, so tiles
 *              can contribute in any order); or NULL
 *   device_ms  CUDA-event time of the kernels only (no copies), summed over batches; or NULL
 * Host pointers may be pageable or pinned (gelcu_host_alloc); pinned buffers copy faster and overlap
 * with rendering.  With pixel_out == z_out == NULL the last batch stays readable via gelcu_read_frame. */
int gelcu_render(gelcu_ctx* ctx, const gelcu_view* views, int nviews,
                 uint32_t* pixel_out, float* z_out, uint64_t* hash_out, float* device_ms);

/* Frame sink: the same render, but each frame leaves the device in presentation form --
 * un-rotated the way schurn presents it (SDL_RenderCopyEx by -90 degrees, main.c:424-432) and packed to 24 bits:
 *   rgb_out    host, nviews * yres rows * xres pixels * 3 bytes, top row first:
 *              rgb_out[k][(wy*xres + wx)*3 + {0,1,2}] = {R,G,B} of view k's pixel[(yres-1-wy) + wx*yres]
 *              (the body of a binary PPM "P6 xres yres 255").  The device -> host copy carries 3 bytes per pixel
 *              instead of 4 and the host does not touch the pixels again.
 *   hash_out, device_ms  as in gelcu_render (checksums are of the sideways XRGB / z frames; device_ms excludes
 *              the sink kernel, like it excludes copies). */
int gelcu_render_rgb8(gelcu_ctx* ctx, const gelcu_view* views, int nviews,
                      uint8_t* rgb_out, uint64_t* hash_out, float* device_ms);

/* Region output: only the part of each frame the mesh can touch crosses PCIe.
 * A view's REGION is the screen bounding box of its transformed vertices (every bbox of main.c:344-347 lies inside it),
 * clipped to the frame and widened to multiples of 8; everything outside holds the reset values of main.c:413-417
 * (pixel 0, z -FLT_MAX).  gelcu_render_region copies just that rectangle of every frame into the caller's full-size
 * frames and keeps the rest of each frame correct with a dirty-rectangle contract:
 *   rect_io[k] in   the part of frame k of pixel_io / z_io that may currently hold non-reset values:
 *                   x1 < x0 = none (the frame is all reset values, e.g. a zeroed pixel buffer);
 *                   { 0, 0, xres-1, yres-1 } = anything (the library resets the whole frame on the host);
 *                   or the rectangle a previous call returned for this frame slot -- then only the strips that the
 *                   old rectangle covers and the new one does not are reset, on the host, while the copies run
 *   rect_io[k] out  the view's region (inclusive; x1 < x0 when the mesh is off screen)
 * On return every frame is complete and bit-identical to what gelcu_render writes.  rgb8 = 0: pixel_io (and z_io, may
 * be NULL) are sideways frames as in gelcu_render; rgb8 = 1: pixel_io is the upright 24-bit layout of
 * gelcu_render_rgb8 (z_io must be NULL).  This is the reference's own frame loop pattern -- one canvas reused every
 * frame (slock/sunlock, main.c:504,523) -- without re-sending the background each time. */
typedef struct { int x0, y0, x1, y1; } gelcu_rect;
int gelcu_render_region(gelcu_ctx* ctx, const gelcu_view* views, int nviews, void* pixel_io, float* z_io,
                        gelcu_rect* rect_io, int rgb8, uint64_t* hash_out, float* device_ms);

/* Copies frame `slot` (0-based within the LAST batch of the previous gelcu_render) to the host. */
int gelcu_read_frame(gelcu_ctx* ctx, int slot, uint32_t* pixel_out, float* z_out);

/* Tunables, by name (returns GELCU_E_INVALID for unknown names):
 *   "batch_views"   views rendered per kernel launch set (default: sized so frames fit ~8 GB)
 *   "raster_ctas_per_sm"   persistent rasteriser CTAs per SM, 1..16 (default: 1024 threads per SM)
 *   "stage_timing"  1 = record per-stage CUDA events into gelcu_stats (default 1)
 *   "pipeline"      0 = chosen from the mesh (default), 1 = tile pipeline, 2 = direct pipeline
 *   "compact_records"  1 (default) = meshes with fewer than 2^21 distinct vertices get 32-byte per-triangle shading
 *                   records (three 21-bit indices), 0 = always the 64-byte form; read by the next gelcu_set_mesh
 *   "graph_small_calls"  1 (default) = calls of up to 4 views replay a captured CUDA graph (one launch instead of ~20 API calls:
 *                   the interactive one-view-per-frame use); stage timings are not recorded for such calls
 *   "fill_mode", "fill_ctas_per_sm", "fill_sleep_ns", "red_hint", "store_hint"   direct pipeline experiments with the background
 *                   reset and L2 cache-policy hints (defaults: plain trailing fill, evict-last hint on the key REDs)
 *   "raster_mode"   tile pipeline: 1 (default) = a warp per band (raster_band_kernel), 0 = a CTA per tile (raster_kernel)
 *   "tma_reset"     tile pipeline: 1 (default) = untouched tiles are reset by TMA tensor stores, 0 = by a store loop
 *   "near_carveout", "band_carveout"   shared-memory carve-out hint (percent, -1 = driver default) of the direct pipeline's near pass / the
 *                   band rasteriser: measurement knobs (profiles/README.md, session 3); the defaults are the fastest */
int gelcu_set_option(gelcu_ctx* ctx, const char* name, int value);
int gelcu_get_stats(gelcu_ctx* ctx, gelcu_stats* out);

/* Stage introspection for the parity tests (not used by the render path):
 *   gelcu_debug_transform: runs the vertex-transform kernel for one view and expands the result per
 *       corner: vew receives 9 floats per triangle (screen x, y, z of a, b, c -- the reference's `vew`,
 *       main.c:519), shade 3 floats per triangle (vdot(lights, nrm.{a,b,c}), main.c:358).
 *   gelcu_debug_bins: runs transform + binning for one view; counts receives one int per screen tile
 *       (tile = tx*tiles_y + ty), entries the triangle indices tile by tile (ascending per tile), at
 *       most `cap` of them; *total is the number of (triangle, tile) pairs. */
int gelcu_debug_transform(gelcu_ctx* ctx, const gelcu_view* view, float* vew, float* shade);
int gelcu_debug_bins(gelcu_ctx* ctx, const gelcu_view* view, int* counts, int* entries, int cap, int* total);
int gelcu_tile_grid(gelcu_ctx* ctx, int* tile_w, int* tile_h, int* tiles_x, int* tiles_y);

/* Page-locked host memory for pixel_out / z_out. */
int  gelcu_host_alloc(void** p, size_t bytes);
void gelcu_host_free(void* p);

void gelcu_destroy(gelcu_ctx* ctx);
const char* gelcu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GELCU_H */
