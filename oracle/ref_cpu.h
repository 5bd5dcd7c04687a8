/* TEST INFRASTRUCTURE (oracle) -- CPU restatement of gel's per-frame render path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the reported CPU baseline; the product path never does.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit (tests/test_oracle.py) against the
 * UNMODIFIED reference main.c built with the headless SDL shim (oracle/verbatim, oracle/_ref) at 800x600,
 * and against the committed golden vectors in tests/golden/ that were produced by that verbatim binary.
 * The one un-pinned boundary is texture decode (SDL2_image, un-vendored): see verbatim/shim.c.
 */
#ifndef GEL_ORACLE_REF_CPU_H
#define GEL_ORACLE_REF_CPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-frame workload counters (what SURVEY.md §8(a) calls tested / inside / z-pass / lit) */
typedef struct
{
    uint64_t tested;  /* bbox pixels visited              main.c:348-351 */
    uint64_t inside;  /* passed v,w,u >= 0                main.c:352     */
    uint64_t zpass;   /* passed z > zbuff                 main.c:356     */
    uint64_t lit;     /* pixels with zbuff != -FLT_MAX at end of frame   */
}
RefCounters;

/* basis[12] = x[3], y[3], z[3], eye[3]                    main.c:506-512 */
void ref_view_basis(float xt, float yt, float basis[12]);

/* per-triangle transform only; vew/nrm receive 9 floats per triangle (a,b,c)    main.c:515-519 */
void ref_transform(const float* tv, const float* tn, int ntri, const float basis[12],
                   int xres, int yres, float* vew, float* nrm);

/* one frame: reset + transform + raster                   main.c:505-522
 * tv/tn/tt: 9 floats per triangle; tex: XRGB8888 top-down tw x th; pixel/zbuff: xres*yres, index y + x*yres.
 * Returns 0 for in-domain input, else flag bits for the two cases where the reference itself is undefined (it would
 * access memory out of bounds); this restatement defines both the way the product does, so the two stay comparable:
 *   1  a triangle's bbox left [0,xres-1]x[0,yres-1]: the bbox is clipped to the screen        
 *   2  a texel coordinate left [0,tw-1]x[0,th-1]: it is clamped                                (main.c:360-366) */
int ref_render(const float* tv, const float* tn, const float* tt, int ntri,
               const uint32_t* tex, int tw, int th, int xres, int yres, const float basis[12],
               uint32_t* pixel, float* zbuff, RefCounters* counters);

/* frames-parallel batch: nviews bases (12 floats each) over nthreads POSIX threads, private framebuffers.
 * pixel_out/z_out: nviews frames or NULL; hash_pixel/hash_z: nviews position-salted checksums or NULL.
 * seconds: wall time of the batch (clock_gettime) or NULL. */
int ref_render_views(const float* tv, const float* tn, const float* tt, int ntri,
                     const uint32_t* tex, int tw, int th, int xres, int yres,
                     const float* bases, int nviews, int nthreads,
                     uint32_t* pixel_out, float* z_out, uint64_t* hash_pixel, uint64_t* hash_z,
                     double* seconds);

/* FNV-1a-64 folded over 32-bit words (the survey's KAT hash) and the order-independent position-salted
 * checksum the CUDA library also reports (include/gelcu.h, gelcu_render hash_out). */
uint64_t ref_fnv1a64_words(const uint32_t* w, uint64_t n);
uint64_t ref_salted_sum(const uint32_t* w, uint64_t n);

/* Host-flow restatement: OBJ text -> three triangle soups     main.c:84-180, 227-286.
 * On success returns ntri and malloc'd arrays of 9*ntri floats (caller frees with ref_free). */
int ref_load_obj(const char* path, float** tv, float** tn, float** tt);
/* 24-bit BMP -> XRGB8888 top-down (the shim's definition of main.c:471-484) */
int ref_load_bmp(const char* path, uint32_t** xrgb, int* w, int* h);
void ref_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
