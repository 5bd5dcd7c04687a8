/* gel_host.h -- host-side C flow around the render path: the parts of the reference's main.c that stay on
 * the CPU when lines 505-522 move to the GPU.  Pure C99, no CUDA, no SDL.
 *
 *   gel_obj_load      oload + oparse + tvgen/ttgen/tngen     main.c:460-469, 84-180, 227-286   (files > 4 MB: parsed in
 *                     line-aligned chunks by up to 32 threads, GEL_PARSE_THREADS overrides; same arrays as one pass)
 *   gel_bmp_load      sload (IMG_Load + convert to RGB888)    main.c:471-484   (uncompressed 8/24/32-bit BMP, no SDL_image)
 *   gel_view_basis    camera basis from (xt, yt)              main.c:506-512
 *   gel_input_step    ipump's angle update                    main.c:408-409
 *   gel_upright       schurn's -90 degree rotation            main.c:424-432   (for image export)
 *
 * Build with -ffp-contract=off: gel_view_basis feeds both the CPU oracle and the device and must be the
 * reference's exact fp32 sequence.
 */
#ifndef GEL_HOST_H
#define GEL_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    float* tv;     /* positions, scaled by 1.0f/(int)maxlen (main.c:244,253); 9 floats per triangle */
    float* tn;     /* normals                                                                        */
    float* tt;     /* texture coordinates (x, y, z-as-parsed)                                        */
    int ntri;
    int nv, nvt, nvn;   /* counts of v / vt / vn lines read */
}
GelMesh;

/* The indexed OBJ as the reference holds it after oparse (main.c:129-180, the `Obj` struct at :37-43): three floats per
 * v / vt / vn line and one `Face` per f line -- 9 ints { va,vb,vc, ta,tb,tc, na,nb,nc }, 0-based (main.c:22-28,165-170).
 * This is what gelcu_set_mesh_indexed takes: the soups (main.c:242-286) are then generated on the device. */
typedef struct
{
    float *v, *vt, *vn;
    int* faces;
    int nv, nvt, nvn, nfaces;
    int parse_threads;      /* threads the parse used (gel_obj_expand reuses the count) */
}
GelObj;

typedef struct { uint32_t* pixels; int w, h; } GelTexture;   /* XRGB8888 top-down, pitch 4*w */

/* 0 on success; -1 cannot open, -2 malformed (index out of range, no faces' data), -3 out of memory,
 * -4 max|v| < 1 (the reference divides by (int)maxlen == 0, main.c:244,253). */
int  gel_obj_load(const char* path, GelMesh* out);
void gel_mesh_free(GelMesh* m);

/* The two halves of gel_obj_load: text -> indexed arrays (oparse), indexed arrays -> soups (tvgen/ttgen/tngen on the
 * host).  gel_obj_parse alone + gelcu_set_mesh_indexed skips the host expansion.  Same return codes. */
int  gel_obj_parse(const char* path, GelObj* out);
int  gel_obj_expand(const GelObj* obj, GelMesh* out);
void gel_obj_free(GelObj* o);

/* 0 on success; -1 cannot open, -2 not an uncompressed 8/24/32-bit BMP (or larger than 32768 per side),
 * -3 out of memory, -4 truncated / inconsistent header. */
int  gel_bmp_load(const char* path, GelTexture* out);
void gel_texture_free(GelTexture* t);

/* basis[12] = x[3], y[3], z[3], eye[3]; layout-compatible with gelcu_view. */
void gel_view_basis(float xt, float yt, float basis[12]);

/* xt -= sens*dx; yt += sens*dy with sens = 0.005f (main.c:394, 408-409). */
void gel_input_step(float* xt, float* yt, int dx, int dy);

/* FNV-1a-64 folded over 32-bit words. */
uint64_t gel_fnv1a64_words(const uint32_t* w, uint64_t n);

/* Sideways canvas (index y + x*yres) -> upright row-major xres*yres image, as schurn presents it:
 * window(wx, wy) = canvas[(yres-1-wy) + wx*yres]. */
void gel_upright(const uint32_t* canvas, int xres, int yres, uint32_t* upright);
int  gel_write_ppm(const char* path, const uint32_t* canvas, int xres, int yres);

#ifdef __cplusplus
}
#endif
#endif
