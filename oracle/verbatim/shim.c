/* TEST INFRASTRUCTURE (oracle) -- never linked into the product.
 *
 * Headless implementation of the 19 SDL2 / SDL2_image symbols the reference program imports
 * (/root/reference/main.c:1-2, used at :395, :404, :407, :421, :431, :437-441, :449, :456, :473-482,
 * :503, :528), so that main.c builds and runs UNMODIFIED without a display.  Written from scratch;
 * none of this is SDL code.
 *
 * Behaviour is scripted through environment variables:
 *   GELSHIM_FRAMES=N        render N frames, then raise SDL_QUIT              (default 1)
 *   GELSHIM_DX=i GELSHIM_DY=j   relative mouse motion reported after every frame (default 0 0);
 *                           the reference turns it into xt -= 0.005f*dx, yt += 0.005f*dy (main.c:408-409)
 *   GELSHIM_SCRIPT=path     text file of "dx dy" pairs, one per frame transition (overrides DX/DY)
 *   GELSHIM_DUMP=path       append each presented frame, raw uint32[xres*yres] in the reference's own
 *                           sideways order (index y + x*yres, main.c:356,441)
 *   GELSHIM_BARRIER=dir:N   before the first frame, rendezvous with N-1 sibling processes through
 *                           marker files in dir (used by the frames-parallel CPU baseline)
 * One JSON line per frame goes to stdout: frame index, FNV-1a-64 over the uint32 words, count of
 * non-zero pixels and the wall time between SDL_LockTexture and SDL_UnlockTexture (= reset +
 * transform + raster, main.c:504-523).
 *
 * Texture decode ("IMG_Load + SDL_ConvertSurface to RGB888", main.c:473-480) is defined here as:
 * uncompressed 24-bit bottom-up (or top-down, negative height) BMP -> 0x00RRGGBB, row 0 = top of the
 * image, pitch = 4*w.  SDL2_image is un-vendored and unpinned in the reference (Makefile:5), so this
 * boundary is "parity unpinned" (see DESIGN.md).
 */
#define _DEFAULT_SOURCE
#include <SDL2/SDL.h>
#include <SDL2/SDL_image.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <dirent.h>

struct SDL_Texture { int w, h; uint32_t* pixels; };
struct SDL_Window { int unused; };
struct SDL_Renderer { int unused; };
struct SDL_PixelFormat { uint32_t format; };

static struct SDL_Window the_window;
static struct SDL_Renderer the_renderer;
static struct SDL_Texture* the_canvas;
static int frames_wanted = 1, frames_done = 0;
static int const_dx = 0, const_dy = 0;
static int* script = NULL; static int script_len = 0;
static FILE* dump = NULL;
static double t_lock = 0.0, render_ms = 0.0;
static char img_error[256] = "no error";

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int env_int(const char* name, int dflt)
{
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

static void rendezvous(void)
{
    const char* spec = getenv("GELSHIM_BARRIER");
    if(!spec) return;
    char dir[512];
    const char* colon = strrchr(spec, ':');
    if(!colon) return;
    const int n = atoi(colon + 1);
    snprintf(dir, sizeof dir, "%.*s", (int)(colon - spec), spec);
    char path[600];
    snprintf(path, sizeof path, "%s/ready.%d", dir, (int) getpid());
    FILE* f = fopen(path, "w"); if(f) fclose(f);
    for(;;)
    {
        int seen = 0;
        DIR* d = opendir(dir);
        if(!d) return;
        for(struct dirent* e; (e = readdir(d)); ) if(strncmp(e->d_name, "ready.", 6) == 0) seen++;
        closedir(d);
        if(seen >= n) return;
        usleep(2000);
    }
}

int SDL_Init(uint32_t flags)
{
    (void) flags;
    frames_wanted = env_int("GELSHIM_FRAMES", 1);
    const_dx = env_int("GELSHIM_DX", 0);
    const_dy = env_int("GELSHIM_DY", 0);
    const char* sp = getenv("GELSHIM_SCRIPT");
    if(sp)
    {
        FILE* f = fopen(sp, "r");
        if(f)
        {
            int cap = 1024, a, b;
            script = (int*) malloc(sizeof(int) * 2 * cap);
            while(fscanf(f, "%d %d", &a, &b) == 2)
            {
                if(script_len == cap) script = (int*) realloc(script, sizeof(int) * 2 * (cap *= 2));
                script[2 * script_len] = a; script[2 * script_len + 1] = b; script_len++;
            }
            fclose(f);
        }
    }
    const char* dp = getenv("GELSHIM_DUMP");
    if(dp) dump = fopen(dp, "wb");
    return 0;
}

int SDL_CreateWindowAndRenderer(int w, int h, uint32_t flags, SDL_Window** win, SDL_Renderer** ren)
{
    (void) w; (void) h; (void) flags;
    *win = &the_window; *ren = &the_renderer;
    return 0;
}

void SDL_SetWindowTitle(SDL_Window* win, const char* title) { (void) win; (void) title; }

SDL_Texture* SDL_CreateTexture(SDL_Renderer* ren, uint32_t format, int access, int w, int h)
{
    (void) ren; (void) format; (void) access;
    struct SDL_Texture* t = (struct SDL_Texture*) malloc(sizeof *t);
    t->w = w; t->h = h;
    t->pixels = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) w * h);
    the_canvas = t;
    return t;
}

int SDL_LockTexture(SDL_Texture* tex, const SDL_Rect* rect, void** pixels, int* pitch)
{
    (void) rect;
    if(frames_done == 0) rendezvous();
    *pixels = tex->pixels; *pitch = tex->w * 4;
    t_lock = now_ms();
    return 0;
}

void SDL_UnlockTexture(SDL_Texture* tex) { (void) tex; render_ms = now_ms() - t_lock; }

int SDL_RenderCopyEx(SDL_Renderer* ren, SDL_Texture* tex, const SDL_Rect* src, const SDL_Rect* dst,
                     double angle, const SDL_Point* center, int flip)
{
    (void) ren; (void) tex; (void) src; (void) dst; (void) angle; (void) center; (void) flip;
    return 0;
}

void SDL_RenderPresent(SDL_Renderer* ren)
{
    (void) ren;
    const size_t n = (size_t) the_canvas->w * the_canvas->h;
    uint64_t h = 0xcbf29ce484222325ull;
    size_t nonzero = 0;
    for(size_t i = 0; i < n; i++)
    {
        h = (h ^ the_canvas->pixels[i]) * 0x100000001b3ull;
        nonzero += the_canvas->pixels[i] != 0;
    }
    if(dump) fwrite(the_canvas->pixels, sizeof(uint32_t), n, dump);
    printf("{\"frame\": %d, \"fnv\": \"%016llx\", \"nonzero\": %zu, \"render_ms\": %.4f}\n",
           frames_done, (unsigned long long) h, nonzero, render_ms);
    frames_done++;
}

int SDL_PollEvent(SDL_Event* event)
{
    /* The reference reads event.type even when no event is pending (main.c:403-405): always write it. */
    event->type = frames_done >= frames_wanted ? SDL_QUIT : 0;
    if(event->type == SDL_QUIT) { if(dump) fclose(dump); dump = NULL; fflush(stdout); }
    return 0;
}

uint32_t SDL_GetRelativeMouseState(int* x, int* y)
{
    const int k = frames_done - 1; /* transition after frame k */
    if(script && k >= 0 && k < script_len) { *x = script[2 * k]; *y = script[2 * k + 1]; }
    else { *x = const_dx; *y = const_dy; }
    return 0;
}

int SDL_SetRelativeMouseMode(SDL_bool enabled) { (void) enabled; return 0; }
uint32_t SDL_GetTicks(void) { return (uint32_t) now_ms(); }
void SDL_Delay(uint32_t ms) { (void) ms; }

SDL_PixelFormat* SDL_AllocFormat(uint32_t format)
{
    struct SDL_PixelFormat* f = (struct SDL_PixelFormat*) malloc(sizeof *f);
    f->format = format;
    return f;
}

SDL_Surface* SDL_ConvertSurface(SDL_Surface* src, const SDL_PixelFormat* fmt, uint32_t flags)
{
    (void) fmt; (void) flags;
    SDL_Surface* s = (SDL_Surface*) malloc(sizeof *s);
    *s = *src;
    s->pixels = malloc((size_t) src->pitch * src->h);
    memcpy(s->pixels, src->pixels, (size_t) src->pitch * src->h);
    return s;
}

void SDL_FreeFormat(SDL_PixelFormat* fmt) { free(fmt); }
void SDL_FreeSurface(SDL_Surface* s) { if(s) { free(s->pixels); free(s); } }

static uint32_t rd32(const unsigned char* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t) p[3] << 24; }

SDL_Surface* IMG_Load(const char* path)
{
    FILE* f = fopen(path, "rb");
    if(!f) { snprintf(img_error, sizeof img_error, "Couldn't open %s", path); return NULL; }
    unsigned char hdr[54];
    if(fread(hdr, 1, 54, f) != 54 || hdr[0] != 'B' || hdr[1] != 'M')
    { snprintf(img_error, sizeof img_error, "Not a BMP: %s", path); fclose(f); return NULL; }
    const uint32_t off = rd32(hdr + 10);
    const int32_t w = (int32_t) rd32(hdr + 18);
    const int32_t hs = (int32_t) rd32(hdr + 22);
    const int bpp = hdr[28] | hdr[29] << 8;
    const uint32_t comp = rd32(hdr + 30);
    if(bpp != 24 || comp != 0 || w <= 0 || hs == 0)
    { snprintf(img_error, sizeof img_error, "Unsupported BMP (need 24-bit uncompressed): %s", path); fclose(f); return NULL; }
    const int h = hs < 0 ? -hs : hs;
    const size_t rowbytes = ((size_t) w * 3 + 3) & ~(size_t) 3;
    unsigned char* row = (unsigned char*) malloc(rowbytes);
    SDL_Surface* s = (SDL_Surface*) malloc(sizeof *s);
    s->flags = 0; s->format = NULL; s->w = w; s->h = h; s->pitch = 4 * w;
    s->pixels = malloc((size_t) 4 * w * h);
    uint32_t* px = (uint32_t*) s->pixels;
    fseek(f, (long) off, SEEK_SET);
    for(int r = 0; r < h; r++)
    {
        if(fread(row, 1, rowbytes, f) != rowbytes)
        { snprintf(img_error, sizeof img_error, "Truncated BMP: %s", path); fclose(f); free(row); SDL_FreeSurface(s); return NULL; }
        const int y = hs < 0 ? r : h - 1 - r;
        for(int x = 0; x < w; x++)
            px[(size_t) y * w + x] = (uint32_t) row[3 * x + 2] << 16 | (uint32_t) row[3 * x + 1] << 8 | row[3 * x];
    }
    free(row);
    fclose(f);
    return s;
}

const char* IMG_GetError(void) { return img_error; }
