/* TEST INFRASTRUCTURE (oracle).  Minimal stand-in for <SDL2/SDL.h>, written from scratch:
 * just the names the reference program (main.c) mentions, so that it compiles UNMODIFIED in a
 * container that has no SDL2.  The semantics live in ../shim.c.  Nothing here is SDL code. */
#ifndef GEL_ORACLE_STUB_SDL_H
#define GEL_ORACLE_STUB_SDL_H
#include <stdint.h>

typedef struct SDL_Window SDL_Window;
typedef struct SDL_Renderer SDL_Renderer;
typedef struct SDL_Texture SDL_Texture;
typedef struct SDL_PixelFormat SDL_PixelFormat;
typedef struct SDL_Point SDL_Point;

typedef struct { uint32_t flags; SDL_PixelFormat* format; int w, h, pitch; void* pixels; } SDL_Surface;
typedef struct { int x, y, w, h; } SDL_Rect;
typedef struct { uint32_t type; } SDL_Event;
typedef enum { SDL_FALSE = 0, SDL_TRUE = 1 } SDL_bool;

enum { SDL_QUIT = 0x100 };
enum { SDL_INIT_VIDEO = 0x20 };
enum { SDL_PIXELFORMAT_ARGB8888 = 1, SDL_PIXELFORMAT_RGB888 = 2 };
enum { SDL_TEXTUREACCESS_STREAMING = 1 };
enum { SDL_FLIP_NONE = 0 };

int SDL_Init(uint32_t flags);
int SDL_CreateWindowAndRenderer(int w, int h, uint32_t flags, SDL_Window** win, SDL_Renderer** ren);
void SDL_SetWindowTitle(SDL_Window* win, const char* title);
SDL_Texture* SDL_CreateTexture(SDL_Renderer* ren, uint32_t format, int access, int w, int h);
int SDL_LockTexture(SDL_Texture* tex, const SDL_Rect* rect, void** pixels, int* pitch);
void SDL_UnlockTexture(SDL_Texture* tex);
int SDL_RenderCopyEx(SDL_Renderer* ren, SDL_Texture* tex, const SDL_Rect* src, const SDL_Rect* dst,
                     double angle, const SDL_Point* center, int flip);
void SDL_RenderPresent(SDL_Renderer* ren);
int SDL_PollEvent(SDL_Event* event);
uint32_t SDL_GetRelativeMouseState(int* x, int* y);
int SDL_SetRelativeMouseMode(SDL_bool enabled);
uint32_t SDL_GetTicks(void);
void SDL_Delay(uint32_t ms);
SDL_PixelFormat* SDL_AllocFormat(uint32_t format);
SDL_Surface* SDL_ConvertSurface(SDL_Surface* src, const SDL_PixelFormat* fmt, uint32_t flags);
void SDL_FreeFormat(SDL_PixelFormat* fmt);
void SDL_FreeSurface(SDL_Surface* s);
#endif
