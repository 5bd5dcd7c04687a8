/* TEST INFRASTRUCTURE (oracle).  Stand-in for <SDL2/SDL_image.h>; see SDL.h beside it. */
#ifndef GEL_ORACLE_STUB_SDL_IMAGE_H
#define GEL_ORACLE_STUB_SDL_IMAGE_H
#include "SDL.h"
SDL_Surface* IMG_Load(const char* path);
const char* IMG_GetError(void);
#endif
