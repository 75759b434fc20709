// oracle/legacy_shim/SDL2/SDL.h -- TEST INFRASTRUCTURE.  Declarations (no behaviour) of the SDL2 names that the reference's legacy
// demo sources mention, so that hello-shs-renderer/shs_renderer.hpp and hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp
// compile where they lie under /root/reference.  Nothing here is ever called by the harness: the raster functions it pins
// (RendererSystem::draw_triangle_tile, the Blinn-Phong shaders, Canvas, ZBuffer, Camera3D) do not touch SDL.  SDL2 itself is a
// third-party dependency of the reference that is absent from this container.
#pragma once
#include <cstdint>
typedef uint8_t Uint8;
typedef uint32_t Uint32;
typedef int32_t Sint32;
struct SDL_PixelFormat { Uint32 format; Uint8 BytesPerPixel; };
struct SDL_Surface { Uint32 flags; SDL_PixelFormat* format; int w, h, pitch; void* pixels; };
struct SDL_Window;
struct SDL_Renderer;
struct SDL_Texture;
struct SDL_Rect { int x, y, w, h; };
struct SDL_Keysym { int sym; };
struct SDL_KeyboardEvent { Uint32 type; SDL_Keysym keysym; };
struct SDL_MouseButtonEvent { Uint32 type; Uint8 button; Sint32 x, y; };
struct SDL_MouseMotionEvent { Uint32 type; Sint32 x, y, xrel, yrel; };
union SDL_Event { Uint32 type; SDL_KeyboardEvent key; SDL_MouseButtonEvent button; SDL_MouseMotionEvent motion; };
enum { SDL_QUIT = 0x100, SDL_KEYDOWN = 0x300, SDL_KEYUP, SDL_MOUSEMOTION = 0x400, SDL_MOUSEBUTTONDOWN, SDL_MOUSEBUTTONUP };
enum { SDLK_ESCAPE = 27, SDLK_a = 'a', SDLK_d = 'd', SDLK_s = 's', SDLK_w = 'w' };
#define SDL_BUTTON_LEFT 1
#define SDL_INIT_VIDEO 0x20u
#define SDL_BIG_ENDIAN 4321
#define SDL_LIL_ENDIAN 1234
#define SDL_BYTEORDER SDL_LIL_ENDIAN
#define SDL_PIXELFORMAT_RGBA32 0x16762004u
int SDL_Init(Uint32);
void SDL_Quit();
Uint32 SDL_GetTicks();
int SDL_PollEvent(SDL_Event*);
int SDL_CreateWindowAndRenderer(int, int, Uint32, SDL_Window**, SDL_Renderer**);
void SDL_DestroyWindow(SDL_Window*);
void SDL_DestroyRenderer(SDL_Renderer*);
void SDL_DestroyTexture(SDL_Texture*);
SDL_Texture* SDL_CreateTextureFromSurface(SDL_Renderer*, SDL_Surface*);
int SDL_UpdateTexture(SDL_Texture*, const SDL_Rect*, const void*, int);
int SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, const SDL_Rect*, const SDL_Rect*);
void SDL_RenderPresent(SDL_Renderer*);
int SDL_RenderClear(SDL_Renderer*);
void SDL_SetWindowTitle(SDL_Window*, const char*);
SDL_Surface* SDL_ConvertSurfaceFormat(SDL_Surface*, Uint32, Uint32);
SDL_Surface* SDL_CreateRGBSurface(Uint32, int, int, int, Uint32, Uint32, Uint32, Uint32);
void SDL_FreeSurface(SDL_Surface*);
const char* SDL_GetError();
void SDL_GetRGBA(Uint32, const SDL_PixelFormat*, Uint8*, Uint8*, Uint8*, Uint8*);
Uint32 SDL_MapRGBA(const SDL_PixelFormat*, Uint8, Uint8, Uint8, Uint8);
