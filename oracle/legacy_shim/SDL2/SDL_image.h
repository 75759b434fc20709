// oracle/legacy_shim -- TEST INFRASTRUCTURE: declarations only (see SDL2/SDL.h).
#pragma once
#include "SDL.h"
SDL_Surface* IMG_Load(const char*);
const char* IMG_GetError();
enum { IMG_INIT_JPG = 1, IMG_INIT_PNG = 2 };
int IMG_Init(int);
void IMG_Quit();
