// oracle/legacy_shim -- TEST INFRASTRUCTURE: declarations only (see SDL2/SDL.h).
#pragma once
#include "SDL.h"
SDL_Surface* IMG_Load(const char*);
const char* IMG_GetError();
