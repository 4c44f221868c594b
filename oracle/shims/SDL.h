// TEST INFRASTRUCTURE ONLY — the four SDL_Surface fields swegl's renderer touches
// (pixels, pitch, w, format->BytesPerPixel; viewport.cpp:61,88-103, renderer.cpp:483).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cmath>
struct SDL_PixelFormat { uint8_t BytesPerPixel; };
struct SDL_Surface { int w, h, pitch; void * pixels; SDL_PixelFormat * format; };
