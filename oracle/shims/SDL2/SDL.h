#pragma once
#include "../SDL.h"
