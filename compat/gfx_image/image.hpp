#pragma once
#include "gfx_image/image.h"
