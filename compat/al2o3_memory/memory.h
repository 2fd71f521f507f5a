// Compatibility shim: the encode path includes this header but uses nothing from it.
#pragma once
#include <stdlib.h>
