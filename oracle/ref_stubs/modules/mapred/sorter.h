// Shadows modules/mapred/sorter.h (test infrastructure; map-reduce runtime, outside the hot path).
#pragma once
#include "modules/mapred/manifest.h"
