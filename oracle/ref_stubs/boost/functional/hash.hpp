#pragma once
#include "ref_stub_prelude.h"
