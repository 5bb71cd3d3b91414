// Stand-in: modules/io/utils.cpp includes Boost's SHA-1 without using it (test infrastructure).
#pragma once
