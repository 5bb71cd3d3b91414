// Shadows modules/io/make_unique.h (test infrastructure): the reference's own make_unique collides with
// std::make_unique under the C++17 this stand-in tree is compiled as.
#pragma once
#include <memory>
using std::make_unique;
