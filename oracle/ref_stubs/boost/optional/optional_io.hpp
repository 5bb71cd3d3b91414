#pragma once
#include "boost/optional.hpp"
