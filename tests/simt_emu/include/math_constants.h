#pragma once
#include <limits>
#define CUDART_NAN (std::numeric_limits<double>::quiet_NaN())
#define CUDART_INF (std::numeric_limits<double>::infinity())
