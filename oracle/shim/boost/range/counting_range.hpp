// Stand-in for a Boost header the reference includes but does not use on the hot path
// (reference src/support/cache.hh:25, src/support/int_part.cc:18-19, src/support/util.hh:27).
// It only supplies the std headers the reference picks up transitively.  Oracle build only.
#pragma once
#include <limits>
#include <tuple>
#include <cassert>
#include <algorithm>
#include <cmath>
#include <cstdlib>
namespace boost {}
