// stand-in for <boost/math/tools/config.hpp> (oracle build only)
#pragma once
#include <boost/config.hpp>
namespace boost { using uintmax_t = std::uintmax_t; }
