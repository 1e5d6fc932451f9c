// Stand-in for Boost.Log (absent from this image): the reference only uses BOOST_LOG_TRIVIAL(level) << ...
// (include/util/log.hpp:29-38).  Test infrastructure for tests/test_boundary_compile.py, nothing ships from here.
#pragma once
#include <iostream>
namespace boost { namespace log { namespace trivial { enum severity_level { trace, debug, info, warning, error, fatal }; } } }
#define BOOST_LOG_TRIVIAL(lvl) std::clog
