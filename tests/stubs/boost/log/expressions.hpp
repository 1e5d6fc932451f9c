#pragma once
#include <boost/log/trivial.hpp>
