#pragma once
#include <boost/icl/interval_set.hpp>
