// Drop-in shadow of the reference's include/ligetron/webgpu/buffer_view.hpp (see buffer_binding.hpp beside it).
#pragma once
#include "buffer_binding.hpp"
