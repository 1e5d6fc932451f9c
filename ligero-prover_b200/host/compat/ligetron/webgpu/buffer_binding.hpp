// Drop-in shadow of the reference's include/ligetron/webgpu/buffer_binding.hpp: put this directory BEFORE the
// reference's include/ on the include path and the stage contexts' `webgpu::buffer_binding` members
// (include/zkp/nonbatch_context.hpp:578-580,863-871,1065-1066) become the CUDA executor's bindings -- no edit to
// nonbatch_context.hpp, and no <webgpu/webgpu.h> (Dawn) in the translation unit.
#pragma once

#include "../../../cuda_executor.hpp"

namespace ligero {
namespace webgpu {

using buffer_view = ::ligero::cuda::buffer_view;
using buffer_binding = ::ligero::cuda::buffer_binding;
using eltwise_offset = ::ligero::cuda::eltwise_offset;
using device_uint256_t = ::ligero::cuda::device_uint256_t;

}  // namespace webgpu
}  // namespace ligero
