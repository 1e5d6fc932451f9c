// Drop-in shadow of the reference's include/wgpu.hpp: with this directory BEFORE the reference's include/ on the include
// path, `#include <wgpu.hpp>` (src/webgpu_prover.cpp:28, src/webgpu_verifier.cpp, include/host_modules/vbn254fr.hpp:25)
// yields the CUDA executor under the reference's own name, so `using executor_t = webgpu_context;`
// (src/webgpu_prover.cpp:54) needs no edit either.  Link -llgr instead of dawn::webgpu_dawn.
#pragma once

#include "../cuda_executor.hpp"
#include "ligetron/webgpu/buffer_binding.hpp"

namespace ligero {

using webgpu_context = cuda_context;

}  // namespace ligero
