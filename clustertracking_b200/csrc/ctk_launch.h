// ctk_launch.h -- declaration of the per-configuration launcher (defined in ctk_kernel.cuh,
// explicitly instantiated per (arithmetic, family) in ctk_inst.cu so the instances build in parallel).
#pragma once
#include <cuda_runtime.h>

#include "ctk_layout.h"

namespace ctk {
template <class C>
int launch_refine(const BatchArgs& args, cudaStream_t stream, char* err, size_t err_len);
// one pass of a global-level fit (ctk_global_pass); instantiated for the full flavour only
template <class C>
int launch_global(const BatchArgs& args, cudaStream_t stream, char* err, size_t err_len);
}
