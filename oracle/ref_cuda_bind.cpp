// oracle/ref_cuda_bind.cpp — TEST INFRASTRUCTURE, not product code.
//
// Binding list for the "reference CUDA build": the reference's own hot-path
// translation units (csrc/attention/attention_kernels.cu, cache_kernels.cu,
// layernorm_kernels.cu, pos_encoding_kernels.cu, activation_kernels.cu,
// cuda_utils_kernels.cu, custom_all_reduce.cu) are compiled UNMODIFIED, from
// where they lie under /root/reference, for sm_100a by oracle/Makefile.  The
// reference's csrc/pybind.cpp cannot be linked as-is because it also binds the
// quantisation / MoE ops whose TUs are out of scope, so this
// file re-declares the module with only the hot-path names.  The function
// declarations come from the reference headers (csrc/ops.h, csrc/cache.h,
// csrc/cuda_utils.h) via -I/root/reference/csrc; nothing is copied.
#include "cache.h"
#include "cuda_utils.h"
#include "ops.h"
#include <torch/extension.h>

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  auto ops = m.def_submodule("ops");
  ops.def("paged_attention_v1", &paged_attention_v1);
  ops.def("paged_attention_v2", &paged_attention_v2);
  ops.def("silu_and_mul", &silu_and_mul);
  ops.def("gelu_and_mul", &gelu_and_mul);
  ops.def("gelu_tanh_and_mul", &gelu_tanh_and_mul);
  ops.def("gelu_new", &gelu_new);
  ops.def("gelu_fast", &gelu_fast);
  ops.def("rms_norm", &rms_norm);
  ops.def("fused_add_rms_norm", &fused_add_rms_norm);
  ops.def("rotary_embedding", &rotary_embedding);
  ops.def("batched_rotary_embedding", &batched_rotary_embedding);

  auto cache_ops = m.def_submodule("cache_ops");
  cache_ops.def("swap_blocks", &swap_blocks);
  cache_ops.def("copy_blocks", &copy_blocks);
  cache_ops.def("reshape_and_cache", &reshape_and_cache);
  cache_ops.def("convert_fp8", &convert_fp8);

  auto cuda_utils = m.def_submodule("cuda_utils");
  cuda_utils.def("get_device_attribute", &get_device_attribute);
  cuda_utils.def("get_max_shared_memory_per_block_device_attribute",
                 &get_max_shared_memory_per_block_device_attribute);

  auto custom_ar = m.def_submodule("custom_ar");
  custom_ar.def("init_custom_ar", &init_custom_ar);
  custom_ar.def("should_custom_ar", &should_custom_ar);
  custom_ar.def("all_reduce_reg", &all_reduce_reg);
  custom_ar.def("all_reduce_unreg", &all_reduce_unreg);
  custom_ar.def("dispose", &dispose);
  custom_ar.def("meta_size", &meta_size);
  custom_ar.def("register_buffer", &register_buffer);
  custom_ar.def("get_graph_buffer_ipc_meta", &get_graph_buffer_ipc_meta);
  custom_ar.def("register_graph_buffers", &register_graph_buffers);
}
