"""``vllm._C.custom_ar`` — tensor-parallel all-reduce surface (csrc/pybind.cpp:108-123,
csrc/ops.h:143-160).  Filled in by vllm_ltr_b200.parallel (TP path); see DESIGN.md §multi-GPU."""
