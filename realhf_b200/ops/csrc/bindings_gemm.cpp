#include <torch/library.h>
void register_gemm_ops(torch::Library& m) {}
