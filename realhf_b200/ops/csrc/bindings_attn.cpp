#include <torch/library.h>
void register_attn_ops(torch::Library& m) {}
