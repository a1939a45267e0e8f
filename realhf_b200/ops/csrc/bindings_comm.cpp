#include <torch/library.h>
void register_comm_ops(torch::Library& m) {}
