"""Small reusable utilities: CUDA-graph registry, logits warpers, the global stats tracker."""
