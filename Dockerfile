# GPU image for realhf_b200 (model workers and the master worker use the same image; see examples/cluster_config.json).
# Needs an NGC PyTorch base with CUDA >= 12.8 (nvcc with sm_100a) -- not built or tested in the development sandbox (no network).
ARG BASE=nvcr.io/nvidia/pytorch:25.03-py3
FROM ${BASE}

ENV DEBIAN_FRONTEND=noninteractive PIP_NO_CACHE_DIR=1
RUN apt-get update && apt-get install -y --no-install-recommends git build-essential ninja-build && rm -rf /var/lib/apt/lists/*

WORKDIR /opt/realhf_b200
COPY . .
# flash-attn is optional (packed varlen attention until the in-tree tcgen05 kernels are the default)
RUN pip install networkx pyzmq psutil transformers pybind11 tensorboard && \
    python -m realhf_b200.ops.build --sass && \
    pip install -e . --no-build-isolation --no-deps

ENV PYTHONPATH=/opt/realhf_b200 REAL_FILEROOT=/workspace/realhf_b200
CMD ["python", "-m", "realhf_b200.apps.quickstart"]
