"""realhf_b200 — a B200-native (sm_100a) RLHF training framework.

Same capabilities as openpsi-project/ReaLHF (dataflow graphs of model function calls,
3D parallelism, parameter reallocation, CUDA-graph generation, SFT/RW/DPO/PPO/GRPO), built
from scratch around hand-written sm_100a kernels, one process per GPU and NCCL/NVLink peer
memory.  See DESIGN.md for the architecture and SURVEY.md for the parity map.
"""

__version__ = "0.1.0"
