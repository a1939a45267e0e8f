"""realhf_b200 — a B200-native (sm_100a) RLHF training framework.

Same capabilities as openpsi-project/ReaLHF (dataflow graphs of model function calls,
3D parallelism, parameter reallocation, CUDA-graph generation, SFT/RW/DPO/PPO/GRPO), built
from scratch around hand-written sm_100a kernels, one process per GPU and NCCL/NVLink peer
memory.  See DESIGN.md for the architecture and SURVEY.md for the parity map.
"""

__version__ = "0.1.0"

# Public names, resolved on first use so that `import realhf_b200` stays light (the reference re-exports the same names from
# its package root, realhf/__init__.py:4-36: user code written as `from realhf import PPOConfig, MFCDef, ...` ports by changing
# the package name only).
_EXPORTS = {
    "ModelFamily": "api.config", "ModelName": "api.config", "ModelShardID": "api.config",
    "SequenceSample": "api.data",
    "MFCDef": "api.dfg",
    "FinetuneSpec": "api.model", "GenerationHyperparameters": "api.model", "Model": "api.model", "ModelBackend": "api.model",
    "ModelInterface": "api.model", "ModelVersion": "api.model", "PipelinableEngine": "api.model", "ReaLModelConfig": "api.model",
    "PairedComparisonDatasetConfig": "api.quickstart", "PromptAnswerDatasetConfig": "api.quickstart",
    "PromptOnlyDatasetConfig": "api.quickstart", "MFCConfig": "api.quickstart", "ModelTrainEvalConfig": "api.quickstart",
    "OptimizerConfig": "api.quickstart", "ParallelismConfig": "api.quickstart",
    "CommonExperimentConfig": "experiments.common", "ExperimentSaveEvalControl": "api.system",
    "DPOConfig": "experiments.algos", "GenerationConfig": "experiments.algos", "PPOConfig": "experiments.algos",
    "PPOHyperparameters": "experiments.algos", "RWConfig": "experiments.algos", "SFTConfig": "experiments.algos",
}
__all__ = sorted(_EXPORTS) + ["__version__"]


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module 'realhf_b200' has no attribute '{name}'")
    import importlib
    value = getattr(importlib.import_module(f"realhf_b200.{mod}"), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(set(globals()) | set(_EXPORTS))
