"""PPO with an arbitrary external reward function instead of a reward *model*.

Parity: reference `examples/customized_exp/ppo_sentiment.py`: the reward MFC runs on model type `tokenizer` with the
`null` backend, and the interface calls user code (here: a length / keyword heuristic; plug a HF classifier in `score`).
"""

import dataclasses
from typing import Optional

import torch

from realhf_b200.api.config import ModelAbstraction, ModelBackendAbstraction, ModelInterfaceAbstraction
from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import Model, ModelInterface, register_interface
from realhf_b200.api.quickstart import register_quickstart_exp
from realhf_b200.experiments.algos import PPOConfig


@dataclasses.dataclass
class HeuristicRewardInterface(ModelInterface):
    target_len: int = 64

    def score(self, text: str, n_tokens: int) -> float:
        return 1.0 - abs(n_tokens - self.target_len) / self.target_len

    @torch.no_grad()
    def inference(self, model: Model, data: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        ids = data.data["packed_input_ids"]
        lens = data.flat_seqlens("packed_input_ids")
        off, scores = 0, []
        for l in lens:
            toks = ids[off:off + l].tolist()
            text = model.tokenizer.decode(toks, skip_special_tokens=True) if model.tokenizer is not None else ""
            scores.append(self.score(text, l))
            off += l
        return SequenceSample.from_default(ids=data.ids, seqlens=lens, data=dict(rewards=torch.tensor(scores, device=ids.device)))


register_interface("heuristic_reward", HeuristicRewardInterface)


@dataclasses.dataclass
class PPOExternalRewardConfig(PPOConfig):
    @property
    def rpcs(self):
        rpcs = super().rpcs
        rpcs["rew_inf"].interface_impl = ModelInterfaceAbstraction("heuristic_reward")
        return rpcs


register_quickstart_exp("ppo-external-reward", PPOExternalRewardConfig)

if __name__ == "__main__":
    import os
    import sys

    os.environ["REAL_USER_CODE"] = os.path.abspath(__file__)   # workers re-import this file before building anything
    from realhf_b200.apps.quickstart import main
    main(sys.argv[1:])
