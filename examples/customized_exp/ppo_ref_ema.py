"""PPO whose reference model tracks the actor by an exponential moving average.

Parity: reference `examples/customized_exp/ppo_ref_ema.py` — a one-way `ParamReallocHook(target=ref, eta)` after
`actor_train` performs `ref <- eta * actor + (1 - eta) * ref` inside the parameter-reallocation kernel (the EMA mode of
`ops/csrc/segcopy.cu`), whatever the two models' layouts are.

    python examples/customized_exp/ppo_ref_ema.py ppo-ref-ema experiment_name=ema trial_name=t0 ref_ema_eta=0.001 actor.path=...
"""

import dataclasses

from realhf_b200.api.config import ModelName
from realhf_b200.api.dfg import ParamReallocHook
from realhf_b200.api.quickstart import register_quickstart_exp
from realhf_b200.experiments.algos import PPOConfig


@dataclasses.dataclass
class PPORefEMAConfig(PPOConfig):
    ref_ema_eta: float = 0.001

    @property
    def rpcs(self):
        rpcs = super().rpcs
        rpcs["actor_train"].add_post_hook(ParamReallocHook(target=ModelName("ref", 0), eta=self.ref_ema_eta))
        return rpcs


register_quickstart_exp("ppo-ref-ema", PPORefEMAConfig)

if __name__ == "__main__":
    import os
    import sys

    os.environ["REAL_USER_CODE"] = os.path.abspath(__file__)
    from realhf_b200.apps.quickstart import main
    main(sys.argv[1:])
