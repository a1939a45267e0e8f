"""Best-effort import probe of the UNMODIFIED reference installed under baseline/_ref (see DESIGN.md "Reference arm").

    python baseline/probe_reference.py

Adds baseline/shims (pure-Python stand-ins for packages that are only used for logging / CLI sugar) to sys.path and imports the
modules the reference's PPO quickstart needs, one stage at a time.  Every package that is missing from this offline image is
recorded and replaced by an empty stub so that the probe can continue and list ALL blockers, and every attribute the reference
then asks of such a stub is recorded too -- this is the exact list of symbols a shim would have to implement with real
behaviour (for megatron.core / deepspeed / flash_attn internals that is the distributed optimizer, the pipeline engine and the
fused kernels themselves: not shimmable).  Prints one JSON object.
"""
import importlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.path.join(HERE, "_ref"))

STAGES = ["realhf.base.logging", "realhf.api.core.config", "realhf.api.core.dfg", "realhf.api.quickstart.entrypoint",
          "realhf.experiments.common.ppo_exp", "realhf.apps.quickstart", "realhf.impl.model.nn.real_llm_api",
          "realhf.impl.model.backend.megatron", "realhf.impl.model.backend.deepspeed", "realhf.impl.model.interface.ppo_interface",
          "realhf.system.master_worker", "realhf.system.model_worker"]
missing, asked = [], {}


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        asked.setdefault(self.__name__, set()).add(k)
        sub = _Stub(f"{self.__name__}.{k}")
        sub.__path__ = []
        sys.modules[sub.__name__] = sub
        return sub

    def __call__(self, *a, **k):
        return self

    def __mro_entries__(self, bases):
        return (object,)


def main():
    report = {}
    for stage in STAGES:
        for _ in range(60):
            try:
                importlib.import_module(stage)
                report[stage] = "imported"
                break
            except ModuleNotFoundError as e:
                if e.name is None or e.name.startswith("realhf"):
                    report[stage] = f"ModuleNotFoundError: {e}"
                    break
                missing.append(e.name)
                stub = _Stub(e.name)
                stub.__path__ = []
                sys.modules[e.name] = stub
                for k in [k for k in sys.modules if k.startswith("realhf")]:
                    del sys.modules[k]
            except Exception as e:  # a stubbed symbol was USED at import time, or python 3.12 incompatibility
                report[stage] = f"{type(e).__name__}: {str(e)[:200]}"
                break
    print(json.dumps(dict(python=sys.version.split()[0], stages=report, missing_packages=sorted(set(missing)),
                          symbols_asked_of_stubs={k: sorted(v) for k, v in asked.items()}), indent=1))


if __name__ == "__main__":
    main()
