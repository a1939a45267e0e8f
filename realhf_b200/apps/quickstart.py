"""`python -m realhf_b200.apps.quickstart <sft|rw|dpo|ppo|gen|profile> key=value ...`

Parity: `realhf/apps/quickstart.py` + `api/quickstart/entrypoint.py`: the first argument selects a registered
experiment dataclass, the remaining `a.b.c=value` arguments override its (nested) fields.
"""

from __future__ import annotations

import sys

from realhf_b200.api.quickstart import QUICKSTART_EXPERIMENTS, parse_overrides


def build_experiment(argv):
    import realhf_b200.experiments.algos  # noqa: F401  (registers sft / rw / dpo / ppo / gen)
    import realhf_b200.experiments.profile  # noqa: F401  (registers profile)
    if not argv or argv[0] in ("-h", "--help"):
        print("usage: python -m realhf_b200.apps.quickstart {" + ",".join(sorted(QUICKSTART_EXPERIMENTS)) + "} key=value ...")
        sys.exit(0)
    name, overrides = argv[0], argv[1:]
    if name not in QUICKSTART_EXPERIMENTS:
        raise SystemExit(f"unknown experiment `{name}`; choices: {sorted(QUICKSTART_EXPERIMENTS)}")
    cfg = QUICKSTART_EXPERIMENTS[name]()
    parse_overrides(cfg, overrides)
    for n in (cfg.experiment_name, cfg.trial_name):
        if "_" in n:
            raise SystemExit(f"experiment_name / trial_name must not contain `_` (got `{n}`)")
    return cfg


def main(argv=None):
    from realhf_b200.apps.main import main_start
    cfg = build_experiment(sys.argv[1:] if argv is None else argv)
    if hasattr(cfg, "run_local"):  # in-process experiments (profile sweeps)
        return cfg.run_local()
    return main_start(cfg)


if __name__ == "__main__":
    main()
