"""`python -m realhf_b200.apps.quickstart <sft|rw|dpo|ppo|gen|profile> key=value ...`

Parity: `realhf/apps/quickstart.py` + `api/quickstart/entrypoint.py`: the first argument selects a registered
experiment dataclass, the remaining `a.b.c=value` arguments override its (nested) fields.
"""

from __future__ import annotations

import sys

from realhf_b200.api.quickstart import QUICKSTART_EXPERIMENTS, parse_overrides


def option_table(cls, prefix=""):
    """(dotted option, type, default) for every leaf field of an experiment dataclass (also feeds docs/expconfig.md)."""
    import dataclasses
    rows = []
    for f in dataclasses.fields(cls):
        if f.name.startswith("_"):
            continue
        default = f.default if f.default is not dataclasses.MISSING else (
            f.default_factory() if f.default_factory is not dataclasses.MISSING else None)
        if dataclasses.is_dataclass(default):
            rows.extend(option_table(type(default), prefix + f.name + "."))
        else:
            t = f.type if isinstance(f.type, str) else (str(f.type).replace("typing.", "") if "[" in str(f.type) else getattr(f.type, "__name__", str(f.type)))
            rows.append((prefix + f.name, t, repr(default)))
    return rows


def _revalidate(obj, path=""):
    """Overrides are applied with setattr, after the dataclasses validated their defaults: run the generation options'
    checks again now, so that e.g. `max_new_tokens=6` with the default `min_new_tokens=256` fails here, on the command line,
    and not inside a worker minutes later."""
    import dataclasses
    from realhf_b200.api.model import GenerationHyperparameters
    if isinstance(obj, GenerationHyperparameters):
        try:
            obj.__post_init__()
        except ValueError as e:
            raise SystemExit(f"invalid generation options at `{path or 'gen'}`: {e}")
    elif dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        for f in dataclasses.fields(obj):
            _revalidate(getattr(obj, f.name, None), f"{path}.{f.name}" if path else f.name)


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
    if any(o in ("-h", "--help", "help") for o in overrides):
        print(f"usage: python -m realhf_b200.apps.quickstart {name} key=value ...\n\noptions (dotted name, type, default):")
        for opt, typ, default in option_table(type(cfg)):
            print(f"  {opt:<55} {typ:<28} {default}")
        sys.exit(0)
    parse_overrides(cfg, overrides)
    _revalidate(cfg)
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
