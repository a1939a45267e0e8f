"""Entry point of worker processes (parity: `realhf/apps/remote.py`): load the pickled system config written by the
launcher, run the requested worker, publish status keys for the controller's liveness checks."""

from __future__ import annotations

import argparse
import os
import pickle
import sys
import traceback

from realhf_b200.base import constants, logging, name_resolve

logger = logging.getLogger("remote")


def config_path(exp: str, trial: str) -> str:
    return os.path.join(constants.run_dirs(exp, trial)["log"], "experiment_config.pkl")


def status_key(exp, trial, worker_type, index):
    return f"{exp}/{trial}/status/{worker_type}/{index}"


def control_key(exp, trial, worker_type, index):
    """Commands for a worker: `pause` / `resume` / `exit` (written by the controller, polled by the worker between steps)."""
    return f"{exp}/{trial}/control/{worker_type}/{index}"


def status_ttl() -> float:
    return float(os.environ.get("REAL_STATUS_TTL", "120"))


def _watch_controller(exp: str, trial: str, extra_keys=None):
    """Exit when the launcher's liveness key disappears (controller killed / its node lost): orphaned workers would otherwise
    hold their GPUs forever (reference: `watch_names`, system/worker_base.py:660-666).  `extra_keys`: the worker's
    `WorkerInformation.watch_keys` -- more name-resolve keys whose disappearance ends this worker."""
    ttl = status_ttl()
    keys = [status_key(exp, trial, "controller", 0)]
    if extra_keys:
        keys += [extra_keys] if isinstance(extra_keys, str) else list(extra_keys)

    def _die():
        logger.error(f"a watched liveness key expired ({keys}): exiting")
        os._exit(3)
    name_resolve.watch_names(keys, _die, poll_frequency=max(0.5, ttl / 3), wait_timeout=300)


def main_worker(args):
    # register everything the configs may name
    import realhf_b200.datasets  # noqa: F401
    import realhf_b200.engine.engine  # noqa: F401
    import realhf_b200.interfaces.basic  # noqa: F401
    import realhf_b200.interfaces.ppo  # noqa: F401
    import realhf_b200.models.factory  # noqa: F401
    user_code = os.environ.get("REAL_USER_CODE")
    if user_code:  # custom experiments / interfaces registered by the user's script
        from realhf_b200.base.importing import import_usercode
        import_usercode(user_code, "real_user_code")
    with open(config_path(args.experiment_name, args.trial_name), "rb") as f:
        cfg = pickle.load(f)
    key = status_key(args.experiment_name, args.trial_name, args.worker_type, args.jobstep_id)
    name_resolve.add(key, "RUNNING", replace=True, keepalive_ttl=status_ttl())
    if os.environ.get("REAL_WATCH_CONTROLLER", "1") == "1":
        wcfg = cfg.model_worker[args.jobstep_id] if args.worker_type == "model_worker" else cfg.master_worker[0]
        info = getattr(wcfg, "worker_info", None)
        _watch_controller(args.experiment_name, args.trial_name, getattr(info, "watch_keys", None))
    try:
        if args.worker_type == "model_worker":
            from realhf_b200.system.model_worker import ModelWorker
            ModelWorker(cfg.model_worker[args.jobstep_id]).run()
        elif args.worker_type == "master_worker":
            from realhf_b200.system.master_worker import MasterWorker
            MasterWorker(cfg.master_worker[0]).run()
        else:
            raise ValueError(args.worker_type)
        name_resolve.add(key, "COMPLETED", replace=True)
    except Exception:
        logger.error(traceback.format_exc())
        name_resolve.add(key, "ERROR", replace=True)
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser(prog="realhf_b200.apps.remote")
    sub = ap.add_subparsers(dest="cmd", required=True)
    w = sub.add_parser("worker")
    w.add_argument("-w", "--worker_type", required=True)
    w.add_argument("-e", "--experiment_name", required=True)
    w.add_argument("-f", "--trial_name", required=True)
    w.add_argument("-i", "--jobstep_id", type=int, required=True)
    w.add_argument("-g", "--n_jobsteps", type=int, default=1)
    w.add_argument("-r", "--worker_submission_index", type=int, default=0)
    w.add_argument("-p", "--wprocs_per_jobstep", type=int, default=1)
    w.add_argument("-j", "--wprocs_in_job", type=int, default=1)
    w.add_argument("-o", "--wproc_offset", type=int, default=0)
    args = ap.parse_args()
    main_worker(args)


if __name__ == "__main__":
    main()
