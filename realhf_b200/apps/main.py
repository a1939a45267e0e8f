"""Launcher (parity: `realhf/apps/main.py` main_start :74-230 and `system/controller.py`): resolve the experiment into
system configs, start the workers through a scheduler client, watch them, and re-enter on failure when
`recover_mode` asks for it."""

from __future__ import annotations

import os
import pickle
import signal
import time
from typing import Optional

from realhf_b200.api.system import Experiment
from realhf_b200.apps.remote import config_path, control_key, status_key, status_ttl
from realhf_b200.base import constants, logging, name_resolve
from realhf_b200.scheduler import client as sched_client

logger = logging.getLogger("main", "system")


def _command(exp: str, trial: str, cmd: str):
    """Deliver `pause` / `resume` / `exit` to the master worker: through the control key it polls between steps (works from any
    machine that sees the name store), and through its RPC endpoint when that is reachable (immediate acknowledgement)."""
    name_resolve.add(control_key(exp, trial, "master_worker", 0), cmd, replace=True)
    try:
        from realhf_b200.system.worker_control import WorkerControlPanel
        panel = WorkerControlPanel(exp, trial, timeout=2.0)
        panel.connect(["master_worker/0"], timeout=0.5)
        ack = panel.request("master_worker/0", cmd)
        panel.close()
        return ack
    except Exception:  # no endpoint published (yet) / not reachable from here: the control key does the job
        return None


def worker_panel_status(exp: str, trial: str):
    """{worker: {status, uptime_s, served, progress}} through the workers' RPC endpoints (`apps.main status --rpc`)."""
    from realhf_b200.system.worker_control import WorkerControlPanel
    panel = WorkerControlPanel(exp, trial, timeout=5.0)
    names = panel.connect(timeout=1.0)
    st = panel.group_request("status")
    pr = panel.group_request("progress")
    panel.close()
    return {n: dict(st[n] if isinstance(st[n], dict) else dict(status="LOST"), progress=pr[n] if isinstance(pr[n], dict) else None) for n in names}


def pause_experiment(exp: str, trial: str):
    """The master finishes its current step, publishes PAUSED and stops issuing MFCs until `resume_experiment`; model workers
    idle on their request streams (parity: WorkerControlPanel pause / resume, system/worker_base.py:217-455)."""
    _command(exp, trial, "pause")


def resume_experiment(exp: str, trial: str):
    _command(exp, trial, "resume")


def stop_experiment(exp: str, trial: str):
    """Graceful stop after the current step: recover states are saved if the run was launched with recover_mode save / auto."""
    _command(exp, trial, "exit")


class Controller:
    """Polls worker status keys; raises on ERROR (reference: Controller.start polling loop, controller.py:275-318)."""

    def __init__(self, exp: str, trial: str, sched: sched_client.SchedulerClient, n_model_workers: int,
                 ignore_worker_error: bool = False):
        self.exp, self.trial, self.sched, self.n = exp, trial, sched, n_model_workers
        # debugging aid of the reference launcher (`--ignore_worker_error`, apps/main.py:375): a failing MODEL worker is reported but
        # does not end the trial, so that the surviving processes can be inspected; a failing master always ends it
        self.ignore_worker_error = ignore_worker_error
        self._reported = set()
        self._seen = set()
        self._lost_since = {}
        # liveness lease of the launcher itself: workers exit when it expires (apps/remote.py::_watch_controller)
        name_resolve.add(status_key(exp, trial, "controller", 0), "RUNNING", replace=True, keepalive_ttl=status_ttl())

    def pause(self):
        pause_experiment(self.exp, self.trial)

    def resume(self):
        resume_experiment(self.exp, self.trial)

    def stop(self):
        stop_experiment(self.exp, self.trial)

    def statuses(self):
        out = {}
        for wt, cnt in (("master_worker", 1), ("model_worker", self.n)):
            for i in range(cnt):
                k = f"{wt}/{i}"
                try:
                    out[k] = name_resolve.get(status_key(self.exp, self.trial, wt, i))
                    self._seen.add(k)
                except name_resolve.NameEntryNotFoundError:
                    # a status that was published and whose lease then ran out: the worker (or its node) died without a word
                    out[k] = "LOST" if k in self._seen else "UNKNOWN"
        return out

    def _filter_ignored(self, bad: dict) -> dict:
        if not self.ignore_worker_error:
            return bad
        for k, v in bad.items():
            if not k.startswith("master_worker") and k not in self._reported:
                self._reported.add(k)
                logger.error(f"{k} is {v}; continuing because the trial was started with ignore_worker_error")
        return {k: v for k, v in bad.items() if k.startswith("master_worker")}

    def wait(self, timeout: Optional[float] = None, poll: float = 0.5, status_poll: float = 5.0):
        t0 = time.monotonic()
        last_status = t0
        while True:
            if time.monotonic() - last_status > status_poll:
                # a worker that caught its own exception publishes ERROR before the scheduler sees the process exit
                last_status = time.monotonic()
                st = self.statuses()
                now = time.monotonic()
                for k, v in st.items():   # LOST must persist for a full extra lease before it counts: a worker that holds the
                    if v == "LOST":       # GIL through a long host-side step (unpickling a big checkpoint) misses touches too
                        self._lost_since.setdefault(k, now)
                    else:
                        self._lost_since.pop(k, None)
                bad = {k: v for k, v in st.items() if v == "ERROR" or (v == "LOST" and now - self._lost_since[k] >= status_ttl())}
                bad = self._filter_ignored(bad)
                if bad:
                    raise sched_client.JobException(self.sched.run_name, sorted(bad)[0], "localhost", sched_client.JobState.FAILED)
            infos = self.sched.find_all()
            master = next((i for i in infos if i.name.startswith("master_worker")), None)
            failed = [i for i in infos if i.state == sched_client.JobState.FAILED]
            failed = [i for i in failed if i.name in self._filter_ignored({i.name: "FAILED"})]
            if failed:
                raise sched_client.JobException(self.sched.run_name, failed[0].name, "localhost", failed[0].state)
            if master is not None and master.state == sched_client.JobState.COMPLETED:
                return
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError(f"experiment did not finish in {timeout}s; statuses: {self.statuses()}")
            time.sleep(poll)


def preflight(exp_cfg) -> None:
    """Cheap checks before any worker is launched: a typo in a path should cost seconds, not a scheduler round trip plus
    model loading on every GPU.  Local mode only (with Slurm the paths may exist on the compute nodes alone)."""
    problems = []
    for role, m in (getattr(exp_cfg, "models", None) or {}).items():
        path = getattr(m, "path", "")
        if path and not os.path.isdir(path):
            problems.append(f"{role}.path: `{path}` is not a directory")
        elif path and not os.path.exists(os.path.join(path, "config.json")):
            problems.append(f"{role}.path: `{path}` has no config.json (not a HuggingFace checkpoint directory)")
        elif not path and not getattr(m, "init_from_scratch", False):
            problems.append(f"{role}.path is empty and {role}.init_from_scratch is not set")
    for ds in list(getattr(exp_cfg, "datasets", None) or []) + list(getattr(exp_cfg, "eval_datasets", None) or []):
        path = (getattr(ds, "args", None) or {}).get("dataset_path")
        if path and not os.path.isfile(path):
            problems.append(f"dataset file `{path}` does not exist")
    if problems:
        raise FileNotFoundError("the experiment cannot start:\n  " + "\n  ".join(problems))


def main_start(exp_cfg: Experiment, recover_count: int = 0, timeout: Optional[float] = None, env_vars=None):
    exp, trial = exp_cfg.experiment_name, exp_cfg.trial_name
    if getattr(exp_cfg, "mode", "local") == "local" and recover_count == 0:
        preflight(exp_cfg)
    mode = getattr(exp_cfg, "mode", "local")
    recover_mode = getattr(exp_cfg, "recover_mode", "disabled")
    if mode == "local" and recover_mode == "auto":
        logger.warning("recover_mode=auto in local mode restarts all local processes on failure")
    os.environ["REAL_MODE"] = mode.upper()
    os.environ["REAL_RECOVER_RUN"] = "1" if (recover_mode == "resume" or recover_count > 0) else "0"
    os.environ["REAL_SAVE_RECOVER_STATES"] = "1" if recover_mode in ("auto", "save") else "0"
    name_resolve.clear_subtree(f"{exp}/{trial}")
    sys_cfg = exp_cfg.initial_setup()
    sched_cfg = exp_cfg.scheduling_setup()
    with open(config_path(exp, trial), "wb") as f:
        pickle.dump(sys_cfg, f)
    sched = sched_client.make(mode, exp, trial, **({"partition": getattr(exp_cfg, "partition", None)} if mode == "slurm" else {}))
    env = {k: os.environ[k] for k in constants.FORWARDED_ENV if k in os.environ}
    if getattr(exp_cfg, "wandb_mode", "disabled") != "disabled":
        env["WANDB_MODE"] = exp_cfg.wandb_mode
    if getattr(exp_cfg, "tensorboard", False):
        env["REAL_TENSORBOARD"] = "1"
    env.update(env_vars or {})
    debug = getattr(exp_cfg, "debug", True)
    def res(s):  # scheduler resources of a task group (the local scheduler ignores everything but `gpu`)
        return dict(cpu=s.cpu, gpu=s.gpu, mem=s.mem, nodelist=s.nodelist, exclude=s.exclude, time_limit=s.time_limit,
                    container_image=s.container_image, env_vars={**env, **s.env_vars})
    sched.submit_array("master_worker", sched_client.remote_worker_cmd(exp, trial, debug, "master_worker"), count=1,
                       **res(sched_cfg.master_worker.scheduling))
    mw = sched_cfg.model_worker
    sched.submit_array("model_worker", sched_client.remote_worker_cmd(exp, trial, debug, "model_worker"), count=mw.count,
                       **res(mw.scheduling))
    ctl = Controller(exp, trial, sched, mw.count, ignore_worker_error=bool(getattr(exp_cfg, "ignore_worker_error", False)))
    try:
        ctl.wait(timeout=timeout)
        sched.wait(timeout=60)
    except (KeyboardInterrupt, sched_client.JobException, TimeoutError) as e:
        # give workers the chance to dump recover states (they listen for SIGINT), then stop everything
        sched.stop_all(signal.SIGINT if recover_mode in ("auto", "save") else signal.SIGTERM)
        if isinstance(sched, sched_client.LocalSchedulerClient):
            names = ["master_worker/0", "model_worker/0"]
            failed = getattr(e, "worker_type", None)  # the worker that failed first tells the story: print it first
            if failed:
                names = [failed] + [n for n in names if n != failed]
            for info_name in names:
                tail = sched.log_tail(info_name)
                if tail:
                    logger.error(f"---- tail of {info_name} ----\n{tail}")
        if recover_mode == "auto" and recover_count < getattr(exp_cfg, "recover_retries", 1) and not isinstance(e, KeyboardInterrupt):
            logger.warning(f"run failed ({e}); recovering (attempt {recover_count + 1})")
            return main_start(exp_cfg, recover_count + 1, timeout, env_vars)
        raise
    finally:
        sched.stop_all()
    return sys_cfg


# ------------------------------------------------------------------------------------------- command line
def start_registered(args):
    """`start`: build the experiment registered under `--experiment_name` and launch it (parity: apps/main.py:76-231, where the
    name selects `config_package.make_experiment(name)` and the flags configure the launcher, not the experiment)."""
    import realhf_b200.experiments.algos  # noqa: F401  (registers the quickstart experiments)
    from realhf_b200.api.quickstart import QUICKSTART_EXPERIMENTS
    from realhf_b200.api.system import ALL_EXPERIMENT_CLASSES
    env_vars = {}
    if args.user_code:
        from realhf_b200.base.importing import import_usercode
        import_usercode(args.user_code, "real_user_code")
        env_vars["REAL_USER_CODE"] = os.path.abspath(args.user_code)
        os.environ["REAL_USER_CODE"] = env_vars["REAL_USER_CODE"]
    name = args.experiment_name
    if name in ALL_EXPERIMENT_CLASSES:
        exp = ALL_EXPERIMENT_CLASSES[name]()
    elif name in QUICKSTART_EXPERIMENTS:
        exp = QUICKSTART_EXPERIMENTS[name]()
    else:
        raise SystemExit(f"no experiment registered as `{name}` (registered: "
                         f"{sorted(set(ALL_EXPERIMENT_CLASSES) | set(QUICKSTART_EXPERIMENTS))}); --user_code FILE imports the module that registers it")
    for n in (name, args.trial_name):
        if "_" in n:
            raise SystemExit(f"experiment_name / trial_name must not contain `_` (got `{n}`)")
    exp.experiment_name, exp.trial_name = name, args.trial_name
    for k in ("mode", "partition", "wandb_mode", "recover_mode", "recover_retries", "debug", "ignore_worker_error"):
        setattr(exp, k, getattr(args, k))
    if args.image_name is not None:
        exp.image_name = args.image_name
    if args.allocation_mode is not None:
        exp.allocation_mode = args.allocation_mode
    return main_start(exp, timeout=args.timeout, env_vars=env_vars)


def main(argv=None):
    """`python -m realhf_b200.apps.main <cmd>`: operate on a trial from another shell (parity: apps/main.py:233-257,330-450;
    experiments are STARTED with `python -m realhf_b200.apps.quickstart <experiment> key=value ...`).

      start -e NAME -f TRIAL [...]      launch a REGISTERED experiment (`api.system.register_experiment`, or a quickstart experiment with
                                        its default options; `--user_code FILE` imports the module that registers it, here and in
                                        every worker) with the reference's launcher flags: --mode, --partition, --wandb_mode,
                                        --image_name, --ignore_worker_error, --debug, --recover_mode, --recover_retries,
                                        --allocation_mode
      status -e EXP -f TRIAL            worker statuses (RUNNING / PAUSED / COMPLETED / ERROR / LOST)
      pause | resume -e EXP -f TRIAL    the master stops issuing model function calls after its current step / continues
      stop -e EXP -f TRIAL [--mode M]   graceful stop through the master; with --mode slurm also `scancel` the trial's jobs
      find_config -r REGEX              list the registered experiment names that match
      profile_layers ...                forwards to `realhf_b200.apps.profile_layers`
    """
    import argparse
    import re
    ap = argparse.ArgumentParser(prog="realhf_b200.apps.main")
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name in ("status", "pause", "resume", "stop"):
        sp = sub.add_parser(name)
        sp.add_argument("--experiment_name", "-e", required=True)
        sp.add_argument("--trial_name", "-f", required=True)
        if name == "stop":
            sp.add_argument("--mode", default="local", choices=["local", "slurm"])
        if name == "status":
            sp.add_argument("--n_model_workers", "-n", type=int, default=None)
            sp.add_argument("--rpc", action="store_true", help="also query every worker's control endpoint (live progress, memory)")
    sp = sub.add_parser("start")
    sp.add_argument("--experiment_name", "-e", required=True, help="name the experiment was registered under")
    sp.add_argument("--trial_name", "-f", required=True)
    sp.add_argument("--mode", default="local", choices=["local", "slurm", "ray"])
    sp.add_argument("--partition", default=None, help="slurm partition")
    sp.add_argument("--wandb_mode", default="disabled", choices=["online", "offline", "disabled"])
    sp.add_argument("--image_name", default=None, help="container image of the workers (slurm)")
    sp.add_argument("--ignore_worker_error", action="store_true", help="keep the other workers running when one fails (debugging)")
    sp.add_argument("--debug", action="store_true", help="run workers without -O (assertions on)")
    sp.add_argument("--recover_mode", default="disabled", choices=["disabled", "auto", "save", "resume"])
    sp.add_argument("--recover_retries", type=int, default=1)
    sp.add_argument("--allocation_mode", default=None, help="override the experiment's allocation mode (manual / heuristic / search / d2m2p1 ...)")
    sp.add_argument("--user_code", default=None, help="python file that registers the experiment (imported here and by every worker)")
    sp.add_argument("--timeout", type=float, default=None)
    sp = sub.add_parser("find_config")
    sp.add_argument("--regex", "-r", required=True)
    sub.add_parser("profile_layers", add_help=False)
    args, rest = ap.parse_known_args(argv)
    if args.cmd == "profile_layers":
        from realhf_b200.apps import profile_layers
        return profile_layers.main(rest)
    if rest:
        ap.error(f"unrecognized arguments: {rest}")
    if args.cmd == "start":
        return start_registered(args)
    if args.cmd == "find_config":
        import realhf_b200.experiments.algos  # noqa: F401
        import realhf_b200.experiments.profile  # noqa: F401
        from realhf_b200.api.quickstart import QUICKSTART_EXPERIMENTS
        from realhf_b200.api.system import ALL_EXPERIMENT_CLASSES
        names = sorted(n for n in set(QUICKSTART_EXPERIMENTS) | set(ALL_EXPERIMENT_CLASSES) if re.match(args.regex, n))
        print("\n".join(names) if names else "No matched experiment names.")
        return names
    exp, trial = args.experiment_name, args.trial_name
    if args.cmd == "status":
        keys = name_resolve.find_subtree(f"{exp}/{trial}/status")
        out = {}
        for k in keys:
            try:
                out[k.split("/status/", 1)[1]] = name_resolve.get(k)
            except name_resolve.NameEntryNotFoundError:
                out[k.split("/status/", 1)[1]] = "LOST"
        for k in sorted(out):
            print(f"{k}: {out[k]}")
        if getattr(args, "rpc", False):
            live = worker_panel_status(exp, trial)
            for k in sorted(live):
                print(f"{k} [rpc]: {live[k]}")
            out = dict(out, rpc=live)
        return out
    if args.cmd == "pause":
        return pause_experiment(exp, trial)
    if args.cmd == "resume":
        return resume_experiment(exp, trial)
    if args.cmd == "stop":
        stop_experiment(exp, trial)
        if args.mode == "slurm":
            sched_client.make("slurm", exp, trial).stop_all()
        return None


if __name__ == "__main__":
    main()
