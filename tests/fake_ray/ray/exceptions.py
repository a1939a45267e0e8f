"""Exception names of `ray.exceptions` used by the scheduler client."""


class RayError(Exception):
    pass


class GetTimeoutError(RayError, TimeoutError):
    pass


class RayTaskError(RayError):
    pass


class TaskCancelledError(RayError):
    pass


class WorkerCrashedError(RayError):
    pass
