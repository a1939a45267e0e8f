"""Addresses and ports.  Parity: `realhf/base/network.py` (find_free_port / gethostname / gethostip)."""

from __future__ import annotations

import os
import socket


def find_free_port(host: str = "127.0.0.1") -> int:
    """A TCP port that was free a moment ago (bind to port 0 and read the kernel's pick back)."""
    with socket.socket() as s:
        s.bind((host, 0))
        return s.getsockname()[1]


def gethostname() -> str:
    return socket.gethostname()


def gethostip() -> str:
    """Address other workers use to reach this process (ZMQ master endpoint, torch.distributed rendezvous, control panel).

    `REAL_HOST_IP` wins.  Local mode (every worker on this host) publishes loopback: container hostnames often do not
    resolve.  Any other mode (slurm, ...) publishes a routable address of this host: the hostname's address, or, when that is
    loopback / unresolvable, the source address of the default route."""
    ip = os.environ.get("REAL_HOST_IP")
    if ip:
        return ip
    if os.environ.get("REAL_MODE", "LOCAL").upper() == "LOCAL":
        return "127.0.0.1"
    try:
        ip = socket.gethostbyname(socket.gethostname())
        if not ip.startswith("127."):
            return ip
    except OSError:
        pass
    try:  # no packet is sent: connect() on a UDP socket only selects the outgoing interface
        with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s:
            s.connect(("10.255.255.255", 1))
            return s.getsockname()[0]
    except OSError:
        return "127.0.0.1"
