"""The reference's network description (src/config.rs:5-9, config/network.json) and how it maps onto
one multi-GPU box.

`NetworkConfig` is the same JSON the reference deserialises with serde: `slaves[i]` is the address
worker i serves the dispatcher on (PlonkSlave), `peers[i]` the address it serves the other workers on
(PlonkPeer, the fftExchange RPC).  With GPU workers on one box the PlonkPeer traffic is replaced by
NVLink (peer-memory stores or one NCCL all-to-all), so the `peers` entries only bootstrap the process
group: rank i = worker i = GPU i, rendezvous at peers[0].  `config/network.local8.json` is the
eight-workers-on-localhost instance."""
from __future__ import annotations

import ipaddress
import json
from dataclasses import dataclass
from typing import List, Tuple

Addr = Tuple[str, int]


def _socket_addr(s: str) -> Addr:
    """Rust `SocketAddr` text form: ip:port or [ipv6]:port"""
    host, sep, port = s.rpartition(":")
    if not sep or not port.isdigit() or not 0 < int(port) < 65536:
        raise ValueError(f"not a socket address: {s!r}")
    host = host[1:-1] if host.startswith("[") and host.endswith("]") else host
    ipaddress.ip_address(host)          # serde rejects host names too
    return host, int(port)


@dataclass
class NetworkConfig:
    slaves: List[Addr]
    peers: List[Addr]

    @classmethod
    def load(cls, path: str) -> "NetworkConfig":
        with open(path) as f:
            raw = json.load(f)
        cfg = cls([_socket_addr(s) for s in raw["slaves"]], [_socket_addr(s) for s in raw["peers"]])
        if len(cfg.slaves) != len(cfg.peers) or not cfg.slaves:
            raise ValueError("network config: need one peer address per slave, and at least one worker")
        return cfg

    @property
    def n_workers(self) -> int:
        return len(self.slaves)

    def workers_on(self, host: str) -> List[int]:
        """indices of the workers this box runs (`worker <me>`, src/worker.rs:443-449)"""
        return [i for i, (h, _) in enumerate(self.slaves) if h == host]

    def gpu_plan(self, host: str, n_gpus: int):
        """worker index -> CUDA device for the workers of `host`, one GPU each"""
        mine = self.workers_on(host)
        if len(mine) > n_gpus:
            raise ValueError(f"{len(mine)} workers on {host} but {n_gpus} GPUs")
        return {w: d for d, w in enumerate(mine)}

    def rendezvous_env(self, me: int) -> dict:
        """torch.distributed environment of worker `me`: rank = worker index, rendezvous at peers[0]"""
        if not 0 <= me < self.n_workers:
            raise ValueError(f"worker {me} not in the config")
        host, port = self.peers[0]
        return {"RANK": str(me), "WORLD_SIZE": str(self.n_workers), "MASTER_ADDR": host, "MASTER_PORT": str(port)}
