"""Multi-replica request front: N independent engines, one per GPU, behind one port.

A request's speculative-decoding loop is a sequential chain over private state and the reference never shards inside a
request (umbrella/api/server.py:43-49 serialises requests on its one engine), so an 8-GPU box serves eight requests at a
time by running eight ``APIServer`` replicas -- each its own process, CUDA context, copy of the weights and KV caches --
and placing every incoming request on the replica with the least work in flight.  No tensor crosses replicas: the front
moves pickled request / response dicts only, NVLink stays idle, and the aggregate rate is the sum of the replicas'.

    replicas = launch_replicas(config, devices=["cuda:0", ..., "cuda:7"], base_port=65440)
    front = ReplicaFront([("127.0.0.1", p) for p in replicas.ports], port=65432).start()
    # clients: the unchanged umbrella.api.client.APIClient(port=65432)

The front speaks the reference's wire protocol on both sides, so a backend may just as well be a reference server.
"""
from __future__ import annotations

import multiprocessing
import socket
import threading
import time

from ..logging_config import setup_logger
from ..utils import TextColors
from .api_utils import GREETING, TERMINATE_KEY, receive_data, send_data

logger = setup_logger()


class _Backend:
    """One persistent connection to a replica.  A replica answers in order and one at a time, so a request/response pair
    holds ``lock``; requests waiting for the same replica queue on that lock."""

    def __init__(self, index: int, host: str, port: int):
        self.index, self.host, self.port = index, host, port
        self.lock = threading.Lock()
        self.in_flight = 0          # guarded by ReplicaFront._pick_lock
        self.served = 0
        self.sock: socket.socket | None = None
        self.alive = False
        self.dead_since = 0.0

    def connect(self, timeout_s: float) -> None:
        deadline = time.monotonic() + timeout_s
        while True:
            sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            try:
                sock.connect((self.host, self.port))
                break
            except ConnectionRefusedError:
                sock.close()
                if time.monotonic() > deadline:
                    raise
                time.sleep(0.2)         # replicas take a while to load 37.7 GB of weights; keep knocking
        try:
            sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            receive_data(sock)          # the replica's greeting
        except (ConnectionError, OSError, EOFError):
            sock.close()
            raise
        self.sock = sock
        self.alive = True

    def call(self, request: dict, reconnect_timeout_s: float) -> dict:
        with self.lock:
            if not self.alive:
                self.connect(reconnect_timeout_s)       # a replica that was restarted on its port comes back here
            try:
                send_data(self.sock, request)
                return receive_data(self.sock)
            except (ConnectionError, OSError, EOFError):
                self.drop()
                raise

    def drop(self) -> None:
        self.alive = False
        self.dead_since = time.monotonic()
        if self.sock is not None:
            try:
                self.sock.close()
            except OSError:
                pass
            self.sock = None

    def close(self) -> None:
        if self.sock is not None:
            try:
                send_data(self.sock, {TERMINATE_KEY: True})
            except OSError:
                pass
            self.sock.close()
            self.sock = None
        self.alive = False


def pick_replica(in_flight: list[int], cursor: int, eligible: list[bool] | None = None) -> int:
    """Least work in flight among the ``eligible`` replicas (default: all); among equals the first at or after
    ``cursor`` (round robin), so an idle front spreads consecutive requests over all replicas instead of reusing
    replica 0.  -1 when no replica is eligible."""
    n = len(in_flight)
    ok = [True] * n if eligible is None else eligible
    loads = [in_flight[i] for i in range(n) if ok[i]]
    if not loads:
        return -1
    low = min(loads)
    for step in range(n):
        i = (cursor + step) % n
        if ok[i] and in_flight[i] == low:
            return i
    raise AssertionError("unreachable")


class ReplicaFront:
    def __init__(self, backends: list[tuple[str, int]], port: int = 65432, host: str = "127.0.0.1", max_client: int = 64,
                 connect_timeout_s: float = 1800.0, retry_dead_s: float = 10.0, reconnect_timeout_s: float = 0.5):
        if not backends:
            raise ValueError("ReplicaFront needs at least one replica address")
        self.backends = [_Backend(i, h, p) for i, (h, p) in enumerate(backends)]
        self.port, self.host, self.max_client = port, host, max_client
        self.connect_timeout_s = connect_timeout_s
        self.retry_dead_s = retry_dead_s                # a replica that dropped its connection is tried again this often
        self.reconnect_timeout_s = reconnect_timeout_s
        self._pick_lock = threading.Lock()
        self._cursor = 0
        self._closing = threading.Event()
        self.ready = threading.Event()
        self.server_socket: socket.socket | None = None
        self._accept_thread: threading.Thread | None = None

    # ---- placement ----------------------------------------------------------------------------------------------
    def _acquire(self, exclude: set[int]) -> _Backend | None:
        now = time.monotonic()
        with self._pick_lock:
            eligible = [b.index not in exclude and (b.alive or now - b.dead_since >= self.retry_dead_s)
                        for b in self.backends]
            i = pick_replica([b.in_flight for b in self.backends], self._cursor, eligible)
            if i < 0:
                return None
            self._cursor = (i + 1) % len(self.backends)
            b = self.backends[i]
            b.in_flight += 1
            if not b.alive:
                b.dead_since = now          # this request probes it; the others wait another retry_dead_s
            return b

    def _release(self, b: _Backend, served: bool) -> None:
        with self._pick_lock:
            b.in_flight -= 1
            b.served += int(served)

    def dispatch(self, request: dict) -> dict:
        """Place one request, wait for its replica's answer, tag it with the replica index.  Generation is a pure
        function of the request (an engine is reset after every request), so a request whose replica drops the
        connection is placed again on another one; only when no replica is left does the client get the error."""
        failed: set[int] = set()
        errors = []
        while True:
            b = self._acquire(failed)
            if b is None:
                return {**request, "processed": False,
                        "response": "no replica reachable" + "".join(f"; replica {i} unreachable: {e}" for i, e in errors)}
            try:
                reply = b.call(request, self.reconnect_timeout_s)
            except (ConnectionError, OSError, EOFError) as e:
                self._release(b, served=False)
                failed.add(b.index)
                errors.append((b.index, e))
                logger.error(TextColors.colorize(f"front: replica {b.index} ({b.host}:{b.port}) dropped: {e}", "red"))
                continue
            self._release(b, served=True)
            if isinstance(reply, dict):
                reply.setdefault("replica", b.index)
                if failed:
                    reply.setdefault("replicas_tried_before", sorted(failed))
            return reply

    def stats(self) -> dict:
        with self._pick_lock:
            return {"served": [b.served for b in self.backends], "in_flight": [b.in_flight for b in self.backends],
                    "alive": [b.alive for b in self.backends]}

    # ---- client side: the reference protocol ---------------------------------------------------------------------
    def _handle_client(self, conn, addr) -> None:
        try:
            send_data(conn, dict(GREETING, replicas=len(self.backends)))
            while not self._closing.is_set():
                request = receive_data(conn)
                if not isinstance(request, dict):
                    send_data(conn, {"processed": False, "response": "a request is a dict of generate() arguments"})
                    continue
                if request.get(TERMINATE_KEY, False):
                    break
                send_data(conn, self.dispatch(request))
        except (ConnectionError, OSError, EOFError) as e:
            if not self._closing.is_set():
                logger.error(TextColors.colorize(f"front: connection {addr} dropped: {e}", "red"))
        finally:
            try:
                conn.close()
            except OSError:
                pass

    def _accept_loop(self) -> None:
        assert self.server_socket is not None
        while not self._closing.is_set():
            try:
                conn, addr = self.server_socket.accept()
            except OSError:
                return
            threading.Thread(target=self._handle_client, args=(conn, addr), daemon=True).start()

    def _open(self) -> None:
        for b in self.backends:
            b.connect(self.connect_timeout_s)
        sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        sock.bind((self.host, self.port))
        sock.listen(self.max_client)
        self.port = sock.getsockname()[1]
        self.server_socket = sock
        logger.info(TextColors.colorize(
            f"replica front on port {self.port} over {len(self.backends)} replica(s): "
            + ", ".join(f"{b.host}:{b.port}" for b in self.backends), "cyan"))
        self.ready.set()

    def run(self) -> None:
        self._open()
        self._accept_loop()

    def start(self) -> "ReplicaFront":
        self._open()
        self._accept_thread = threading.Thread(target=self._accept_loop, name="umbrella-front", daemon=True)
        self._accept_thread.start()
        return self

    def shutdown(self) -> None:
        self._closing.set()
        if self.server_socket is not None:
            for stop in (lambda: self.server_socket.shutdown(socket.SHUT_RDWR), self.server_socket.close):
                try:
                    stop()                                  # shutdown() first: close() alone does not wake accept()
                except OSError:
                    pass
        for b in self.backends:
            b.close()
        if self._accept_thread is not None:
            self._accept_thread.join(timeout=5)


# ---- one server process per GPU -------------------------------------------------------------------------------------
def _replica_main(config, device, port, host, max_client, engine_factory, port_pipe) -> None:
    from .server import APIServer
    server = APIServer(config, device=device, port=port, max_client=max_client, host=host, engine_factory=engine_factory)
    try:
        server._open()
    except BaseException as e:  # noqa: BLE001
        port_pipe.send(("error", f"{type(e).__name__}: {e}"))
        raise
    port_pipe.send(("ok", server.port))
    server._accept_loop()


class ReplicaSet:
    def __init__(self, procs, ports, devices):
        self.procs, self.ports, self.devices = procs, ports, devices

    def addresses(self, host: str = "127.0.0.1") -> list[tuple[str, int]]:
        return [(host, p) for p in self.ports]

    def stop(self) -> None:
        for p in self.procs:
            if p.is_alive():
                p.terminate()
        for p in self.procs:
            p.join(timeout=10)


def launch_replicas(config: dict, devices: list[str], base_port: int = 0, host: str = "127.0.0.1", max_client: int = 4,
                    engine_factory=None, startup_timeout_s: float = 1800.0) -> ReplicaSet:
    """Start one ``APIServer`` process per entry of ``devices`` (``spawn``: every replica gets a fresh CUDA context) on
    ports ``base_port + i`` (``base_port=0``: ephemeral) and return once every replica has built its engine and listens.
    Replicas load concurrently.  ``engine_factory(device, config)`` must be picklable (a module-level function)."""
    ctx = multiprocessing.get_context("spawn")
    procs, pipes = [], []
    for i, dev in enumerate(devices):
        parent, child = ctx.Pipe(duplex=False)
        p = ctx.Process(target=_replica_main, name=f"umbrella-replica-{i}", daemon=True,
                        args=(config, dev, (base_port + i) if base_port else 0, host, max_client, engine_factory, child))
        p.start()
        child.close()
        procs.append(p)
        pipes.append(parent)
    ports = []
    try:
        for i, pipe in enumerate(pipes):
            if not pipe.poll(startup_timeout_s):
                raise TimeoutError(f"replica {i} ({devices[i]}) did not come up within {startup_timeout_s:g} s")
            try:
                status, value = pipe.recv()
            except EOFError:
                raise RuntimeError(f"replica {i} ({devices[i]}) exited while starting") from None
            if status != "ok":
                raise RuntimeError(f"replica {i} ({devices[i]}) failed to start: {value}")
            ports.append(value)
    except BaseException:
        ReplicaSet(procs, ports, devices).stop()
        raise
    return ReplicaSet(procs, ports, list(devices))
