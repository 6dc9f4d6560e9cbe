"""``APIServer``: one engine behind a TCP port (reference: umbrella/api/server.py:10-75; started by app/api.py:19-21).

Protocol as in the reference: greet on connect (``GREETING``), then one request dict in, one response dict out, until
the client sends ``{"terminate": True}``.  A response is the engine's ``generate(**request)`` result plus
``processed`` / ``response``.

Differences in construction, not in protocol:
* the engine is built *and* driven by one worker thread, so every CUDA call of a replica (allocation, graph capture,
  replay, the per-thread current device) comes from the same thread; the reference builds it on the main thread and
  generates on another, serialising with a lock that its single worker never contends for;
* a request that makes the engine raise is answered with ``processed: False`` and the error text, and the worker goes on
  (in the reference the worker thread dies with the exception and every later client waits forever);
* ``port=0`` binds an ephemeral port (``self.port`` holds the real one once ``ready`` is set); ``start()`` /
  ``shutdown()`` run the server in the background for the replica front and the tests, ``run()`` blocks as before.
"""
from __future__ import annotations

import queue
import socket
import threading

from ..logging_config import setup_logger
from ..utils import TextColors
from .api_utils import GREETING, TERMINATE_KEY, receive_data, send_data

logger = setup_logger()
_STOP = object()


def _note(text: str, color: str = "cyan") -> None:
    logger.info(TextColors.colorize(text, color))


def _brief(request) -> str:
    """A request for the log: long token-id lists / contexts are cut (the reference prints them whole, which at a 30k-token
    prompt is a 200 KB log line per request)."""
    if not isinstance(request, dict):
        return repr(request)[:200]
    shown = {}
    for k, v in request.items():
        if isinstance(v, (list, tuple)) and len(v) > 16:
            shown[k] = f"<{len(v)} items: {list(v[:8])}...>"
        elif isinstance(v, str) and len(v) > 200:
            shown[k] = v[:200] + f"...<{len(v)} chars>"
        else:
            shown[k] = v
    return str(shown)


def _default_engine_factory(device: str, config: dict):
    from ..speculation.auto_engine import AutoEngine
    engine = AutoEngine.from_config(device, **config)
    engine.initialize()
    return engine


class APIServer:
    def __init__(self, config, device: str = "cuda:0", port: int = 65432, max_client: int = 4, host: str = "127.0.0.1",
                 engine_factory=None):
        self.config = config
        self.device = device
        self.port = port
        self.max_client = max_client
        self.host = host
        self.engine = None
        self.engine_factory = engine_factory or _default_engine_factory
        self.ready = threading.Event()          # set once the engine is built and the port is listening
        self.failed: BaseException | None = None
        self.requests_served = 0
        self._requests: queue.Queue = queue.Queue()
        self._closing = threading.Event()
        self._threads: list[threading.Thread] = []
        self._conns: set[socket.socket] = set()
        self._conns_lock = threading.Lock()
        self.server_socket: socket.socket | None = None

    # ---- the engine's thread ------------------------------------------------------------------------------------
    def _engine_loop(self) -> None:
        try:
            if self.device.startswith("cuda"):
                import torch
                torch.cuda.set_device(self.device)
            self.engine = self.engine_factory(self.device, self.config)
        except BaseException as e:  # noqa: BLE001 -- reported to run()/start(), which re-raise it
            self.failed = e
            self._engine_built.set()
            return
        self._engine_built.set()
        while True:
            item = self._requests.get()
            if item is _STOP:
                return
            addr, conn, request = item
            _note(f"Processing message from {addr}: {_brief(request)}")
            try:
                reply = {**self.engine.generate(**request), "processed": True, "response": "Processed successfully"}
            except Exception as e:  # noqa: BLE001 -- one bad request must not take the replica down
                logger.error(TextColors.colorize(f"Request from {addr} failed: {type(e).__name__}: {e}", "red"))
                reply = {**request, "processed": False, "response": f"{type(e).__name__}: {e}"}
            self.requests_served += 1
            try:
                send_data(conn, reply)
            except OSError as e:
                logger.error(TextColors.colorize(f"Error sending data to {addr}: {e}", "red"))

    # ---- one thread per client connection: reads requests, the engine thread writes the answers ------------------
    def handle_client(self, conn, addr) -> None:
        _note(f"Connection from {addr}")
        with self._conns_lock:
            self._conns.add(conn)
        try:
            send_data(conn, dict(GREETING))
            while not self._closing.is_set():
                request = receive_data(conn)
                _note(f"Received dictionary from {addr}: {_brief(request)}")
                if not isinstance(request, dict):
                    send_data(conn, {"processed": False, "response": "a request is a dict of generate() arguments"})
                    continue
                if request.get(TERMINATE_KEY, False):
                    break
                self._requests.put((addr, conn, request))
        except (ConnectionError, OSError, EOFError) as e:
            if not self._closing.is_set():
                logger.error(TextColors.colorize(f"Error handling data from {addr}: {e}", "red"))
        finally:
            with self._conns_lock:
                self._conns.discard(conn)
            try:
                conn.close()
            except OSError:
                pass
            _note(f"Connection with {addr} closed")

    def _accept_loop(self) -> None:
        assert self.server_socket is not None
        while not self._closing.is_set():
            try:
                conn, addr = self.server_socket.accept()
            except OSError:
                return                                      # the listening socket was closed by shutdown()
            t = threading.Thread(target=self.handle_client, args=(conn, addr), daemon=True)
            t.start()

    def _open(self) -> None:
        self._engine_built = threading.Event()
        worker = threading.Thread(target=self._engine_loop, name="umbrella-engine", daemon=True)
        worker.start()
        self._threads.append(worker)
        self._engine_built.wait()
        if self.failed is not None:
            raise self.failed
        sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        sock.bind((self.host, self.port))
        sock.listen(self.max_client)
        self.port = sock.getsockname()[1]
        self.server_socket = sock
        _note(f"Infini AI LLM server started successfully ({self.device}, port {self.port})")
        self.ready.set()

    def run(self) -> None:
        """Build the engine, listen, serve until the process ends (the reference's blocking entry point)."""
        self._open()
        self._accept_loop()

    def start(self) -> "APIServer":
        """``run()`` on a background thread; returns once the port is listening."""
        self._open()
        t = threading.Thread(target=self._accept_loop, name="umbrella-accept", daemon=True)
        t.start()
        self._threads.append(t)
        return self

    def shutdown(self) -> None:
        self._closing.set()
        if self.server_socket is not None:
            for stop in (lambda: self.server_socket.shutdown(socket.SHUT_RDWR), self.server_socket.close):
                try:
                    stop()                                  # shutdown() first: close() alone does not wake accept()
                except OSError:
                    pass
        with self._conns_lock:
            conns = list(self._conns)
        for c in conns:
            try:
                c.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass
        self._requests.put(_STOP)
        for t in self._threads:
            t.join(timeout=5)
