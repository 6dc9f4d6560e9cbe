"""``APIClient`` (reference: umbrella/api/client.py:8-35): connect (retrying while the server is still loading its
models), read the greeting, then ``get_output(**api_args)`` = one request / one response; ``close()`` says
``{"terminate": True}`` first.  The same client talks to one ``APIServer`` or to a ``ReplicaFront``."""
from __future__ import annotations

import socket
import time

from ..logging_config import setup_logger
from ..utils import TextColors
from .api_utils import TERMINATE_KEY, receive_data, send_data

logger = setup_logger()


class APIClient:
    def __init__(self, port: int, host: str = "127.0.0.1", retry_seconds: float = 5.0, max_retries: int | None = None):
        self.port = port
        self.host = host
        self.retry_seconds = retry_seconds      # the reference waits 5 s between attempts, forever
        self.max_retries = max_retries
        self.client_socket: socket.socket | None = None
        self.greeting = None

    def run(self) -> None:
        attempt = 0
        while True:
            sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            try:
                sock.connect((self.host, self.port))
                break
            except ConnectionRefusedError:
                sock.close()                    # a socket whose connect failed is not reusable on every platform
                attempt += 1
                if self.max_retries is not None and attempt > self.max_retries:
                    raise
                logger.info(TextColors.colorize(
                    f"Server is not available, retrying in {self.retry_seconds:g} seconds...", "red"))
                time.sleep(self.retry_seconds)
        self.client_socket = sock
        self.greeting = receive_data(sock)
        logger.info(TextColors.colorize(f"Server confirmation: {self.greeting}", "cyan"))

    def get_output(self, **api_args) -> dict:
        if self.client_socket is None:
            raise RuntimeError("APIClient.run() has not connected yet")
        send_data(self.client_socket, api_args)
        return receive_data(self.client_socket)

    def close(self) -> None:
        if self.client_socket is None:
            return
        try:
            send_data(self.client_socket, {TERMINATE_KEY: True})
        except OSError:
            pass
        self.client_socket.close()
        self.client_socket = None

    def __enter__(self) -> "APIClient":
        self.run()
        return self

    def __exit__(self, *exc) -> None:
        self.close()
