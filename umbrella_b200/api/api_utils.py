"""Wire format of the request API (reference: umbrella/api/api_utils.py:3-18).

One message = a 4-byte big-endian length followed by ``pickle.dumps(dict)``.  The format is the reference's, so its
clients talk to this server and vice versa; the reading side here differs in that it insists on *exact* counts (a TCP
``recv`` may return fewer bytes than asked for, including for the 4-byte header) and reports a closed peer as
``ConnectionError`` instead of unpickling an empty buffer.

The payload is a pickle because the reference's is: unpickling runs code chosen by the sender, so this protocol is for
trusted peers only -- servers and the replica front bind to 127.0.0.1 unless told otherwise, as the reference does.
"""
from __future__ import annotations

import pickle
import struct

_HEADER = struct.Struct(">I")          # 4 bytes, big endian: the reference's ``int.to_bytes(4, 'big')``
MAX_MESSAGE_BYTES = (1 << 32) - 1


def _read_exact(conn, count: int) -> bytes:
    buf = bytearray(count)
    view, got = memoryview(buf), 0
    while got < count:
        n = conn.recv_into(view[got:], count - got)
        if n == 0:
            raise ConnectionError(f"peer closed the connection after {got} of {count} bytes")
        got += n
    return bytes(buf)


def send_data(conn, data) -> None:
    """Serialise ``data`` and write header + payload with one ``sendall``."""
    payload = pickle.dumps(data)
    if len(payload) > MAX_MESSAGE_BYTES:
        raise ValueError(f"message of {len(payload)} bytes does not fit the 4-byte length header")
    conn.sendall(_HEADER.pack(len(payload)) + payload)


def receive_data(conn):
    """Read one message; raises ``ConnectionError`` when the peer has gone away."""
    (length,) = _HEADER.unpack(_read_exact(conn, _HEADER.size))
    if length == 0:
        raise ConnectionError("zero-length message (no pickle is empty): the stream is out of step")
    return pickle.loads(_read_exact(conn, length))


GREETING = {"status": "connected", "message": "Welcome to the server!"}     # server.py:24 (clients log it)
TERMINATE_KEY = "terminate"                                                  # client.py:35 / server.py:31
