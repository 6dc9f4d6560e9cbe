"""Interoperation with the reference's own API code (SURVEY.md 8b "API wire"): its client against this server, this
client and the replica front against its server.  The reference runs in subprocesses from /root/reference (read-only,
present in the build container only -- skipped elsewhere); engines are stand-ins on both sides."""
import json
import os
import socket
import subprocess
import sys
import time

import pytest

from umbrella_b200.api.client import APIClient
from umbrella_b200.api.front import ReplicaFront
from umbrella_b200.api.server import APIServer

from test_api import make_echo_engine

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "umbrella", "api")),
                                reason="the reference checkout is not on this machine")

REF_CLIENT = r"""
import json, sys
sys.path.insert(0, "/root/reference")
from umbrella.api.client import APIClient          # the reference's client, unmodified
c = APIClient(port=int(sys.argv[1]))
c.run()
outs = [c.get_output(context="abc", max_new_tokens=3, temperature=0.0),
        c.get_output(input_ids=[5, 6, 7], max_new_tokens=2)]
c.close()
print("RESULT " + json.dumps(outs))
"""

REF_SERVER = r"""
import sys
sys.path.insert(0, sys.argv[2])
from oracle.make_golden import import_reference    # stubs for the reference's absent GPU wheels
import_reference()
import umbrella.api.server as srv                  # the reference's server, unmodified
assert srv.__file__.startswith("/root/reference")

class Engine:
    def initialize(self):
        pass
    def generate(self, **api_args):
        ids = api_args.get("input_ids") or [ord(ch) for ch in api_args.get("context", "")]
        api_args["generated_tokens"] = [sum(ids) % 1000, len(ids)]
        api_args["generated_text"] = "from the reference server"
        return api_args

srv.AutoEngine.from_config = staticmethod(lambda device, **config: Engine())
srv.APIServer(config={}, device="cpu", port=int(sys.argv[1]), max_client=4).run()
"""


def test_reference_client_against_this_server():
    server = APIServer({}, device="cpu", port=0, engine_factory=make_echo_engine).start()
    try:
        r = subprocess.run([sys.executable, "-c", REF_CLIENT, str(server.port)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-2000:]
        outs = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
        assert [o["processed"] for o in outs] == [True, True]
        assert outs[0]["context"] == "abc" and len(outs[0]["generated_tokens"]) == 3
        assert outs[1]["generated_tokens"] == [18, 19]
        assert "Server confirmation" in r.stderr            # the reference client logged this server's greeting
        assert server.requests_served == 2
    finally:
        server.shutdown()


def test_this_client_and_front_against_the_reference_server():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    proc = subprocess.Popen([sys.executable, "-c", REF_SERVER, str(port), ROOT], stdout=subprocess.PIPE,
                            stderr=subprocess.STDOUT, text=True)
    front = None
    try:
        c = APIClient(port=port, retry_seconds=0.5, max_retries=240)       # the reference imports torch + flashinfer first
        c.run()
        assert c.greeting == {"status": "connected", "message": "Welcome to the server!"}
        out = c.get_output(input_ids=[1, 2, 3], max_new_tokens=4)
        assert out["processed"] is True and out["response"] == "Processed successfully"
        assert out["generated_tokens"] == [6, 3] and out["generated_text"] == "from the reference server"
        c.close()
        # the replica front treats the reference server as one of its replicas
        mine = APIServer({}, device="cpu:1", port=0, engine_factory=make_echo_engine).start()
        front = ReplicaFront([("127.0.0.1", port), ("127.0.0.1", mine.port)], port=0).start()
        with APIClient(port=front.port, retry_seconds=0.1, max_retries=50) as fc:
            outs = [fc.get_output(input_ids=[i, 1], max_new_tokens=2) for i in range(4)]
        assert [o["replica"] for o in outs] == [0, 1, 0, 1] and all(o["processed"] for o in outs)
        assert outs[0]["generated_text"] == "from the reference server" and outs[1]["device"] == "cpu:1"
        mine.shutdown()
    finally:
        if front is not None:
            front.shutdown()
        proc.kill()
        out, _ = proc.communicate(timeout=30)
    assert "Traceback" not in (out or ""), out[-2000:]
