"""Request API (SURVEY.md 8b "API wire", 8f-4 multi-replica front): wire format against the reference's byte layout,
server / client behaviour, placement over replicas.  CPU only: the engines here are stand-ins with the engine's
``generate(**api_args) -> dict`` contract (umbrella/speculation/static_speculation_engine.py:429-434)."""
import json
import logging
import os
import pickle
import socket
import threading
import time

import pytest

from umbrella_b200.api.api_utils import GREETING, receive_data, send_data
from umbrella_b200.api.client import APIClient
from umbrella_b200.api.front import ReplicaFront, launch_replicas, pick_replica
from umbrella_b200.api.server import APIServer

HERE = os.path.dirname(os.path.abspath(__file__))


class EchoEngine:
    """generate() as the engines define it: the request dict comes back with the result keys added."""

    def __init__(self, device, config):
        self.device, self.delay = device, float(config.get("delay", 0.0))
        self.threads = set()
        self.calls = 0

    def generate(self, **api_args):
        self.threads.add(threading.get_ident())
        self.calls += 1
        if api_args.get("explode"):
            raise ValueError("bad request")
        time.sleep(self.delay)
        ids = api_args.get("input_ids") or [ord(c) for c in api_args.get("context", "")]
        n = int(api_args.get("max_new_tokens", 4))
        toks = [(sum(ids) + i) % 1000 for i in range(n)]
        return {**api_args, "generated_text": " ".join(map(str, toks)), "generated_tokens": toks,
                "avg_accept_tokens": 3.3, "time_per_output_token": self.delay / max(n, 1), "device": self.device,
                "pid": os.getpid()}


def make_echo_engine(device, config):       # module level: launch_replicas pickles it for the spawned replicas
    return EchoEngine(device, config)


# ---- wire format -----------------------------------------------------------------------------------------------------
def test_wire_is_the_reference_layout():
    a, b = socket.socketpair()
    try:
        msg = {"context": "hello", "max_new_tokens": 7, "temperature": 0.0, "input_ids": list(range(50))}
        send_data(a, msg)
        raw = b.recv(1 << 16)
        # umbrella/api/api_utils.py:3-7: 4-byte big-endian length, then the pickle
        assert int.from_bytes(raw[:4], "big") == len(raw) - 4
        assert pickle.loads(raw[4:]) == msg
        # and a message framed the reference's way is read back
        payload = pickle.dumps({"terminate": True})
        a.sendall(len(payload).to_bytes(4, "big"))
        a.sendall(payload)
        assert receive_data(b) == {"terminate": True}
    finally:
        a.close()
        b.close()


def test_wire_survives_short_reads_and_large_messages():
    a, b = socket.socketpair()
    try:
        big = {"input_ids": list(range(200_000)), "context": "x" * 300_000}
        payload = pickle.dumps(big)
        frame = len(payload).to_bytes(4, "big") + payload

        def dribble():      # header split over two segments, body in odd-sized pieces
            a.sendall(frame[:2])
            time.sleep(0.01)
            a.sendall(frame[2:5])
            for i in range(5, len(frame), 70_001):
                a.sendall(frame[i:i + 70_001])

        t = threading.Thread(target=dribble)
        t.start()
        assert receive_data(b) == big
        t.join()
    finally:
        a.close()
        b.close()


def test_wire_reports_a_closed_peer():
    a, b = socket.socketpair()
    a.close()
    with pytest.raises(ConnectionError):
        receive_data(b)
    b.close()
    a, b = socket.socketpair()
    a.sendall((100).to_bytes(4, "big") + b"only a part")
    a.close()
    with pytest.raises(ConnectionError):
        receive_data(b)
    b.close()


# ---- one server ------------------------------------------------------------------------------------------------------
def test_server_client_round_trip_and_error_isolation():
    server = APIServer({"delay": 0.0}, device="cpu", port=0, max_client=2, engine_factory=make_echo_engine).start()
    try:
        assert server.ready.is_set() and server.port > 0
        c = APIClient(port=server.port, retry_seconds=0.05, max_retries=20)
        c.run()
        assert c.greeting == GREETING                                      # server.py:24
        out = c.get_output(context="abc", max_new_tokens=3, temperature=0.0)
        assert out["processed"] is True and out["response"] == "Processed successfully"     # server.py:52
        assert out["context"] == "abc" and len(out["generated_tokens"]) == 3
        assert out["generated_text"] == " ".join(map(str, out["generated_tokens"]))
        # a request the engine rejects is answered, and the replica keeps serving
        bad = c.get_output(explode=True)
        assert bad["processed"] is False and "ValueError" in bad["response"]
        again = c.get_output(input_ids=[1, 2, 3], max_new_tokens=2)
        assert again["processed"] is True and again["generated_tokens"] == [6, 7]
        # a second client while the first stays connected; every generate() ran on the one engine thread
        with APIClient(port=server.port, retry_seconds=0.05, max_retries=20) as c2:
            assert c2.get_output(context="z")["processed"] is True
        c.close()
        assert len(server.engine.threads) == 1 and threading.get_ident() not in server.engine.threads
        assert server.requests_served == 4
    finally:
        server.shutdown()


def test_server_start_reports_engine_failure():
    def broken(device, config):
        raise RuntimeError("no such checkpoint")
    with pytest.raises(RuntimeError, match="no such checkpoint"):
        APIServer({}, device="cpu", port=0, engine_factory=broken).start()


def test_client_gives_up_after_max_retries():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()                                   # nobody listens on this port now
    with pytest.raises(ConnectionRefusedError):
        APIClient(port=port, retry_seconds=0.01, max_retries=2).run()


# ---- replicas --------------------------------------------------------------------------------------------------------
def test_pick_replica_least_loaded_then_round_robin():
    assert pick_replica([0, 0, 0, 0], 0) == 0
    assert pick_replica([0, 0, 0, 0], 2) == 2
    assert pick_replica([1, 0, 1, 0], 2) == 3
    assert pick_replica([1, 0, 1, 0], 0) == 1
    assert pick_replica([2, 2, 1], 0) == 2
    assert pick_replica([3], 0) == 0
    assert pick_replica([0, 5, 1], 0, eligible=[False, True, True]) == 2
    assert pick_replica([0, 0], 0, eligible=[False, False]) == -1


def test_front_spreads_requests_over_replicas():
    delay, n_rep, n_clients, per_client = 0.08, 2, 4, 4
    servers = [APIServer({"delay": delay}, device=f"cpu:{i}", port=0, engine_factory=make_echo_engine).start()
               for i in range(n_rep)]
    front = ReplicaFront([("127.0.0.1", s.port) for s in servers], port=0).start()
    try:
        results = [[] for _ in range(n_clients)]

        def client(k):
            with APIClient(port=front.port, retry_seconds=0.05, max_retries=20) as c:
                assert c.greeting["status"] == "connected" and c.greeting["replicas"] == n_rep
                for j in range(per_client):
                    results[k].append(c.get_output(input_ids=[k, j], max_new_tokens=2))

        t0 = time.perf_counter()
        threads = [threading.Thread(target=client, args=(k,)) for k in range(n_clients)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        wall = time.perf_counter() - t0
        flat = [r for rs in results for r in rs]
        assert len(flat) == n_clients * per_client and all(r["processed"] for r in flat)
        # each answer is the right request's (placement does not mix up connections)
        for k in range(n_clients):
            for j, r in enumerate(results[k]):
                assert r["input_ids"] == [k, j] and r["generated_tokens"] == [k + j, k + j + 1]
                assert r["device"] == f"cpu:{r['replica']}"
        served = front.stats()["served"]
        assert sum(served) == 16 and min(served) >= 6, served          # least-in-flight keeps both replicas busy
        assert front.stats()["in_flight"] == [0, 0]
        serial = n_clients * per_client * delay                        # one engine: 1.28 s
        assert wall < 0.85 * serial, (wall, serial)                    # two replicas: ~0.64 s + overheads
        assert all(len(s.engine.threads) == 1 for s in servers)
    finally:
        front.shutdown()
        [s.shutdown() for s in servers]


def test_front_survives_a_lost_replica_and_takes_it_back():
    servers = [APIServer({}, device=f"cpu:{i}", port=0, engine_factory=make_echo_engine).start() for i in range(2)]
    front = ReplicaFront([("127.0.0.1", s.port) for s in servers], port=0, retry_dead_s=0.2).start()
    revived = None
    try:
        lost_port = servers[1].port
        servers[1].shutdown()
        with APIClient(port=front.port, retry_seconds=0.05, max_retries=20) as c:
            outs = [c.get_output(input_ids=[i], max_new_tokens=1) for i in range(4)]
            # every request is answered: the one placed on the lost replica is placed again on the other
            assert all(o["processed"] and o["replica"] == 0 for o in outs)
            assert any(o.get("replicas_tried_before") == [1] for o in outs)
            assert front.stats()["alive"] == [True, False] and front.stats()["in_flight"] == [0, 0]
            # the replica comes back on its port: after retry_dead_s the front reconnects and uses it again
            revived = APIServer({}, device="cpu:1", port=lost_port, engine_factory=make_echo_engine).start()
            time.sleep(0.3)
            outs = [c.get_output(input_ids=[i], max_new_tokens=1) for i in range(4)]
            assert all(o["processed"] for o in outs) and {o["replica"] for o in outs} == {0, 1}
            assert front.stats()["alive"] == [True, True]
            # nobody left: the client is told, per request
            servers[0].shutdown()
            revived.shutdown()
            gone = c.get_output(input_ids=[9], max_new_tokens=1)
            assert gone["processed"] is False and "no replica reachable" in gone["response"]
            assert gone["input_ids"] == [9]
    finally:
        front.shutdown()
        [s.shutdown() for s in servers]
        if revived is not None:
            revived.shutdown()


@pytest.mark.timeout(120)
def test_launch_replicas_one_process_per_device():
    reps = launch_replicas({"delay": 0.0}, devices=["cpu:0", "cpu:1"], engine_factory=make_echo_engine,
                           startup_timeout_s=90)
    front = None
    try:
        assert len(set(reps.ports)) == 2 and all(p.is_alive() for p in reps.procs)
        front = ReplicaFront(reps.addresses(), port=0).start()
        with APIClient(port=front.port, retry_seconds=0.05, max_retries=20) as c:
            outs = [c.get_output(context=f"request {i}", max_new_tokens=2) for i in range(4)]
        assert all(o["processed"] for o in outs)
        assert [o["replica"] for o in outs] == [0, 1, 0, 1]                # an idle front round-robins
        pids = {o["replica"]: o["pid"] for o in outs}
        assert pids[0] != pids[1] and os.getpid() not in pids.values()
        assert {o["device"] for o in outs} == {"cpu:0", "cpu:1"}
    finally:
        if front is not None:
            front.shutdown()
        reps.stop()
    assert not any(p.is_alive() for p in reps.procs)


# ---- the small front-end modules -------------------------------------------------------------------------------------
def test_templates_match_the_reference_tables():
    from umbrella_b200 import templates
    with open(os.path.join(HERE, "golden", "templates.json")) as f:
        gold = json.load(f)
    for table in ("Prompts", "SysPrompts", "ExtraPrompts"):
        assert getattr(templates, table) == gold[table], table
    text = templates.SysPrompts["meta-llama3"] + templates.Prompts["meta-llama3"].format("hi")
    assert text.startswith("<|begin_of_text|>") and text.endswith("<|start_header_id|>assistant<|end_header_id|>\n\n")


def test_text_colors_and_logger():
    from umbrella_b200.logging_config import setup_logger
    from umbrella_b200.utils import TextColors
    assert TextColors.colorize("x", "red") == "\033[31mx\033[0m"
    assert TextColors.colorize("x", "CYAN") == "\033[36mx\033[0m"
    assert TextColors.colorize("x", "no-such-colour") == "\033[0mx\033[0m"      # utils.py:24: falls back to reset
    a = setup_logger("umbrella_test_logger")
    b = setup_logger("umbrella_test_logger")
    assert a is b and len(a.handlers) == 1 and a.level == logging.INFO


def test_reference_import_paths_resolve():
    """every ``umbrella`` import statement of the reference's app/*.py and examples/*.py."""
    statements = [
        "from umbrella.api.client import APIClient",
        "from umbrella.api.server import APIServer",
        "from umbrella.logging_config import setup_logger",
        "from umbrella.models.auto_model import AutoModelLM",
        "from umbrella.sequoia_utils import measure_acceptance_rate, generate_sequoia_tree",
        "from umbrella.speculation.auto_engine import AutoEngine",
        "from umbrella.speculation.dynamic_speculation_engine import DynamicSpeculationEngine",
        "from umbrella.speculation.speculation_utils import make_causal_mask, is_sentence_complete_regex, "
        "find_first_element_position",
        "from umbrella.speculation.static_speculation_engine import StaticSpeculationEngine",
        "from umbrella.templates import Prompts, SysPrompts",
        "from umbrella.utils import TextColors",
    ]
    for st in statements:
        ns = {}
        exec(st, ns)
    from umbrella.api.client import APIClient as C
    from umbrella.api.server import APIServer as S
    assert C is APIClient and S is APIServer


def test_wire_golden_from_the_reference():
    """tests/golden/wire.bin holds frames written by the reference's own send_data (oracle/make_golden_frontend.py):
    receive_data reads them back, and send_data writes the same bytes."""
    with open(os.path.join(HERE, "golden", "wire.json")) as f:
        messages = json.load(f)
    with open(os.path.join(HERE, "golden", "wire.bin"), "rb") as f:
        raw = f.read()
    a, b = socket.socketpair()
    try:
        t = threading.Thread(target=lambda: (a.sendall(raw), a.shutdown(socket.SHUT_WR)))
        t.start()
        got = [receive_data(b) for _ in messages]
        t.join()
        assert got == messages
        with pytest.raises(ConnectionError):
            receive_data(b)                                  # nothing after the last frame
    finally:
        a.close()
        b.close()
    a, b = socket.socketpair()
    try:
        for m in messages:
            send_data(a, m)
        a.shutdown(socket.SHUT_WR)
        mine = b""
        while True:
            chunk = b.recv(1 << 16)
            if not chunk:
                break
            mine += chunk
        assert mine == raw
    finally:
        a.close()
        b.close()
