"""GPU: the request API in front of a real engine (SURVEY.md 8b "API wire", 8f-4).  A request served over the socket
by ``APIServer`` -- engine built and driven on the server's worker thread, CUDA graph of the iteration captured there --
must return exactly the tokens of a direct ``engine.generate`` call, also through the replica front."""
import pytest
import torch

from umbrella_b200.api.client import APIClient
from umbrella_b200.api.front import ReplicaFront
from umbrella_b200.api.server import APIServer

from test_gpu_engine import DEV, TARGET, _engine

pytestmark = pytest.mark.gpu


def _factory(device, config):
    assert device == DEV
    return _engine(graph=config.get("graph", True))


def test_api_server_serves_the_engine_bit_exactly():
    g = torch.Generator().manual_seed(31)
    prompts = [torch.randint(3, TARGET.vocab_size, (n,), generator=g).tolist() for n in (40, 57)]
    direct = _engine(graph=True)
    want = [direct.generate(input_ids=p, max_new_tokens=32)["generated_tokens"] for p in prompts]
    del direct
    server = APIServer({"graph": True}, device=DEV, port=0, engine_factory=_factory).start()
    front = ReplicaFront([("127.0.0.1", server.port)], port=0).start()
    try:
        with APIClient(port=server.port, retry_seconds=0.1, max_retries=50) as c:
            for p, w in zip(prompts, want):
                out = c.get_output(input_ids=p, max_new_tokens=32, temperature=0.0)
                assert out["processed"] is True and out["generated_tokens"] == w
                assert out["large_model_steps"] >= 1 and out["avg_accept_tokens"] >= 1.0
            # an empty request is the engine's empty answer, not an error (static...:373-390)
            empty = c.get_output(input_ids=[], max_new_tokens=8)
            assert empty["processed"] is True and empty.get("generated_tokens", []) == []
        with APIClient(port=front.port, retry_seconds=0.1, max_retries=50) as c:
            out = c.get_output(input_ids=prompts[0], max_new_tokens=32, temperature=0.0)
            assert out["generated_tokens"] == want[0] and out["replica"] == 0
        assert server.requests_served == 4
    finally:
        front.shutdown()
        server.shutdown()


class _WordTokenizer:
    """One word per token id, so the streamed text maps back to ids."""

    def decode(self, ids, **kwargs):
        return "".join(f" w{int(i)}" for i in ids)


@pytest.mark.parametrize("kind", ["static", "dynamic"])
def test_generate_stream_releases_the_tokens_generate_commits(kind):
    if kind == "static":
        eng = _engine(graph=True)
    else:
        from umbrella_b200.models.llama import SyntheticLlama
        from umbrella_b200.speculation.dynamic_speculation_engine import DynamicSpeculationEngine
        from test_gpu_engine import DRAFT
        eng = DynamicSpeculationEngine("draft", "target", device=DEV, max_length=512, safe_buffer=48, width=4,
                                       num_beams=6, depth=3)
        eng.initialize(draft_model=SyntheticLlama(DRAFT, max_length=512, awq=False, seed=22, device=DEV),
                       target_model=SyntheticLlama(TARGET, max_length=512, awq=True, seed=21, device=DEV))
        eng._set_eos([])
    eng.stop_distance = 8
    g = torch.Generator().manual_seed(33)
    prompt = torch.randint(3, TARGET.vocab_size, (48,), generator=g).tolist()
    want = eng.generate(input_ids=prompt, max_new_tokens=40)["generated_tokens"]
    eng.tokenizer = _WordTokenizer()
    chunks = list(eng.generate_stream(input_ids=prompt, max_new_tokens=40, temperature=0.0))
    texts = [t for t, _ in chunks]
    assert len(texts) >= 2 and all(b.startswith(a) for a, b in zip(texts, texts[1:]))
    ids = [int(w[1:]) for w in texts[-1].split()]
    assert ids == want[:len(ids)] and len(ids) >= 40          # generate() also reports the bonus token of the last step
    assert chunks[-1][1].startswith("Output Tokens ")
    assert eng.num_nodes == 0                                  # reset at the end, like generate()
    assert list(eng.generate_stream(input_ids=[], max_new_tokens=4)) == [("", "")]
