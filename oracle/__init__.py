"""CPU oracle for the speculative-decoding verify path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as the shipped path.  The product path
(``umbrella_b200``) raises if its CUDA library is missing; it never falls back
to anything in here.

What is restated (all citations relative to /root/reference, UMbreLLa @ ff1ac17):

* ``oracle.tree``       -- Sequoia growmaps, tree accept, KV gather, state advance
                           (umbrella/sequoia_utils.py:40-130,
                            umbrella/speculation/static_speculation_engine.py:282-351,
                            umbrella/speculation/dynamic_speculation_engine.py:215-327,
                            umbrella/attn/cache.py:41-49,136-144)
* ``oracle.attention``  -- tree-masked GQA attention (umbrella/attn/cache.py:53-87,169-192)
* ``oracle.awq``        -- AWQ INT4 GEMM format + arithmetic
                           (umbrella/quantization/awq_utils.py:63-86; the arithmetic lives in
                            the un-vendored dependency autoawq-kernels==0.0.8 / autoawq==0.2.7.post3)
* ``oracle.llama``      -- RMSNorm, RoPE, the Llama verify forward
                           (umbrella/models/llama.py:262-322, umbrella/models/model_utils.py:17-64)

Pinning status (see DESIGN.md "Oracle"):

* tree / accept / gather / mask / RoPE / attention: PINNED against outputs of the
  reference's own Python run on CPU in the build container
  (``oracle/make_golden.py`` -> ``tests/golden/*``) and against the six shipped
  growmap JSONs.
* AWQ: the *format* (nibble order, packing) is pinned against the in-image
  ``vllm`` re-implementation of the AutoAWQ packer; the GEMM *arithmetic* of
  ``awq_ext.gemm_forward_cuda`` has no golden vector anywhere in the reference
  -> "parity unpinned" for that one function (restated from its published
  algorithm; checked on the GPU box against vLLM's build of the same kernel).
"""
