"""TEST INFRASTRUCTURE.  Front-end fixtures made by the reference's own code, so the product can be pinned byte for byte on
a box without /root/reference:
  tests/golden/templates.json   the chat-template tables (umbrella/templates.py:1-47)
  tests/golden/wire.bin         what the reference's send_data (umbrella/api/api_utils.py:3-7) puts on a socket for the
                                messages in tests/golden/wire.json (a request, a response, the greeting, terminate)
      python oracle/make_golden_frontend.py
"""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_templates", "/root/reference/umbrella/templates.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
out = {"Prompts": ref.Prompts, "SysPrompts": ref.SysPrompts, "ExtraPrompts": ref.ExtraPrompts}
with open(os.path.join(ROOT, "tests", "golden", "templates.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print({k: sorted(v) for k, v in out.items()})

# ---- wire frames written by the reference's send_data
import socket

spec = importlib.util.spec_from_file_location("ref_api_utils", "/root/reference/umbrella/api/api_utils.py")
ref_wire = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_wire)
messages = [
    {"context": "Tell me what you know about LSH in 100 words.", "max_new_tokens": 512, "temperature": 0.0},
    {"input_ids": list(range(3, 300)), "max_new_tokens": 64, "temperature": 0.6, "topp": 0.9, "topk": 32,
     "repetition_penalty": 1.05},
    {"generated_text": "ok", "generated_tokens": [5, 6, 7], "avg_accept_tokens": 3.25, "time_per_output_token": 4.5,
     "processed": True, "response": "Processed successfully"},
    {"status": "connected", "message": "Welcome to the server!"},
    {"terminate": True},
]
a, b = socket.socketpair()
for m in messages:
    ref_wire.send_data(a, m)
a.close()
raw = b""
while True:
    chunk = b.recv(1 << 16)
    if not chunk:
        break
    raw += chunk
b.close()
with open(os.path.join(ROOT, "tests", "golden", "wire.bin"), "wb") as f:
    f.write(raw)
with open(os.path.join(ROOT, "tests", "golden", "wire.json"), "w") as f:
    json.dump(messages, f, indent=1)
print("wire.bin", len(raw), "bytes,", len(messages), "messages")
