"""Chat templates of the front-ends: ``Prompts[name].format(user_text)`` after ``SysPrompts[name]``
(reference: umbrella/templates.py:1-47; used by app/*.py, examples/*.py and ``configs/*.json``'s "template" key).

The tables are assembled from each model family's role markers instead of being written out, so the marker of a
family is stated once; tests/golden/templates.json (made from the reference by oracle/make_golden_frontend.py) pins
the resulting strings byte for byte.
"""
from __future__ import annotations


def _llama3_header(role: str) -> str:
    return f"<|start_header_id|>{role}<|end_header_id|>\n\n"


def _chatml(role: str, body: str) -> str:
    return f"<|im_start|>{role}\n{body}<|im_end|>\n"


_LLAMA3_EOT = "<|eot_id|>"
_LLAMA3_SYSTEM_TEXT = ("You are a knowledgeable, efficient, and direct AI assistant. Provide concise answers, focusing on "
                       "the key information needed. Offer suggestions tactfully when appropriate to improve outcomes. "
                       "Engage in productive collaboration with the user. ")

# one user turn, left open at the start of the assistant's answer; "{}" receives the user's text
Prompts = {
    "meta-llama3": "\n" + _llama3_header("user") + "{}" + _LLAMA3_EOT + "\n" + _llama3_header("assistant"),
    # code template: no newlines between the turns, and the answer starts inside the python tool tag
    "llama3-code": _llama3_header("user") + "{}" + _LLAMA3_EOT + _llama3_header("assistant") + "<|python_tag|>",
    "qwen": _chatml("user", "{}") + "<|im_start|>assistant\n",
    "gemma2-it": "<start_of_turn>user\n{}<end_of_turn>\n<start_of_turn>model\n",
    "gemma2": "{}",
    "mistral": "[INST] {} [/INST]",
}

# what precedes the first user turn (BOS marker + system turn where the family has one)
SysPrompts = {
    "meta-llama3": "<|begin_of_text|>" + _llama3_header("system") + _LLAMA3_SYSTEM_TEXT + _LLAMA3_EOT,
    "llama3-code": "<|begin_of_text|>" + _llama3_header("system") + "Environment: ipython" + _LLAMA3_EOT,
    "qwen": _chatml("system", "You are a helpful assistant."),
    "gemma2": "",
    "gemma2-it": "",
    "mistral": "",
}

# appended to the user's text by the code front-ends
ExtraPrompts = {
    "llama3-code": "\nAlways try to wrap what you write in a function.",
}
