"""``setup_logger`` (reference: umbrella/logging_config.py:2-11): one console handler per logger name, added once."""
from __future__ import annotations

import logging

_FORMAT = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"


def setup_logger(name: str = "infini_ai_chatbot", level: int = logging.INFO) -> logging.Logger:
    log = logging.getLogger(name)
    log.setLevel(level)
    if not log.handlers:                       # repeated calls (every module asks for the logger) add nothing
        handler = logging.StreamHandler()
        handler.setLevel(level)
        handler.setFormatter(logging.Formatter(_FORMAT))
        log.addHandler(handler)
    return log
