"""Host helpers of the engines (mirror of the parts of reference
umbrella/speculation/speculation_utils.py the hot path and the front-ends import)."""
from __future__ import annotations

import re

import torch


def make_causal_mask(input_ids_shape, device):
    """speculation_utils.py:4-14: bool [T, T], mask[i, j] = (j <= i)."""
    _, tgt_len = input_ids_shape
    idx = torch.arange(tgt_len, device=device)
    return idx[None, :] < (idx + 1)[:, None]


def find_first_element_position(tensor, elements):
    """speculation_utils.py:316-338."""
    if not isinstance(tensor, torch.Tensor):
        raise ValueError("Input 'tensor' must be a PyTorch Tensor.")
    if not all(isinstance(e, int) for e in elements):
        raise ValueError("All elements in 'elements' must be integers.")
    for idx, value in enumerate(tensor.flatten().tolist()):
        if value in elements:
            return idx
    return -1


def apply_repetition_penalty(input_ids: torch.LongTensor, logits: torch.FloatTensor, penalty: float):
    """speculation_utils.py:340-345."""
    logit = torch.gather(logits, 1, input_ids)
    logit = torch.where(logit < 0, logit * penalty, logit / penalty)
    return logits.scatter(1, input_ids, logit)


def apply_topk(logits: torch.FloatTensor, topk: int):
    """speculation_utils.py:347-352."""
    top_k = min(topk, logits.size(-1))
    remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
    return logits.masked_fill(remove, torch.finfo(logits.dtype).min)


def is_sentence_complete_regex(text):
    """speculation_utils.py:356-358."""
    return bool(re.search(r'[.?!。？！]\s*$', text))
