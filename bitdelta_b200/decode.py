"""Multi-tenant greedy decode loop around the fused modules -- the reference's ``streaming_generator``
(demo/demo_backend.py:190-258) without the FastAPI process around it.

Convention kept from the reference (:101-102): batch row ``t`` is tenant ``t``; the model's leaves have been replaced by
``register_diff_compress`` (``DiffCompressModule`` for the projections, ``DataParallelModule`` for embed / norms / lm_head),
so one ``model(...)`` call advances every tenant by one token.  Logits of tenants with a smaller vocabulary are padded
with ``finfo.min`` (``DataParallelModule``), so the argmax never lands in the padding.

``greedy_steps`` is the loop itself (prefill, argmax, attention-mask growth, stop bookkeeping :231-243); ``streaming_generator``
wraps it into the NDJSON lines the reference's ``/generate`` endpoint streams (:209-227), decoding only the tokens a tenant
has produced so far.  The HTTP server, prompt templates and tokenizer loading stay out of scope.
"""
from __future__ import annotations

import json
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import torch


def greedy_steps(model, input_ids: torch.Tensor, attention_mask: torch.Tensor, max_new_tokens: int,
                 stop_token_ids: Optional[Sequence[Iterable[int]]] = None) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    """Yields ``(next_token [T], stopped_pos [T])`` once per generated position.

    ``stopped_pos[t]`` is -1 while tenant ``t`` is running, else the step index at which it produced one of its stop
    tokens (the reference keeps decoding stopped rows -- the batch is one kernel launch -- and just reports them as
    stopped; so do we).  The loop ends after ``max_new_tokens`` steps or when every tenant has stopped (:242-243).
    """
    T = input_ids.shape[0]
    stops = [frozenset(int(v) for v in s) for s in stop_token_ids] if stop_token_ids is not None else [frozenset()] * T
    assert len(stops) == T, "one stop-token set per tenant"
    device = input_ids.device
    with torch.inference_mode():
        stopped_pos = torch.full((T,), -1, dtype=torch.long, device=device)
        outputs = model(input_ids, attention_mask=attention_mask, use_cache=True)
        for i in range(max_new_tokens):
            next_token = torch.argmax(outputs.logits[:, -1, :], dim=-1)
            yield next_token, stopped_pos.clone()
            attention_mask = torch.cat([attention_mask, attention_mask.new_ones((T, 1))], dim=-1)
            tok = next_token.tolist()
            for t in range(T):
                if tok[t] in stops[t] and stopped_pos[t] == -1:
                    stopped_pos[t] = i
            if bool(torch.all(stopped_pos != -1)):
                break
            if i + 1 == max_new_tokens:
                break
            outputs = model(next_token[:, None], past_key_values=outputs.past_key_values, attention_mask=attention_mask,
                            use_cache=True)


def greedy_decode(model, input_ids, attention_mask, max_new_tokens, stop_token_ids=None) -> torch.Tensor:
    """All generated tokens ``[T, n_steps]`` (rows keep receiving tokens after their stop token, like the reference's batch)."""
    cols: List[torch.Tensor] = [tok for tok, _ in greedy_steps(model, input_ids, attention_mask, max_new_tokens, stop_token_ids)]
    if not cols:
        return input_ids.new_zeros((input_ids.shape[0], 0))
    return torch.stack(cols, dim=1)


def streaming_generator(model, tokenizers: Sequence, input_ids, attention_mask, max_new_tokens,
                        stop_token_ids: Optional[Sequence[Iterable[int]]] = None) -> Iterator[str]:
    """NDJSON lines ``{"response": [[text, "continue" | "stop"], ...]}\\n\\n`` exactly as the reference streams them
    (:209-227): the text is the tenant's whole generated prefix decoded with its own tokenizer
    (``skip_special_tokens=False``), ``("", "stop")`` once the tenant has stopped, and the reference's end-of-turn
    rewrite for ``<|end_of_turn|>`` / ``<|im_end|>`` (:218-220)."""
    T = input_ids.shape[0]
    assert len(tokenizers) == T, "one tokenizer per tenant"
    generated: List[List[int]] = [[] for _ in range(T)]
    for next_token, stopped_pos in greedy_steps(model, input_ids, attention_mask, max_new_tokens, stop_token_ids):
        tok = next_token.tolist()
        stopped = stopped_pos.tolist()
        response = []
        for t in range(T):
            generated[t].append(tok[t])
            if stopped[t] != -1:
                response.append(("", "stop"))
                continue
            resp = tokenizers[t].decode(generated[t], skip_special_tokens=False)
            if "<|end_of_turn|>" in resp or "<|im_end|>" in resp:
                resp = resp.split("<|")[0] + "</s>"
            response.append((resp, "continue"))
        yield json.dumps({"response": response}) + "\n\n"
