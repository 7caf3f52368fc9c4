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


class GraphedDecoder:
    """The decode step of ``greedy_steps`` as ONE CUDA graph over a static KV cache.

    The reference re-enters Python for every token of every layer (demo_backend.py:231-243: ``model(...)`` with a growing
    ``past_key_values`` and a re-concatenated attention mask).  Here the prefill runs eagerly into a ``StaticCache`` of
    ``max_cache_len`` positions and every following step -- embedding, all decoder layers (fused BinaryDiff launches,
    per-tenant norms), the per-tenant lm_heads, the greedy argmax and the bookkeeping for the next step -- is one graph
    replay on static buffers:

        tok [T,1]        the token fed to the step (the previous step's argmax)
        mask [T,L]       1 for every cache position a row may attend to (left-padded prompt, generated tokens)
        cache_pos [1]    the cache slot the step writes
        pos_ids [T,1]    the row's rotary position

    Batch row t is tenant t, as everywhere else.  Results are identical to the eager loop (``greedy_decode``): same
    modules, same kernels, only the launch mechanism differs.
    """

    def __init__(self, model, max_cache_len: int):
        from transformers import StaticCache

        self.model = model
        self.max_cache_len = int(max_cache_len)
        self.cache = StaticCache(config=model.config, max_cache_len=self.max_cache_len)
        self.graph = None
        self.tok = self.mask = self.cache_pos = self.pos_ids = self.next_tok = None

    def _forward_step(self):
        out = self.model(self.tok, attention_mask=self.mask, past_key_values=self.cache, cache_position=self.cache_pos,
                         position_ids=self.pos_ids, use_cache=True)
        nxt = torch.argmax(out.logits[:, -1, :], dim=-1)
        self.next_tok.copy_(nxt)
        # state of the NEXT step, advanced on the device so that a replay is a whole step
        self.tok.copy_(nxt[:, None])
        self.cache_pos.add_(1)
        self.pos_ids.add_(1)
        self.mask.index_fill_(1, self.cache_pos, 1)

    @torch.inference_mode()
    def prefill(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """Runs the prompt (left-padded like ``generate``) through the model into the static cache; returns the first
        generated token [T] and arms the step state."""
        T, P = input_ids.shape
        assert P < self.max_cache_len, "prompt does not fit the static cache"
        dev = input_ids.device
        self.cache.reset()
        pos = (attention_mask.long().cumsum(-1) - 1).clamp(min=0)
        out = self.model(input_ids, attention_mask=attention_mask, past_key_values=self.cache,
                         cache_position=torch.arange(P, device=dev), position_ids=pos, use_cache=True)
        first = torch.argmax(out.logits[:, -1, :], dim=-1)
        if self.tok is None:
            self.tok = torch.empty((T, 1), dtype=torch.long, device=dev)
            self.mask = torch.zeros((T, self.max_cache_len), dtype=attention_mask.dtype, device=dev)
            self.cache_pos = torch.empty((1,), dtype=torch.long, device=dev)
            self.pos_ids = torch.empty((T, 1), dtype=torch.long, device=dev)
            self.next_tok = torch.empty((T,), dtype=torch.long, device=dev)
        assert self.tok.shape[0] == T, "the graph is captured for a fixed number of tenants"
        self.tok.copy_(first[:, None])
        self.mask.zero_()
        self.mask[:, :P] = attention_mask
        self.mask[:, P] = 1
        self.cache_pos.fill_(P)
        self.pos_ids.copy_(pos[:, -1:] + 1)
        self._steps_left = self.max_cache_len - P - 1
        return first

    @torch.inference_mode()
    def capture(self, stream: Optional["torch.cuda.Stream"] = None):
        """Captures one decode step.  Call after a ``prefill`` (the capture itself runs two warm-up steps and one captured
        step on the armed state, then restores it).  Returns self."""
        assert self.tok is not None, "prefill first: the static buffers are sized by the prompt batch"
        layers = list(getattr(self.cache, "layers", []))
        assert not any(getattr(l, "is_sliding", False) for l in layers), \
            "sliding-window cache layers keep their write position on the host: not capturable (build the config with sliding_window=None)"
        # transformers' StaticLayer advances its own device-side write position (`cumulative_length`) with every update:
        # the warm-up steps below move it, so it is part of the state to put back
        state = [self.tok, self.mask, self.cache_pos, self.pos_ids]
        state += [l.cumulative_length for l in layers if isinstance(getattr(l, "cumulative_length", None), torch.Tensor)]
        saved = [t.clone() for t in state]
        stream = stream or torch.cuda.Stream(device=self.tok.device)
        stream.wait_stream(torch.cuda.current_stream(self.tok.device))
        with torch.cuda.stream(stream):
            for _ in range(2):  # warm-up on the capture stream: workspaces, lazily created buffers, cuBLAS handles
                self._forward_step()
            stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream):
                self._forward_step()
        torch.cuda.current_stream(self.tok.device).wait_stream(stream)
        # the warm-up / capture steps wrote cache slots past the prompt; they are overwritten before they are ever attended to
        for t, s in zip(state, saved):
            t.copy_(s)
        return self

    @torch.inference_mode()
    def step(self) -> torch.Tensor:
        """One decode step for every tenant: returns the [T] tokens it produced (a view of a static buffer -- clone to keep)."""
        assert self._steps_left > 0, "static cache exhausted"
        self._steps_left -= 1
        if self.graph is not None:
            self.graph.replay()
        else:
            self._forward_step()
        return self.next_tok

    @torch.inference_mode()
    def decode(self, input_ids, attention_mask, max_new_tokens: int, use_graph: bool = True) -> torch.Tensor:
        """Greedy tokens [T, max_new_tokens] (prefill + graph replays), the graphed counterpart of ``greedy_decode``.
        ``use_graph=False`` runs the same fixed-shape steps eagerly (what a CPU model always does)."""
        cols = [self.prefill(input_ids, attention_mask).clone()]
        if not use_graph:
            self.graph = None
        elif self.graph is None and input_ids.is_cuda:
            self.capture()
        for _ in range(max_new_tokens - 1):
            cols.append(self.step().clone())
        return torch.stack(cols, dim=1)
