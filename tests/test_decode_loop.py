"""The multi-tenant greedy decode loop (reference demo/demo_backend.py:190-258) on a plain tiny HF Llama on the CPU: the
loop is host logic around ``model(...)`` calls, so it is checked here against HF's own greedy ``generate``."""
import json

import pytest
import torch

from bitdelta_b200.decode import greedy_decode, greedy_steps, streaming_generator


@pytest.fixture(scope="module")
def tiny():
    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=96, max_position_embeddings=128, tie_word_embeddings=False, pad_token_id=0)
    model = LlamaForCausalLM(cfg).eval()
    T, L = 3, 9
    ids = torch.randint(1, 96, (T, L))
    mask = torch.ones(T, L, dtype=torch.long)
    mask[0, :3] = 0  # left padding like the reference's prompt batching (:297-311)
    mask[2, :5] = 0
    ids = ids * mask
    return model, ids, mask


def test_greedy_loop_matches_hf_generate(tiny):
    model, ids, mask = tiny
    n = 12
    ours = greedy_decode(model, ids, mask, n)
    ref = model.generate(ids, attention_mask=mask, max_new_tokens=n, do_sample=False, pad_token_id=0, eos_token_id=None)
    assert ours.shape == (3, n)
    assert torch.equal(ours, ref[:, ids.shape[1]:])


def test_stop_bookkeeping_and_early_exit(tiny):
    model, ids, mask = tiny
    full = greedy_decode(model, ids, mask, 10)
    # tenant 0 stops on its 3rd token, tenant 1 never, tenant 2 on its 1st
    stops = [{int(full[0, 2])}, set(), {int(full[2, 0])}]
    steps = list(greedy_steps(model, ids, mask, 10, stops))
    assert len(steps) == 10  # tenant 1 keeps the batch alive
    first0 = [int(v) for v in full[0]].index(int(full[0, 2]))  # the stop token may already occur earlier
    for i, (tok, stopped) in enumerate(steps):
        assert torch.equal(tok, full[:, i])  # stopped rows keep decoding: one launch serves the whole batch
        assert int(stopped[0]) == (first0 if i > first0 else -1)
        assert int(stopped[2]) == (0 if i > 0 else -1)
        assert int(stopped[1]) == -1
    # everybody stops -> the loop ends right after the step in which the last tenant stopped
    stops_all = [{int(full[t, 1])} | {int(full[t, 0])} for t in range(3)]
    assert len(list(greedy_steps(model, ids, mask, 10, stops_all))) == 1


class _Tok:
    def __init__(self, tag):
        self.tag = tag

    def decode(self, toks, skip_special_tokens=False):
        assert skip_special_tokens is False
        s = " ".join(f"{self.tag}{t}" for t in toks)
        return s + ("<|end_of_turn|>" if self.tag == "b" and len(toks) >= 2 else "")


def test_streaming_generator_ndjson_contract(tiny):
    model, ids, mask = tiny
    full = greedy_decode(model, ids, mask, 4)
    stops = [{int(full[0, 1])}, set(), set()]
    first0 = [int(v) for v in full[0]].index(int(full[0, 1]))
    lines = list(streaming_generator(model, [_Tok("a"), _Tok("b"), _Tok("c")], ids, mask, 4, stops))
    assert len(lines) == 4 and all(l.endswith("\n\n") for l in lines)
    for i, line in enumerate(lines):
        resp = json.loads(line)["response"]
        assert len(resp) == 3
        # tenant 0: whole generated prefix until the step after its stop token, then ("", "stop")
        if i > first0:
            assert resp[0] == ["", "stop"]
        else:
            assert resp[0] == [" ".join(f"a{int(t)}" for t in full[0, : i + 1]), "continue"]
        # tenant 1: the reference's end-of-turn rewrite
        want_b = " ".join(f"b{int(t)}" for t in full[1, : i + 1])
        assert resp[1] == [want_b + ("</s>" if i >= 1 else ""), "continue"]
        assert resp[2][1] == "continue"


def test_static_cache_decoder_matches_hf_generate(tiny):
    """GraphedDecoder without a graph (CPU): prefill into a StaticCache, then fixed-shape steps whose bookkeeping (token,
    cache slot, rotary position, attention mask) advances in place -- the same tokens as HF's greedy generate."""
    from bitdelta_b200.decode import GraphedDecoder

    model, ids, mask = tiny
    n = 12
    ref = model.generate(ids, attention_mask=mask, max_new_tokens=n, do_sample=False, pad_token_id=0, eos_token_id=None)[:, ids.shape[1]:]
    dec = GraphedDecoder(model, max_cache_len=40)
    assert torch.equal(dec.decode(ids, mask, n), ref)
    assert torch.equal(dec.decode(ids, mask, n), ref)  # the cache and the step state are re-armed by every prefill
    with pytest.raises(AssertionError):
        GraphedDecoder(model, max_cache_len=ids.shape[1]).prefill(ids, mask)
