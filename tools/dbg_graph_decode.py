import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bitdelta_b200 as bd
from bitdelta_b200 import demo_backend as db
from bitdelta_b200.decode import GraphedDecoder
from transformers import LlamaConfig, LlamaForCausalLM
DEV = torch.device("cuda:0")
g = np.load(os.path.join("tests", "golden", "tiny_llama.npz"))
def t_bf16(bits): return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)
cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=96, max_position_embeddings=64, tie_word_embeddings=False)
sd = {k[len("basesd::"):]: t_bf16(g[k]) for k in g.files if k.startswith("basesd::")}
path = os.path.join("tests", "golden", "tiny_llama_diff.pt")
for fuse in (False, True):
    model = LlamaForCausalLM(cfg).to(torch.bfloat16); model.load_state_dict(sd); model = model.to(DEV).eval()
    ckpts = []
    for i in range(3):
        d = torch.load(path, weights_only=False)
        ck = {k: (v.detach().to(DEV).to(torch.bfloat16) if v.is_floating_point() else v.to(DEV)) for k, v in d.items()}
        for k in ck:
            if k.endswith(".coeff"): ck[k] = ck[k] * (1.0 + 0.5 * i)
        ckpts.append(ck)
    ids = torch.from_numpy(g["ids"]).to(DEV)[:1, :8].repeat(3, 1)
    mask = torch.ones_like(ids); mask[1, :3] = 0; ids = ids * mask
    n = 10
    db.cached_modules.clear()
    db.register_diff_compress(model, ckpts)
    if fuse: db.fuse_sibling_projections(model)
    print("fuse", fuse)
    print("dynamic ", bd.greedy_decode(model, ids, mask, n).tolist())
    print("dynamic2", bd.greedy_decode(model, ids, mask, n).tolist())
    dec = GraphedDecoder(model, max_cache_len=32)
    print("static1 ", dec.decode(ids, mask, n, use_graph=False).tolist())
    print("static2 ", dec.decode(ids, mask, n, use_graph=False).tolist())
    print("graph1  ", dec.decode(ids, mask, n).tolist())
    print("graph2  ", dec.decode(ids, mask, n).tolist())
    print("static3 ", dec.decode(ids, mask, n, use_graph=False).tolist())
    db.unregister_diff_compress(model); db.cached_modules.clear()
