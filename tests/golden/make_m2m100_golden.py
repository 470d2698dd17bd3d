"""Generate tests/golden/m2m100_small.pt -- an INDEPENDENT-implementation golden.

The reference's own notebook uses HuggingFace ``M2M100Encoder`` as the SONAR
text encoder (``/root/reference/examples/finetune_sonar_as_toxicity_classifier.ipynb``
cells 0/50/53), so its forward on shared random weights pins the op sequence the
oracle restates (position offset 2, [sin|cos] layout, sqrt(d) scale, pre-LN +
model-level final LN, key-only masking).  Run from the repo root:

    python tests/golden/make_m2m100_golden.py

Inputs/outputs are tiny (d=64, 3 layers) so the fixture stays < 1 MB.  HF pads
with id 1 and masks via ``attention_mask``; fairseq2 pads with 0 -- padded
positions are masked everywhere so the pad value must not matter, which the
test checks by feeding the oracle 0-padded ids.
"""

import os

import torch
from transformers import M2M100Config
from transformers.models.m2m_100.modeling_m2m_100 import M2M100Encoder

D, L, H, F_, V, S = 64, 3, 4, 128, 100, 12


def main() -> None:
    torch.manual_seed(1234)
    cfg = M2M100Config(
        vocab_size=V, d_model=D, encoder_layers=L, encoder_attention_heads=H,
        encoder_ffn_dim=F_, max_position_embeddings=32, pad_token_id=1,
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        activation_function="relu", scale_embedding=True, encoder_layerdrop=0.0,
    )
    cfg._attn_implementation = "eager"
    enc = M2M100Encoder(cfg).eval().float()
    with torch.no_grad():
        for name, p in enc.named_parameters():  # give LN / biases non-trivial values
            is_ln_gain = name.endswith("layer_norm.weight")
            p.copy_(torch.randn_like(p) * 0.1 + (1.0 if is_ln_gain else 0.0))

    lens = torch.tensor([12, 9, 5, 1, 7], dtype=torch.int64)
    B = lens.numel()
    ids_hf = torch.full((B, S), 1, dtype=torch.int64)  # HF pad id
    ids_fs = torch.zeros((B, S), dtype=torch.int64)  # fairseq2 collater pad id (0)
    for i, n in enumerate(lens.tolist()):
        row = torch.randint(4, V, (n,))
        ids_hf[i, :n] = row
        ids_fs[i, :n] = row
    mask = (torch.arange(S)[None, :] < lens[:, None]).long()
    with torch.no_grad():
        out = enc(input_ids=ids_hf, attention_mask=mask).last_hidden_state  # [B,S,D]
    valid = mask.bool()[:, :, None]
    emb = torch.where(valid, out, torch.zeros_like(out)).sum(1) / lens[:, None].float()

    # HF names -> fairseq2 names
    hf = enc.state_dict()
    sd = {"encoder_frontend.embed.weight": hf["embed_tokens.weight"].clone()}
    m = {"self_attn.q_proj": "self_attn.q_proj", "self_attn.k_proj": "self_attn.k_proj",
         "self_attn.v_proj": "self_attn.v_proj", "self_attn.out_proj": "self_attn.output_proj",
         "self_attn_layer_norm": "self_attn_layer_norm", "fc1": "ffn.inner_proj",
         "fc2": "ffn.output_proj", "final_layer_norm": "ffn_layer_norm"}
    for i in range(L):
        for a, b in m.items():
            for wb in ("weight", "bias"):
                sd[f"encoder.layers.{i}.{b}.{wb}"] = hf[f"layers.{i}.{a}.{wb}"].clone()
    sd["layer_norm.weight"] = hf["layer_norm.weight"].clone()
    sd["layer_norm.bias"] = hf["layer_norm.bias"].clone()

    here = os.path.dirname(os.path.abspath(__file__))
    torch.save({"config": dict(model_dim=D, vocab_size=V, max_seq_len=29, pad_idx=1,
                               num_layers=L, num_heads=H, ffn_inner_dim=F_),
                "state_dict": sd, "ids": ids_fs, "seq_lens": lens,
                "encoded_seqs": out, "sentence_embeddings": emb,
                "generator": "transformers %s M2M100Encoder" % __import__("transformers").__version__},
               os.path.join(here, "m2m100_small.pt"))
    print("wrote m2m100_small.pt", out.shape, emb.shape)


if __name__ == "__main__":
    main()
