"""Generate tests/golden/m2m100_decoder_small.pt: HuggingFace ``M2M100Decoder`` (same fairseq lineage as the
SONAR text decoder: pre-LN layers with self-attention, encoder attention, ReLU FFN, final LayerNorm, sinusoidal
positions with offset 2, sqrt(d) embedding scale) on shared random weights, conditioned on a SINGLE encoder
position per sentence exactly as ``EmbeddingToTextModelPipeline`` does (``sonar/models/sonar_translation/model.py:48-53``).

    python tests/golden/make_m2m100_decoder_golden.py
"""

import os

import torch
from transformers import M2M100Config
from transformers.models.m2m_100.modeling_m2m_100 import M2M100Decoder

D, L, H, F_, V, S, B = 64, 2, 4, 128, 120, 9, 3


def main() -> None:
    torch.manual_seed(4321)
    cfg = M2M100Config(vocab_size=V, d_model=D, decoder_layers=L, decoder_attention_heads=H, decoder_ffn_dim=F_,
                       max_position_embeddings=32, pad_token_id=1, dropout=0.0, attention_dropout=0.0,
                       activation_dropout=0.0, activation_function="relu", scale_embedding=True, decoder_layerdrop=0.0)
    cfg._attn_implementation = "eager"
    dec = M2M100Decoder(cfg).eval().float()
    with torch.no_grad():
        for name, p in dec.named_parameters():
            p.copy_(torch.randn_like(p) * 0.1 + (1.0 if name.endswith("layer_norm.weight") else 0.0))
    tokens = torch.randint(4, V, (B, S))
    enc = torch.randn(B, 1, D) * 0.25
    with torch.no_grad():
        hid = dec(input_ids=tokens, encoder_hidden_states=enc).last_hidden_state  # [B,S,D] after layer_norm
        logits = hid @ dec.embed_tokens.weight.T

    hf = dec.state_dict()
    sd = {"decoder_frontend.embed.weight": hf["embed_tokens.weight"].clone()}
    m = {"self_attn.q_proj": "self_attn.q_proj", "self_attn.k_proj": "self_attn.k_proj",
         "self_attn.v_proj": "self_attn.v_proj", "self_attn.out_proj": "self_attn.output_proj",
         "self_attn_layer_norm": "self_attn_layer_norm",
         "encoder_attn.q_proj": "encoder_decoder_attn.q_proj", "encoder_attn.k_proj": "encoder_decoder_attn.k_proj",
         "encoder_attn.v_proj": "encoder_decoder_attn.v_proj", "encoder_attn.out_proj": "encoder_decoder_attn.output_proj",
         "encoder_attn_layer_norm": "encoder_decoder_attn_layer_norm",
         "fc1": "ffn.inner_proj", "fc2": "ffn.output_proj", "final_layer_norm": "ffn_layer_norm"}
    for i in range(L):
        for a, b in m.items():
            for wb in ("weight", "bias"):
                sd[f"decoder.layers.{i}.{b}.{wb}"] = hf[f"layers.{i}.{a}.{wb}"].clone()
    sd["decoder.layer_norm.weight"] = hf["layer_norm.weight"].clone()
    sd["decoder.layer_norm.bias"] = hf["layer_norm.bias"].clone()
    sd["final_proj.weight"] = sd["decoder_frontend.embed.weight"]
    here = os.path.dirname(os.path.abspath(__file__))
    torch.save({"config": dict(model_dim=D, vocab_size=V, max_seq_len=29, pad_idx=1, num_layers=L, num_heads=H,
                               ffn_inner_dim=F_),
                "state_dict": sd, "tokens": tokens, "encoder_output": enc, "hidden": hid, "logits": logits},
               os.path.join(here, "m2m100_decoder_small.pt"))
    print("wrote m2m100_decoder_small.pt", hid.shape)


if __name__ == "__main__":
    main()
