"""Generate tests/golden/pooler_layers_small.pt: a 2-layer POST-LN decoder stack (self-attention, cross-attention with a
key-padding mask, ReLU FFN; residual + LayerNorm after each) computed by an INDEPENDENT implementation, HuggingFace
`BartDecoderLayer`, to pin `oracle/speech_encoder.py::pooler_layers` (the layers of the reference's
`AttentionEncoderOutputPooler`, sonar/nn/encoder_pooler.py:47-89, built at sonar/models/sonar_speech/factory.py:110-137 with
`norm_order=POST`).

    python tests/golden/make_pooler_golden.py

Weights are exported under the fairseq2 names the SONAR speech checkpoint uses (sonar_speech/handler.py:87-100)."""

import os

import torch
from transformers import BartConfig
from transformers.models.bart.modeling_bart import BartDecoderLayer

NAME_MAP = {  # HF -> fairseq2 (per layer)
    "self_attn.q_proj": "self_attn.q_proj", "self_attn.k_proj": "self_attn.k_proj", "self_attn.v_proj": "self_attn.v_proj",
    "self_attn.out_proj": "self_attn.output_proj", "self_attn_layer_norm": "self_attn_layer_norm",
    "encoder_attn.q_proj": "encoder_decoder_attn.q_proj", "encoder_attn.k_proj": "encoder_decoder_attn.k_proj",
    "encoder_attn.v_proj": "encoder_decoder_attn.v_proj", "encoder_attn.out_proj": "encoder_decoder_attn.output_proj",
    "encoder_attn_layer_norm": "encoder_decoder_attn_layer_norm", "fc1": "ffn.inner_proj", "fc2": "ffn.output_proj",
    "final_layer_norm": "ffn_layer_norm",
}


def main() -> None:
    torch.manual_seed(11)
    d, heads, ffn, layers = 64, 4, 128, 2
    cfg = BartConfig(d_model=d, decoder_attention_heads=heads, decoder_ffn_dim=ffn, activation_function="relu", dropout=0.0,
                     attention_dropout=0.0, activation_dropout=0.0, decoder_layers=layers)
    cfg._attn_implementation = "eager"
    hf = [BartDecoderLayer(cfg, layer_idx=i).eval() for i in range(layers)]
    sd = {}
    for i, layer in enumerate(hf):
        for p in layer.parameters():  # default init is near-identity LayerNorms / small weights: make it a real test
            torch.nn.init.normal_(p, 0.0, 0.3) if p.dim() > 1 else torch.nn.init.normal_(p, 0.5, 0.3)
        for k, v in layer.state_dict().items():
            mod, leaf = k.rsplit(".", 1)
            sd[f"encoder_pooler.decoder.layers.{i}.{NAME_MAP[mod]}.{leaf}"] = v.detach().clone()
    b, s = 3, 9
    lens = torch.tensor([9, 4, 1])
    x0 = torch.randn(b, 1, d)          # the single query position per utterance
    enc = torch.randn(b, s, d)
    key_ok = torch.arange(s)[None, :] < lens[:, None]
    add_mask = torch.zeros(b, 1, 1, s).masked_fill(~key_ok[:, None, None, :], float("-inf"))
    x = x0
    with torch.no_grad():
        for layer in hf:
            x = layer(x, attention_mask=None, encoder_hidden_states=enc, encoder_attention_mask=add_mask, use_cache=False)
            x = x[0] if isinstance(x, tuple) else x
    out = {"config": {"model_dim": d, "pooler_heads": heads, "pooler_ffn_inner_dim": ffn, "pooler_layers": layers},
           "state_dict": sd, "x0": x0, "enc": enc, "lens": lens, "out": x, "generator": "transformers BartDecoderLayer (post-LN)"}
    here = os.path.dirname(os.path.abspath(__file__))
    torch.save(out, os.path.join(here, "pooler_layers_small.pt"))
    print("wrote pooler_layers_small.pt", tuple(x.shape))


if __name__ == "__main__":
    main()
