"""Generate tests/golden/conformer_layer_small.pt: one HuggingFace ``SeamlessM4TConformerEncoderLayer`` (relative
position embeddings, BatchNorm conv module, swish) on random weights -- the independent implementation whose
parameter names equal the fairseq names mapped in ``sonar/models/sonar_speech/handler.py:66-85`` (SURVEY probe E4).

    python tests/golden/make_conformer_golden.py
"""

import os

import torch
from transformers import SeamlessM4TConfig
from transformers.models.seamless_m4t import modeling_seamless_m4t as m

D, H, F_, K, B, S = 64, 4, 128, 31, 3, 40


def main() -> None:
    torch.manual_seed(777)
    cfg = SeamlessM4TConfig(hidden_size=D, speech_encoder_attention_heads=H, speech_encoder_intermediate_size=F_,
                            conv_depthwise_kernel_size=K, position_embeddings_type="relative",
                            speech_encoder_hidden_act="swish", speech_encoder_dropout=0.0, max_source_positions=128)
    layer = m.SeamlessM4TConformerEncoderLayer(cfg).eval().float()
    with torch.no_grad():
        for name, p in layer.named_parameters():
            p.copy_(torch.randn_like(p) * 0.15 + (1.0 if name.endswith("layer_norm.weight") or name.endswith("batch_norm.weight") else 0.0))
        layer.conv_module.batch_norm.running_mean.copy_(torch.randn(D) * 0.1)
        layer.conv_module.batch_norm.running_var.copy_(1.0 + torch.rand(D))
    relpos = m.SeamlessM4TConformerRelPositionalEmbedding(cfg)
    lens = torch.tensor([40, 23, 7])
    x = torch.randn(B, S, D)
    ok = torch.arange(S)[None, :] < lens[:, None]
    x = x * ok[:, :, None]
    attn_mask = torch.zeros(B, 1, 1, S).masked_fill(~ok[:, None, None, :], float("-inf"))
    with torch.no_grad():
        out, _ = layer(x, attention_mask=attn_mask, relative_position_embeddings=relpos(x), conv_attention_mask=ok)
    hf = layer.state_dict()
    p = "encoder.layers.0."
    mp = {"ffn1_layer_norm": "ffn1_layer_norm", "ffn1.intermediate_dense": "ffn1.inner_proj", "ffn1.output_dense": "ffn1.output_proj",
          "ffn2_layer_norm": "ffn2_layer_norm", "ffn2.intermediate_dense": "ffn2.inner_proj", "ffn2.output_dense": "ffn2.output_proj",
          "self_attn_layer_norm": "self_attn_layer_norm", "self_attn.linear_q": "self_attn.q_proj",
          "self_attn.linear_k": "self_attn.k_proj", "self_attn.linear_v": "self_attn.v_proj",
          "self_attn.linear_out": "self_attn.output_proj", "self_attn.linear_pos": "self_attn.sdpa.r_proj",
          "conv_module.layer_norm": "conv_layer_norm", "conv_module.pointwise_conv1": "conv.pointwise_conv1",
          "conv_module.depthwise_conv": "conv.depthwise_conv", "conv_module.batch_norm": "conv.batch_norm",
          "conv_module.pointwise_conv2": "conv.pointwise_conv2", "final_layer_norm": "layer_norm"}
    sd = {}
    for k, v in hf.items():
        if k == "self_attn.pos_bias_u":
            sd[p + "self_attn.sdpa.u_bias"] = v.clone()
        elif k == "self_attn.pos_bias_v":
            sd[p + "self_attn.sdpa.v_bias"] = v.clone()
        elif k.endswith("num_batches_tracked"):
            continue
        else:
            mod, leaf = k.rsplit(".", 1)
            sd[p + mp[mod] + "." + leaf] = v.clone()
    here = os.path.dirname(os.path.abspath(__file__))
    torch.save({"config": dict(model_dim=D, num_heads=H, ffn_inner_dim=F_, conv_kernel=K, num_layers=1),
                "state_dict": sd, "x": x, "lens": lens, "out": out,
                "generator": "transformers SeamlessM4TConformerEncoderLayer"},
               os.path.join(here, "conformer_layer_small.pt"))
    print("wrote conformer_layer_small.pt", out.shape)


if __name__ == "__main__":
    main()
