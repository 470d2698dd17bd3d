"""Generate tests/golden/conformer_encoder_small.pt: HuggingFace ``SeamlessM4TConformerFeatureProjection`` +
``SeamlessM4TConformerEncoder`` (the w2v-BERT speech encoder of SeamlessM4T: LayerNorm(160) -> Linear on the 2-frame
stacked fbank, N Conformer blocks with relative position embeddings, final LayerNorm) on random weights -- an independent
implementation of the WHOLE stack ``oracle/speech_encoder.py::OracleSpeechEncoder.forward`` restates before the pooler
(frontend stacking and projection, block composition, the re-homed final LayerNorm of
``sonar/models/sonar_speech/handler.py:102-108``).  The frame stacking itself is the reshape HuggingFace's
``SeamlessM4TFeatureExtractor`` applies with ``stride=2``.

    python tests/golden/make_conformer_encoder_golden.py
"""

import os

import torch
from transformers import SeamlessM4TConfig
from transformers.models.seamless_m4t import modeling_seamless_m4t as m

D, H, F_, K, L, B, T = 64, 4, 128, 31, 2, 3, 60  # T fbank frames -> 30 positions

LAYER_MAP = {"ffn1_layer_norm": "ffn1_layer_norm", "ffn1.intermediate_dense": "ffn1.inner_proj", "ffn1.output_dense": "ffn1.output_proj",
             "ffn2_layer_norm": "ffn2_layer_norm", "ffn2.intermediate_dense": "ffn2.inner_proj", "ffn2.output_dense": "ffn2.output_proj",
             "self_attn_layer_norm": "self_attn_layer_norm", "self_attn.linear_q": "self_attn.q_proj",
             "self_attn.linear_k": "self_attn.k_proj", "self_attn.linear_v": "self_attn.v_proj",
             "self_attn.linear_out": "self_attn.output_proj", "self_attn.linear_pos": "self_attn.sdpa.r_proj",
             "conv_module.layer_norm": "conv_layer_norm", "conv_module.pointwise_conv1": "conv.pointwise_conv1",
             "conv_module.depthwise_conv": "conv.depthwise_conv", "conv_module.batch_norm": "conv.batch_norm",
             "conv_module.pointwise_conv2": "conv.pointwise_conv2", "final_layer_norm": "layer_norm"}


def main() -> None:
    torch.manual_seed(4242)
    cfg = SeamlessM4TConfig(hidden_size=D, speech_encoder_attention_heads=H, speech_encoder_intermediate_size=F_,
                            speech_encoder_layers=L, conv_depthwise_kernel_size=K, position_embeddings_type="relative",
                            speech_encoder_hidden_act="swish", speech_encoder_dropout=0.0, speech_encoder_layerdrop=0.0,
                            feature_projection_input_dim=160, max_source_positions=128)
    proj = m.SeamlessM4TConformerFeatureProjection(cfg).eval().float()
    enc = m.SeamlessM4TConformerEncoder(cfg).eval().float()
    with torch.no_grad():
        for mod in (proj, enc):
            for name, p in mod.named_parameters():
                is_gain = name.endswith("layer_norm.weight") or name.endswith("batch_norm.weight")
                p.copy_(torch.randn_like(p) * 0.15 + (1.0 if is_gain else 0.0))
        for layer in enc.layers:
            layer.conv_module.batch_norm.running_mean.copy_(torch.randn(D) * 0.1)
            layer.conv_module.batch_norm.running_var.copy_(1.0 + torch.rand(D))
    frame_lens = [60, 34, 12]
    fbank = torch.randn(B, T, 80)
    for i, n in enumerate(frame_lens):
        fbank[i, n:] = 0
    stacked = fbank.reshape(B, T // 2, 160)  # SeamlessM4TFeatureExtractor(stride=2)
    mask = (torch.arange(T // 2)[None, :] < (torch.tensor(frame_lens) // 2)[:, None]).long()
    with torch.no_grad():
        out = enc(proj(stacked), attention_mask=mask).last_hidden_state
    sd = {"encoder_frontend.post_extract_layer_norm.weight": proj.layer_norm.weight.clone(),
          "encoder_frontend.post_extract_layer_norm.bias": proj.layer_norm.bias.clone(),
          "encoder_frontend.model_dim_proj.weight": proj.projection.weight.clone(),
          "encoder_frontend.model_dim_proj.bias": proj.projection.bias.clone(),
          "layer_norm.weight": enc.layer_norm.weight.clone(), "layer_norm.bias": enc.layer_norm.bias.clone()}
    for i, layer in enumerate(enc.layers):
        p = f"encoder.layers.{i}."
        for k, v in layer.state_dict().items():
            if k == "self_attn.pos_bias_u":
                sd[p + "self_attn.sdpa.u_bias"] = v.clone()
            elif k == "self_attn.pos_bias_v":
                sd[p + "self_attn.sdpa.v_bias"] = v.clone()
            elif k.endswith("num_batches_tracked"):
                continue
            else:
                mod, leaf = k.rsplit(".", 1)
                sd[p + LAYER_MAP[mod] + "." + leaf] = v.clone()
    here = os.path.dirname(os.path.abspath(__file__))
    torch.save({"config": dict(model_dim=D, num_heads=H, ffn_inner_dim=F_, conv_kernel=K, num_layers=L),
                "state_dict": sd, "fbank": fbank, "frame_lens": frame_lens, "out": out,
                "generator": "transformers SeamlessM4TConformerFeatureProjection + SeamlessM4TConformerEncoder"},
               os.path.join(here, "conformer_encoder_small.pt"))
    print("wrote conformer_encoder_small.pt", out.shape)


if __name__ == "__main__":
    main()
