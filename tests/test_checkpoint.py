"""CPU tests of the legacy-fairseq checkpoint converters (sonar_b200/checkpoint.py) against the key maps of
sonar/models/sonar_text/handler.py:52-94,119-172 and sonar/models/sonar_speech/handler.py:47-110."""

import torch

from oracle.speech_encoder import OracleSpeechConfig, make_synthetic_speech_state_dict
from oracle.text_decoder import OracleDecoderConfig, make_synthetic_decoder_state_dict
from oracle.text_encoder import OracleEncoderConfig, make_synthetic_state_dict
from sonar_b200.checkpoint import (convert_sonar_speech_checkpoint, convert_sonar_text_decoder_checkpoint,
                                   convert_sonar_text_encoder_checkpoint)


def _unswap(e):  # inverse of (BOS,PAD,EOS,UNK) -> (PAD,UNK,BOS,EOS)
    f = e.clone()
    f[[1, 3, 0, 2]] = e[[0, 1, 2, 3]]
    return f


def test_text_encoder_converter_roundtrip():
    cfg = OracleEncoderConfig(model_dim=64, vocab_size=40, num_layers=2, num_heads=1, ffn_inner_dim=128)
    sd = make_synthetic_state_dict(cfg, seed=1)
    legacy = {}
    for k, v in sd.items():
        k2 = (k.replace("encoder.layers.", "layers.").replace("self_attn.output_proj", "self_attn.out_proj")
              .replace("ffn.inner_proj", "fc1").replace("ffn.output_proj", "fc2").replace("ffn_layer_norm", "final_layer_norm")
              .replace("encoder_frontend.embed.", "embed_tokens."))
        legacy[k2] = v
    legacy["embed_tokens.weight"] = _unswap(sd["encoder_frontend.embed.weight"])
    legacy["version"] = torch.tensor([3.0])
    out = convert_sonar_text_encoder_checkpoint({"state_dict": legacy})["model"]
    assert set(out) == set(sd)
    for k in sd:
        assert torch.equal(out[k], sd[k]), k
    same = {"model": sd}
    assert convert_sonar_text_encoder_checkpoint(same) is same  # already fairseq2: returned unchanged


def test_text_decoder_converter_roundtrip():
    cfg = OracleDecoderConfig(model_dim=64, vocab_size=40, num_layers=2, num_heads=1, ffn_inner_dim=128)
    sd = make_synthetic_decoder_state_dict(cfg, seed=1)
    legacy = {}
    for k, v in sd.items():
        if k == "final_proj.weight":
            continue
        k2 = (k.replace("decoder.layers.", "layers.").replace("encoder_decoder_attn_layer_norm", "encoder_attn_layer_norm")
              .replace("encoder_decoder_attn.", "encoder_attn.").replace(".output_proj", ".out_proj")
              .replace("ffn.inner_proj", "fc1").replace("ffn.out_proj", "fc2").replace("ffn_layer_norm", "final_layer_norm")
              .replace("decoder_frontend.embed.", "embed_tokens.").replace("decoder.layer_norm.", "layer_norm."))
        legacy[k2] = v
    legacy["embed_tokens.weight"] = _unswap(sd["decoder_frontend.embed.weight"])
    out = convert_sonar_text_decoder_checkpoint({"state_dict": legacy})["model"]
    assert set(out) == set(sd)
    for k in sd:
        assert torch.equal(out[k], sd[k]), k


def _legacy_speech_name(k: str) -> str:
    """fairseq2 name -> the fairseq name it came from (inverse of sonar_speech/handler.py:63-108)."""
    if k.startswith("encoder_frontend.post_extract_layer_norm."):
        return k.replace("encoder_frontend.post_extract_layer_norm.", "encoder.w2v_model.layer_norm.")
    if k.startswith("encoder_frontend.model_dim_proj."):
        return k.replace("encoder_frontend.model_dim_proj.", "encoder.w2v_model.post_extract_proj.")
    if k.startswith("layer_norm."):
        return "encoder.w2v_model.encoder." + k
    if k.startswith("encoder_pooler.decoder_frontend.embed."):
        return k.replace("encoder_pooler.decoder_frontend.embed.", "decoder.embed_tokens.")
    if k == "encoder_pooler.projection_out.weight":
        return "decoder.embed_out"
    if k.startswith("encoder_pooler.decoder.layers."):
        return (k.replace("encoder_pooler.decoder.layers.", "decoder.layers.")
                .replace("encoder_decoder_attn_layer_norm", "encoder_attn_layer_norm")
                .replace("encoder_decoder_attn.", "encoder_attn.").replace(".output_proj", ".out_proj")
                .replace("ffn.inner_proj", "fc1").replace("ffn.out_proj", "fc2").replace("ffn_layer_norm", "final_layer_norm"))
    assert k.startswith("encoder.layers.")
    n, rest = k[len("encoder.layers."):].split(".", 1)
    rest = (rest.replace("conv.batch_norm", "conv_module.batch_norm").replace("conv.depthwise_conv", "conv_module.depthwise_conv")
            .replace("conv_layer_norm", "conv_module.layer_norm").replace("conv.pointwise_conv", "conv_module.pointwise_conv")
            .replace("ffn1_layer_norm", "ffn1.layer_norm").replace("ffn2_layer_norm", "ffn2.layer_norm")
            .replace("ffn1.inner_proj", "ffn1.w_1").replace("ffn1.output_proj", "ffn1.w_2")
            .replace("ffn2.inner_proj", "ffn2.w_1").replace("ffn2.output_proj", "ffn2.w_2")
            .replace("self_attn.q_proj", "self_attn.linear_q").replace("self_attn.k_proj", "self_attn.linear_k")
            .replace("self_attn.v_proj", "self_attn.linear_v").replace("self_attn.output_proj", "self_attn.linear_out")
            .replace("self_attn.sdpa.r_proj", "self_attn.linear_pos").replace("self_attn.sdpa.u_bias", "self_attn.pos_bias_u")
            .replace("self_attn.sdpa.v_bias", "self_attn.pos_bias_v"))
    if rest.startswith("layer_norm."):
        rest = "final_" + rest
    return f"encoder.w2v_model.encoder.layers.{n}.{rest}"


def test_speech_converter_roundtrip():
    cfg = OracleSpeechConfig(model_dim=64, num_layers=2, num_heads=1, ffn_inner_dim=128, pooler_layers=1, pooler_heads=1,
                             pooler_ffn_inner_dim=128, pooler_vocab=16)
    sd = make_synthetic_speech_state_dict(cfg, seed=1)
    legacy = {_legacy_speech_name(k): v for k, v in sd.items()}
    assert len(legacy) == len(sd)
    legacy["encoder.w2v_model.mask_emb"] = torch.zeros(3)  # dropped by the converter (handler.py:55-56)
    out = convert_sonar_speech_checkpoint({"model": legacy})["model"]
    assert set(out) == set(sd), sorted(set(out) ^ set(sd))[:8]
    for k in sd:
        assert torch.equal(out[k], sd[k]), k
