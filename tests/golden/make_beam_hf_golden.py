"""Generate tests/golden/beam_hf_small.pt: beam-search outputs of HuggingFace `generate(num_beams=...)` on a tiny
`M2M100ForConditionalGeneration` whose decoder weights are exported under the fairseq2 names, to pin the beam-search BOOKKEEPING
of `oracle/text_decoder.py::beam_search` (and through it the product's `sonar_b200/generation.py`) against an independent
implementation in the regime where HuggingFace's algorithm and fairseq2's coincide:

    python tests/golden/make_beam_hf_golden.py

The decoder is conditioned on ONE encoder position per sentence (`encoder_outputs=[N,1,d]`), exactly how
EmbeddingToTextModelPipeline drives it (sonar/models/sonar_translation/model.py:48-53).

Case A  `length_penalty=0`, prompt = decoder_input_ids [</s>, lang]: raw cumulative log-probs; HF does not score the prompt, so
        oracle score - log P(lang | </s>) == HF score.
Case B  `length_penalty=1`, decoder_input_ids [</s>] + `forced_bos_token_id=lang`: HF then normalises by the generated length
        INCLUDING the forced language token = P + g, which is fairseq2's divisor (seq_len - 1); HF gives the forced token
        log-prob 0, so the oracle is run with `score_prompt=False`.
Both: `early_stopping=True` (close a sentence at `num_beams` finished hypotheses), PAD and id 1 suppressed (HF's position ids
treat id 1 = its padding_idx specially).  Sentences for which ANY HuggingFace hypothesis reaches `max_new_tokens` are left out
of that case (`sentences` lists the kept indices): at the length limit HF finalises the unfinished beams as they are while
fairseq2 forces EOS as the last token -- a real divergence of the two algorithms, not part of what is pinned here."""

import os

import torch
from transformers import M2M100Config, M2M100ForConditionalGeneration
from transformers.modeling_outputs import BaseModelOutput

D, L, H, F_, V = 32, 2, 2, 64, 48
EOS, LANG, PAD = 3, 17, 0
N, MAX_NEW = 20, 24

NAME_MAP = {"self_attn.q_proj": "self_attn.q_proj", "self_attn.k_proj": "self_attn.k_proj",
            "self_attn.v_proj": "self_attn.v_proj", "self_attn.out_proj": "self_attn.output_proj",
            "self_attn_layer_norm": "self_attn_layer_norm",
            "encoder_attn.q_proj": "encoder_decoder_attn.q_proj", "encoder_attn.k_proj": "encoder_decoder_attn.k_proj",
            "encoder_attn.v_proj": "encoder_decoder_attn.v_proj", "encoder_attn.out_proj": "encoder_decoder_attn.output_proj",
            "encoder_attn_layer_norm": "encoder_decoder_attn_layer_norm",
            "fc1": "ffn.inner_proj", "fc2": "ffn.output_proj", "final_layer_norm": "ffn_layer_norm"}


def build():
    torch.manual_seed(2024)
    cfg = M2M100Config(vocab_size=V, d_model=D, encoder_layers=1, decoder_layers=L, encoder_attention_heads=H,
                       decoder_attention_heads=H, encoder_ffn_dim=F_, decoder_ffn_dim=F_, max_position_embeddings=64,
                       pad_token_id=1, eos_token_id=EOS, bos_token_id=2, decoder_start_token_id=EOS, dropout=0.0,
                       attention_dropout=0.0, activation_dropout=0.0, activation_function="relu", scale_embedding=True,
                       decoder_layerdrop=0.0, tie_word_embeddings=True)
    cfg._attn_implementation = "eager"
    model = M2M100ForConditionalGeneration(cfg).eval().float()
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(torch.randn_like(p) * 0.3 + (1.0 if name.endswith("layer_norm.weight") else 0.0))
        # EOS competes at every step so hypotheses finish well before the length limit: its (tied) embedding row points along
        # the final LayerNorm's bias direction, which every decoder state shares
        model.model.shared.weight[EOS] = model.model.shared.weight[EOS] * 0.5 + 0.9 * model.model.decoder.layer_norm.bias
    model.tie_weights()
    return cfg, model


def export_state_dict(model):
    dec = model.model.decoder.state_dict()
    sd = {"decoder_frontend.embed.weight": model.model.shared.weight.detach().clone()}
    for i in range(L):
        for a, b in NAME_MAP.items():
            for wb in ("weight", "bias"):
                sd[f"decoder.layers.{i}.{b}.{wb}"] = dec[f"layers.{i}.{a}.{wb}"].clone()
    sd["decoder.layer_norm.weight"] = dec["layer_norm.weight"].clone()
    sd["decoder.layer_norm.bias"] = dec["layer_norm.bias"].clone()
    sd["final_proj.weight"] = sd["decoder_frontend.embed.weight"]
    return sd


def run(model, enc, beams, length_penalty, forced):
    n = enc.shape[0]
    kw = dict(encoder_outputs=BaseModelOutput(last_hidden_state=enc), num_beams=beams, num_return_sequences=beams,
              early_stopping=True, length_penalty=length_penalty, max_new_tokens=MAX_NEW + (1 if forced else 0), do_sample=False,
              suppress_tokens=[PAD, 1], output_scores=True, return_dict_in_generate=True, use_cache=True,
              pad_token_id=PAD, eos_token_id=EOS)
    if forced:
        ids = torch.full((n, 1), EOS)
        out = model.generate(decoder_input_ids=ids, forced_bos_token_id=LANG, **kw)
    else:
        ids = torch.tensor([[EOS, LANG]]).repeat(n, 1)
        out = model.generate(decoder_input_ids=ids, **kw)
    seqs = out.sequences.view(n, beams, -1)
    scores = out.sequences_scores.view(n, beams)
    hyps = []
    for i in range(n):
        row = []
        for b in range(beams):
            s = seqs[i, b].tolist()[2:]  # strip </s>, lang
            if EOS in s:
                s = s[: s.index(EOS) + 1]
            row.append((float(scores[i, b]), s))
        hyps.append(row)
    return hyps


def main() -> None:
    cfg, model = build()
    g = torch.Generator().manual_seed(7)
    enc = torch.randn(N, 1, D, generator=g) * 0.5
    cases = {}
    for beams in (2, 3, 5):
        for tag, lp, forced in (("A", 0.0, False), ("B", 1.0, True)):
            hyps = run(model, enc, beams, lp, forced=forced)
            keep = [i for i, row in enumerate(hyps) if all(len(h[1]) < MAX_NEW and h[1][-1] == EOS for h in row)]
            cases[f"{tag}_beam{beams}"] = {"beams": beams, "length_penalty": lp, "score_prompt": not forced, "sentences": keep,
                                           "hyps": [hyps[i] for i in keep]}
    here = os.path.dirname(os.path.abspath(__file__))
    torch.save({"config": dict(model_dim=D, vocab_size=V, max_seq_len=61, pad_idx=1, num_layers=L, num_heads=H, ffn_inner_dim=F_),
                "state_dict": export_state_dict(model), "encoder_output": enc, "prompt": [EOS, LANG], "max_gen_len": MAX_NEW,
                "cases": cases, "generator": "transformers M2M100ForConditionalGeneration.generate"},
               os.path.join(here, "beam_hf_small.pt"))
    import collections
    lens = [len(h[1]) for c in cases.values() for row in c["hyps"] for h in row]
    print("wrote beam_hf_small.pt; kept sentences per case", {k: len(c["sentences"]) for k, c in cases.items()})
    print("hypothesis lengths (incl. EOS):", sorted(collections.Counter(lens).items()))


if __name__ == "__main__":
    main()
