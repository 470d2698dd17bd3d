"""`NllbTokenizer` (sonar_b200/tokenizer.py) on a SentencePiece model trained inside the test (pip `sentencepiece`; no model file
ships and none can be downloaded).  Pins the NLLB id layout SURVEY App. A.2 / F1 describes and the pipelines rely on
(/root/reference/sonar/inference_pipelines/text.py:85-87,199-201,241): ids pad=0 unk=1 bos=2 eos=3, SentencePiece pieces shifted
by one, `__lang__` control symbols after the pieces; source encoding [__lang__] + pieces + [</s>]; target-mode prompt
[</s>, __lang__]."""

import pytest
import torch

spm = pytest.importorskip("sentencepiece")

LANGS = ["eng_Latn", "fra_Latn", "deu_Latn"]


@pytest.fixture(scope="module")
def spm_model(tmp_path_factory):
    d = tmp_path_factory.mktemp("spm")
    words = ["the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog", "sonar", "embeds", "sentences", "into", "vectors",
             "speech", "text", "translation", "model", "encoder", "decoder", "language", "hello", "world", "again", "and"]
    g = torch.Generator().manual_seed(0)
    lines = [" ".join(words[int(i)] for i in torch.randint(0, len(words), (int(torch.randint(3, 12, (1,), generator=g)),), generator=g))
             for _ in range(3000)]
    corpus = d / "corpus.txt"
    corpus.write_text("\n".join(lines) + "\n")
    prefix = str(d / "toy")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=prefix, vocab_size=300, model_type="bpe",
                                   character_coverage=1.0, hard_vocab_limit=False, minloglevel=2)
    return prefix + ".model"


def test_nllb_layout_source_and_target_modes(spm_model):
    from sonar_b200.tokenizer import NllbTokenizer

    sp = spm.SentencePieceProcessor(model_file=spm_model)
    n = sp.get_piece_size()
    assert (sp.unk_id(), sp.bos_id(), sp.eos_id()) == (0, 1, 2)  # the SPM-side ids the +1 shift assumes
    tok = NllbTokenizer(spm_model, LANGS)
    vi = tok.vocab_info
    assert (vi.pad_idx, vi.unk_idx, vi.bos_idx, vi.eos_idx) == (0, 1, 2, 3)
    assert vi.size == n + 1 + len(LANGS) + 3  # pieces shifted by one, then __lang__ symbols, then 3 data-tag symbols

    text = "the quick brown fox jumps over the lazy dog"
    enc = tok.create_encoder(lang="fra_Latn", device="cpu")
    ids = enc(text)
    assert ids.dtype == torch.int64 and ids.dim() == 1
    lang_id = n + 1 + LANGS.index("fra_Latn")
    assert int(ids[0]) == lang_id and int(ids[-1]) == 3
    assert ids[1:-1].tolist() == [i + 1 for i in sp.encode(text)]
    assert enc.prefix_indices.tolist() == [lang_id] and enc.suffix_indices.tolist() == [3]
    assert int(ids[1:-1].min()) >= 4 and int(ids[1:-1].max()) <= n  # pieces never collide with the control ids

    tgt = tok.create_encoder(task="translation", lang="deu_Latn", mode="target")
    tids = tgt(text)
    deu = n + 1 + LANGS.index("deu_Latn")
    assert tids[:2].tolist() == [3, deu] and int(tids[-1]) == 3   # decoder input: </s> __lang__ pieces </s>
    assert tgt.prefix_indices.tolist() == [3, deu]                # = the generator prompt
    assert tids[2:-1].tolist() == ids[1:-1].tolist()

    dec = tok.create_decoder()
    assert dec(ids) == text and dec(tids) == text  # control symbols are dropped
    assert dec(torch.tensor([lang_id, 3])) == ""

    with pytest.raises(ValueError):
        tok.create_encoder(lang="xxx_Latn")
    with pytest.raises(ValueError):
        tok.create_encoder()


def test_nllb_tokenizer_drives_the_text_pipeline_batcher(spm_model):
    """The pipeline's host stages (truncate -> dynamic_bucket -> collate with the TOKENIZER's pad id 0, text.py:241) on real
    SentencePiece ids: ragged batch, right padding, lengths."""
    from sonar_b200.batching import collate, dynamic_bucket
    from sonar_b200.tokenizer import NllbTokenizer

    tok = NllbTokenizer(spm_model, LANGS)
    enc = tok.create_encoder(lang="eng_Latn")
    sents = ["hello world", "the quick brown fox jumps over the lazy dog again and again", "sonar"]
    toks = [enc(s) for s in sents]
    (group,) = list(dynamic_bucket(iter(toks), 2 ** 31, len, max_num_examples=5))
    ids, lens, ragged = collate(group, tok.vocab_info.pad_idx)
    assert ragged and lens == [int(t.numel()) for t in toks] and ids.shape == (3, max(lens))
    for i, t in enumerate(toks):
        assert torch.equal(ids[i, : lens[i]], t) and bool((ids[i, lens[i]:] == 0).all())
    (g1, g2) = list(dynamic_bucket(iter(toks), 5, len))  # batch_max_tokens=5: the crossing example is included
    assert [len(g1), len(g2)] == [2, 1]
