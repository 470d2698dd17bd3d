"""Host-side mirror of the reference text encoder model, backed by the CUDA engine.

``B200TextEncoderModel`` is what a user passes as ``encoder=`` to
``TextToEmbeddingModelPipeline`` in place of fairseq2's
``SonarTextTransformerEncoderModel`` (``sonar/models/sonar_text/model.py:30-143``).
It exposes exactly the seam the pipeline touches:

* ``.eval()``, ``.dtype`` (``sonar/models/encoder_model.py:56-58``)
* ``.encoder_frontend.pos_encoder.max_seq_len`` (``sonar/inference_pipelines/text.py:202``)
* ``__call__(SequenceBatch) -> SonarEncoderOutput`` (``text.py:244-245``)

All arithmetic happens in ``libsonar_b200.so`` (``sb_encoder_forward``); this file
only repacks weights (bf16, fused QKV) and owns device buffers.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from enum import Enum
from pathlib import Path
from typing import Dict, List, Optional, Union

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from .sequence import PaddingMask, SequenceBatch, SonarEncoderOutput


class Pooling(Enum):
    """Same members/values as the reference enum (``model.py:23-27``)."""

    MAX = 1
    MEAN = 2
    LAST = 3
    ATTENTION = 4


@dataclass
class VocabularyInfo:
    size: int
    unk_idx: Optional[int] = None
    bos_idx: Optional[int] = None
    eos_idx: Optional[int] = None
    pad_idx: Optional[int] = None


@dataclass
class SonarTextEncoderConfig:
    """Field-for-field mirror of the reference dataclass
    (``sonar/models/sonar_text/config.py:14-84``); defaults = arch ``basic`` (``:92-116``)."""

    model_dim: int = 1024
    max_seq_len: int = 512
    vocab_info: VocabularyInfo = field(
        default_factory=lambda: VocabularyInfo(size=256206, unk_idx=1, bos_idx=2, eos_idx=3, pad_idx=1))
    num_encoder_layers: int = 24
    num_decoder_layers: int = 24
    num_encoder_attn_heads: int = 16
    num_decoder_attn_heads: int = 16
    ffn_inner_dim: int = 1024 * 8
    pooling: str = "mean"
    embedding_dim: Optional[int] = None
    decoder_ffn_inner_dim: Optional[int] = None
    activation_fn: str = "ReLU"
    layernorm_embedding: bool = False
    no_scale_embedding: bool = False
    no_token_positional_embeddings: bool = False
    learned_pos: bool = False
    emb_dropout_p: float = 0.1
    attention_dropout_p: float = 0.1
    activation_dropout_p: float = 0.1
    normalize_before: bool = False
    _from_fairseq: bool = True


def sonar_text_encoder_config(arch: str = "basic", **overrides) -> SonarTextEncoderConfig:
    """Named archs of ``register_sonar_text_encoder_configs`` (``config.py:87-127``)."""
    if arch == "basic":
        cfg = SonarTextEncoderConfig()
    elif arch == "small":  # config.py:118-127
        cfg = SonarTextEncoderConfig(
            vocab_info=VocabularyInfo(size=32005, unk_idx=1, bos_idx=2, eos_idx=3, pad_idx=1),
            num_encoder_layers=6, num_decoder_layers=6, ffn_inner_dim=1024 * 4)
    else:
        raise ValueError(f"unknown sonar text encoder arch {arch!r}")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown config field {k!r}")
        setattr(cfg, k, v)
    return cfg


def sinusoidal_position_table(num_pos: int, dim: int, legacy_pad_idx: int) -> Tensor:
    """fp32 table whose row t encodes position ``t + legacy_pad_idx + 1``: fairseq2
    ``SinusoidalPositionEncoder(_legacy_pad_idx=pad_idx)`` as built at ``factory.py:88-92``
    ([sin | cos] halves, frequencies ``exp(-j ln(1e4) / (dim/2 - 1))``)."""
    half = dim // 2
    start = legacy_pad_idx + 1
    steps = torch.arange(start, start + num_pos, dtype=torch.float32)
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))
    ang = steps[:, None] * freq[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).contiguous()


class _PosEncoderInfo:
    def __init__(self, max_seq_len: int) -> None:
        self.max_seq_len = max_seq_len


class _FrontendInfo:
    """Carries ``encoder_frontend.pos_encoder.max_seq_len`` (read at ``text.py:202``)."""

    def __init__(self, max_seq_len: int, model_dim: int) -> None:
        self.pos_encoder = _PosEncoderInfo(max_seq_len)
        self.model_dim = model_dim


def _check_supported(cfg: SonarTextEncoderConfig) -> None:
    bad = []
    if cfg.pooling.lower() not in ("mean", "max", "last"):
        bad.append(f"pooling={cfg.pooling!r} (attention pooling is not on the B200 hot path yet)")
    if cfg.activation_fn != "ReLU":
        bad.append(f"activation_fn={cfg.activation_fn!r}")
    if cfg.layernorm_embedding or cfg.no_token_positional_embeddings or cfg.learned_pos:
        bad.append("layernorm_embedding / no_token_positional_embeddings / learned_pos")
    if cfg.embedding_dim not in (None, cfg.model_dim):
        bad.append("embedding_dim != model_dim")
    if cfg.model_dim != 64 * cfg.num_encoder_attn_heads:
        bad.append("head_dim != 64")
    if bad:
        raise NotImplementedError("sonar_b200 text encoder does not support: " + "; ".join(bad))


class B200TextEncoderModel(torch.nn.Module):
    """SONAR text encoder (24-layer pre-LN Transformer + final LN + pooling) on sm_100a kernels."""

    def __init__(self, config: SonarTextEncoderConfig, state_dict: Dict[str, Tensor],
                 device: Union[str, torch.device] = "cuda", *, cta_group: int = 2, ln_fold: Union[bool, int] = False,
                 epi_groups: Optional[int] = None) -> None:
        """``ln_fold=True``: the engine folds every encoder-layer LayerNorm into the GEMMs around it (see
        ``SbEncoderConfig.ln_fold`` in ``include/sonar_b200.h``); the default runs the separate LayerNorm kernels, which is
        the faster schedule as measured.  ``epi_groups``: epilogue warpgroups per GEMM CTA -- 1 (six mainloop stages) is the
        default here: inside the power-capped 24-layer step it is 2 % faster than 2 (five stages), although 2 wins by 9-33 %
        when a GEMM is timed alone at boost clocks (``bench.py`` A/Bs all of these in every run, ``ab_schedule_variants``).
        The folded schedules exist with two warpgroups only, so ``None`` means 1 without and 2 with ``ln_fold``."""
        super().__init__()
        if epi_groups is None:
            epi_groups = 2 if int(ln_fold) else 1
        self.ln_fold = int(ln_fold)  # 0 = separate LayerNorm kernels, 1 = both folded, 2 = only the attention-block one
        self.epi_groups = int(epi_groups)
        _check_supported(config)
        self.config = config
        self.model_dim = config.model_dim
        self.pooling = getattr(Pooling, config.pooling.upper())
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("B200TextEncoderModel needs a CUDA device (there is no CPU path)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        pad_idx = config.vocab_info.pad_idx if config.vocab_info.pad_idx is not None else 1
        max_len = config.max_seq_len + (pad_idx + 1 if config._from_fairseq else 0)  # factory.py:53-59
        self.encoder_frontend = _FrontendInfo(max_len, config.model_dim)
        self._lib = _lib.load()

        sd = state_dict
        d, L = config.model_dim, config.num_encoder_layers

        def bf(t: Tensor) -> Tensor:
            return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()

        def f32(t: Tensor) -> Tensor:
            return t.detach().to(device=dev, dtype=torch.float32).contiguous()

        embed = sd["encoder_frontend.embed.weight"]
        if embed.shape != (config.vocab_info.size, d):
            raise ValueError(f"embedding shape {tuple(embed.shape)} != ({config.vocab_info.size}, {d})")
        self.register_buffer("embed", bf(embed), persistent=False)
        self.register_buffer("pos_table", sinusoidal_position_table(max_len, d, pad_idx).to(dev), persistent=False)
        self.register_buffer("final_ln_g", f32(sd["layer_norm.weight"]), persistent=False)
        self.register_buffer("final_ln_b", f32(sd["layer_norm.bias"]), persistent=False)
        self._layer_bufs: List[Dict[str, Tensor]] = []
        for i in range(L):
            p = f"encoder.layers.{i}."
            a = p + "self_attn."
            bufs = {
                "wqkv": bf(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0)),
                "bqkv": f32(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0)),
                "wo": bf(sd[a + "output_proj.weight"]), "bo": f32(sd[a + "output_proj.bias"]),
                "w1": bf(sd[p + "ffn.inner_proj.weight"]), "b1": f32(sd[p + "ffn.inner_proj.bias"]),
                "w2": bf(sd[p + "ffn.output_proj.weight"]), "b2": f32(sd[p + "ffn.output_proj.bias"]),
                "ln1_g": f32(sd[p + "self_attn_layer_norm.weight"]), "ln1_b": f32(sd[p + "self_attn_layer_norm.bias"]),
                "ln2_g": f32(sd[p + "ffn_layer_norm.weight"]), "ln2_b": f32(sd[p + "ffn_layer_norm.bias"]),
            }
            for k, v in bufs.items():
                self.register_buffer(f"l{i}_{k}", v, persistent=False)
            self._layer_bufs.append(bufs)

        cfg_c = _lib.SbEncoderConfig(
            model_dim=d, num_layers=L, num_heads=config.num_encoder_attn_heads, ffn_inner_dim=config.ffn_inner_dim,
            vocab_size=config.vocab_info.size, pos_rows=max_len, pooling=self.pooling.value, ln_eps=1e-5,
            embed_scale=1.0 if config.no_scale_embedding else math.sqrt(d), cta_group=cta_group, num_sms=0,
            ln_fold=int(ln_fold), epi_groups=int(epi_groups))
        layers_c = (_lib.SbLayerWeights * max(L, 1))()
        for i, bufs in enumerate(self._layer_bufs):
            for k, v in bufs.items():
                setattr(layers_c[i], k, v.data_ptr())
        w_c = _lib.SbEncoderWeights(embed=self.embed.data_ptr(), pos_table=self.pos_table.data_ptr(),
                                    final_ln_g=self.final_ln_g.data_ptr(), final_ln_b=self.final_ln_b.data_ptr(),
                                    layers=layers_c)
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self._lib.sb_encoder_create(C.byref(cfg_c), C.byref(w_c), C.byref(handle)), "sb_encoder_create")
        self._handle = handle
        self._workspace: Optional[Tensor] = None
        self.return_encoded_seqs = False

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_checkpoint(cls, path: Union[str, Path], config: Optional[SonarTextEncoderConfig] = None,
                        device: Union[str, torch.device] = "cuda", **kw) -> "B200TextEncoderModel":
        """Load a fairseq2-layout checkpoint ``{"model": state_dict}``."""
        ckpt = torch.load(str(path), map_location="cpu", weights_only=True)
        sd = ckpt["model"] if "model" in ckpt else ckpt
        return cls(config or sonar_text_encoder_config("basic"), sd, device, **kw)

    # ------------------------------------------------------------------ nn.Module-ish surface
    @property
    def dtype(self) -> torch.dtype:
        """Compute dtype of the engine (bf16 operands, fp32 accumulation and residual stream)."""
        return torch.bfloat16

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            if getattr(self, "_handle", None):
                self._lib.sb_encoder_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _ensure_workspace(self, batch: int, tokens: int) -> Tensor:
        need = C.c_size_t()
        _lib.check(self._lib.sb_encoder_workspace_bytes(self._handle, batch, tokens, C.byref(need)),
                   "sb_encoder_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need.value:
            self._workspace = None  # release before growing
            self._workspace = torch.empty(int(need.value * 1.1) + 4096, dtype=torch.uint8, device=self.device)
        return self._workspace

    @torch.inference_mode()
    def forward(self, batch: SequenceBatch) -> SonarEncoderOutput:
        seqs = batch.seqs
        if seqs.dim() != 2:
            raise ValueError("expected token ids of shape [N, S]")
        if not seqs.is_cuda:
            seqs = seqs.to(self.device, non_blocking=True)
        if seqs.dtype != torch.int64:
            seqs = seqs.to(torch.int64)
        if seqs.stride(1) != 1:
            seqs = seqs.contiguous()
        n, s = seqs.shape
        pm = batch.padding_mask
        if pm is not None:
            lens_np = np.ascontiguousarray(pm.seq_lens_host, dtype=np.int32)  # one vectorised conversion, no per-item Python
            lens_c = lens_np.ctypes.data_as(C.POINTER(C.c_int32))
            tokens = int(lens_np.sum(dtype=np.int64))
        else:
            lens_c = None
            tokens = n * s
        ws = self._ensure_workspace(n, max(tokens, 1))
        out = torch.empty((n, self.model_dim), dtype=torch.float32, device=self.device)
        enc = (torch.empty((n, s, self.model_dim), dtype=torch.float32, device=self.device)
               if self.return_encoded_seqs else None)
        with torch.cuda.device(self.device):
            rc = self._lib.sb_encoder_forward(
                self._handle, seqs.data_ptr(), seqs.stride(0), lens_c, n, s, out.data_ptr(),
                enc.data_ptr() if enc is not None else None, ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "sb_encoder_forward")
        return SonarEncoderOutput(encoded_seqs=enc, sentence_embeddings=out, padding_mask=pm)

    def profile_ffn1(self, start: Optional[torch.cuda.Event], stop: Optional[torch.cuda.Event]) -> None:
        """Record `start`/`stop` (timing-enabled events that have been recorded once, so their handles exist) around
        the FFN inner-projection GEMM of the middle layer in every following forward; (None, None) switches it off."""
        _lib.check(self._lib.sb_encoder_profile_ffn1(self._handle, start.cuda_event if start is not None else None,
                                                     stop.cuda_event if stop is not None else None),
                   "sb_encoder_profile_ffn1")

    def check_inputs(self) -> None:
        """Raise ``ValueError`` if the last batch contained a token id outside the vocabulary."""
        if self._workspace is None:
            return
        rc = self._lib.sb_encoder_check_inputs(self._handle, self._workspace.data_ptr(),
                                               torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "sb_encoder_check_inputs")

    # ------------------------------------------------------------------ reference static API
    @staticmethod
    def static_pooling(seqs: Tensor, padding_mask: Optional[PaddingMask], pooling: Pooling) -> Tensor:
        """``SonarTextTransformerEncoderModel.static_pooling`` (``model.py:86-128``) on the GPU kernel.
        ``seqs`` is a padded CUDA tensor [N, S, D] with D a multiple of 128."""
        if pooling == Pooling.ATTENTION:
            raise NotImplementedError(pooling)
        if not seqs.is_cuda:
            raise RuntimeError("static_pooling runs on CUDA tensors only")
        n, s, d = seqs.shape
        lens = padding_mask.seq_lens_host if padding_mask is not None else [s] * n
        valid = (torch.arange(s)[None, :] < torch.tensor(lens)[:, None]).to(seqs.device)
        packed = seqs.float()[valid].contiguous()  # [T, D] (layout change only)
        cu = ops.cu_seqlens_of(lens).to(seqs.device)
        return ops.pool_packed(packed, cu, pooling.name).to(seqs.dtype)
