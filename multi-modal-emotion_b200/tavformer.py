"""Drop-in replacements for the two encoders of the reference's utils/TAVFormer.py, executed by the sm_100a kernel
library through ``engine.EncoderStackFn``.

* ``VideoMAEEncoder(config, num_layers)``  — reference utils/TAVFormer.py:171-223 (+ layer classes :230-439).
  Same constructor, ``forward`` signature and ``state_dict`` keys (``layer.{i}.layernorm_before.*``,
  ``.attention.attention.{query,key,value}.weight``, ``.attention.attention.{q_bias,v_bias}``,
  ``.attention.output.dense.*``, ``.layernorm_after.*``, ``.intermediate.dense.*``, ``.output.dense.*``).
  The reference adds ``attention_mask`` to the attention PROBABILITIES (:372-375); that is reproduced exactly as
  unmasked attention + a rank-1 fp32 term.
* ``TransformerEncoder(embed_dim, num_layers, expansion_factor, n_heads, dropout, early_div)`` — reference
  utils/TAVFormer.py:144-166 (+ :10-142): post-LN blocks, bias-free q/k/v, pre-softmax additive mask, and the
  scrambled head "concat" (:86).

The nn.Linear / nn.LayerNorm children are parameter containers only (so initialisation and checkpoints match the
reference); their own ``forward`` is never called."""
import torch
from torch import nn

from . import engine
from .engine import LayerSpec


class _SelfAttentionParams(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0 and not hasattr(config, "embedding_size"):
            raise ValueError(
                f"The hidden size {config.hidden_size,} is not a multiple of the number of attention "
                f"heads {config.num_attention_heads}.")
        H = config.hidden_size
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = H // config.num_attention_heads
        self.all_head_size = H
        self.query = nn.Linear(H, H, bias=False)
        self.key = nn.Linear(H, H, bias=False)
        self.value = nn.Linear(H, H, bias=False)
        if config.qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(H))
            self.v_bias = nn.Parameter(torch.zeros(H))
        else:
            self.q_bias = None
            self.v_bias = None


class _Dense(nn.Module):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.dense = nn.Linear(n_in, n_out)


class _AttentionParams(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = _SelfAttentionParams(config)
        self.output = _Dense(config.hidden_size, config.hidden_size)


class VideoMAELayer(nn.Module):
    """Parameter container of one fusion layer (reference utils/TAVFormer.py:230-241)."""

    def __init__(self, config):
        super().__init__()
        self.layernorm_before = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.attention = _AttentionParams(config)
        self.layernorm_after = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.intermediate = _Dense(config.hidden_size, config.intermediate_size)
        self.output = _Dense(config.intermediate_size, config.hidden_size)

    def slots(self):
        a = self.attention.attention
        return [self.layernorm_before.weight, self.layernorm_before.bias, a.query.weight, a.key.weight, a.value.weight,
                a.q_bias, None, a.v_bias, self.attention.output.dense.weight, self.attention.output.dense.bias,
                self.layernorm_after.weight, self.layernorm_after.bias, self.intermediate.dense.weight,
                self.intermediate.dense.bias, self.output.dense.weight, self.output.dense.bias]


class VideoMAEEncoder(nn.Module):
    def __init__(self, config, num_layers: int) -> None:
        super().__init__()
        self.config = config
        if config.hidden_size // config.num_attention_heads != 64:
            raise ValueError("the sm_100a attention kernel supports head_dim 64 only")
        if not isinstance(config.hidden_act, str):
            # the reference maps every string activation to exact-erf nn.GELU() (utils/TAVFormer.py:397-398)
            raise NotImplementedError("only the exact-erf GELU FFN is implemented by the kernel path")
        self.layer = nn.ModuleList([VideoMAELayer(config) for _ in range(num_layers)])
        self.gradient_checkpointing = False
        self._shadows = [engine.LayerShadow() for _ in range(num_layers)]

    def _spec(self, masked):
        c = self.config
        return LayerSpec(hidden=c.hidden_size, heads=c.num_attention_heads, inter=c.intermediate_size, pre_ln=True,
                         eps=c.layer_norm_eps, mask_mode="rank1" if masked else "none")

    def forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions: bool = False,
                output_hidden_states: bool = False, return_dict: bool = True):
        if head_mask is not None or output_attentions:
            raise NotImplementedError("head_mask / output_attentions are not produced by the fused attention kernel")
        B, S, _ = hidden_states.shape
        mask2d = None
        if attention_mask is not None:
            mask2d = attention_mask.to(device=hidden_states.device, dtype=torch.float32).reshape(B, S)
        spec = self._spec(mask2d is not None)
        x = hidden_states.float()
        all_hidden = () if output_hidden_states else None
        if output_hidden_states:
            for i, lyr in enumerate(self.layer):
                all_hidden = all_hidden + (x,)
                x = engine.run_stack(spec, self._shadows[i:i + 1], x, mask2d, [lyr.slots()])
            all_hidden = all_hidden + (x,)
        else:
            x = engine.run_stack(spec, self._shadows, x, mask2d, [lyr.slots() for lyr in self.layer])
        if not return_dict:
            return tuple(v for v in [x, all_hidden, None] if v is not None)
        return x


class _MHAParams(nn.Module):
    def __init__(self, embed_dim=768, n_heads=12):
        super().__init__()
        self.embed_dim = embed_dim
        self.n_heads = n_heads
        self.single_head_dim = int(embed_dim / n_heads)
        self.query_matrix = nn.Linear(embed_dim, embed_dim, bias=False)
        self.key_matrix = nn.Linear(embed_dim, embed_dim, bias=False)
        self.value_matrix = nn.Linear(embed_dim, embed_dim, bias=False)
        self.out = nn.Linear(embed_dim, embed_dim)


class TransformerBlock(nn.Module):
    """Parameter container of one post-LN block (reference utils/TAVFormer.py:93-118)."""

    def __init__(self, embed_dim, expansion_factor=4, n_heads=12, dropout=0.2):
        super().__init__()
        self.dropout = dropout
        self.attention = _MHAParams(embed_dim, n_heads)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.feed_forward = nn.Sequential(nn.Dropout(dropout), nn.Linear(embed_dim, expansion_factor * embed_dim),
                                          nn.GELU(), nn.Linear(expansion_factor * embed_dim, embed_dim))
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(embed_dim)

    def slots(self):
        a = self.attention
        return [self.norm1.weight, self.norm1.bias, a.query_matrix.weight, a.key_matrix.weight, a.value_matrix.weight,
                None, None, None, a.out.weight, a.out.bias, self.norm2.weight, self.norm2.bias,
                self.feed_forward[1].weight, self.feed_forward[1].bias, self.feed_forward[3].weight,
                self.feed_forward[3].bias]


class TransformerEncoder(nn.Module):
    def __init__(self, embed_dim, num_layers=2, expansion_factor=4, n_heads=12, dropout=0.2, early_div=False):
        super().__init__()
        if embed_dim // n_heads != 64:
            raise ValueError("the sm_100a attention kernel supports head_dim 64 only")
        self.early_div = early_div  # scaling Q before QK^T vs the scores after: identical up to fp32 rounding
        self.embed_dim, self.n_heads, self.expansion_factor, self.p = embed_dim, n_heads, expansion_factor, dropout
        self.layers = nn.ModuleList([TransformerBlock(embed_dim, expansion_factor, n_heads, dropout) for _ in range(num_layers)])
        self._shadows = [engine.LayerShadow() for _ in range(num_layers)]
        TransformerEncoder._instances += 1
        self._salt = TransformerEncoder._instances     # construction order: reproducible dropout streams per encoder

    _instances = 0

    def forward(self, x, attention_mask=None):
        B, S, _ = x.shape
        mask2d = None
        if attention_mask is not None:
            if attention_mask.shape[-2] != 1:
                raise NotImplementedError("only key-padding masks of shape [B,1,1,S] are supported")
            mask2d = attention_mask.to(device=x.device, dtype=torch.float32).reshape(B, S)
        spec = LayerSpec(hidden=self.embed_dim, heads=self.n_heads, inter=self.expansion_factor * self.embed_dim,
                         pre_ln=False, eps=1e-5, mask_mode="key_bias" if mask2d is not None else "none",
                         scrambled_concat=True,
                         # training mode: the block's three nn.Dropout(p) (reference utils/TAVFormer.py:107,111,117,130-141)
                         dropout=float(self.p) if (self.training and self.p > 0) else 0.0,
                         dropout_salt=self._salt)
        return engine.run_stack(spec, self._shadows, x.float(), mask2d, [blk.slots() for blk in self.layers])
