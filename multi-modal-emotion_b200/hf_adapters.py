"""Runs the transformer layers of the three HF encoders the reference instantiates (RobertaModel, Wav2Vec2Model,
VideoMAEModel — reference models/tav.py:257,259,263,438,455,456) on the same sm_100a kernel set as the fusion block.

The HF modules are kept as parameter containers (so ``state_dict`` keys ``bert.*`` / ``wav2vec2.*`` / ``videomae.*``
are unchanged); their encoder-layer ``forward`` is bypassed: per-layer parameters are read in ``engine.PARAM_SLOTS``
order and the whole stack runs through ``engine.run_stack``.  The non-transformer front-ends (conv feature extractor,
positional conv, patch-embedding conv, token/position embeddings) still execute as ATen/cuDNN ops — SURVEY.md §8(f)
lists them as the "next" rows.  Layer families and their LayerNorm placement follow SURVEY.md Appendix A."""
import torch

from . import engine, frontends
from .engine import LayerSpec


def _shadows(module, n):
    sh = getattr(module, "_tavk_shadows", None)
    if sh is None or len(sh) != n:
        sh = [engine.LayerShadow() for _ in range(n)]
        module._tavk_shadows = sh
    return sh


def _check_head_dim(hidden, heads):
    if hidden // heads != 64:
        raise ValueError("the sm_100a attention kernel supports head_dim 64 only (got %d/%d)" % (hidden, heads))


# ------------------------------------------------------------------------------------------------ VideoMAE
def videomae_layer_slots(lyr):
    a = lyr.attention.attention
    return [lyr.layernorm_before.weight, lyr.layernorm_before.bias, a.query.weight, a.key.weight, a.value.weight,
            a.q_bias, None, a.v_bias, lyr.attention.output.dense.weight, lyr.attention.output.dense.bias,
            lyr.layernorm_after.weight, lyr.layernorm_after.bias, lyr.intermediate.dense.weight,
            lyr.intermediate.dense.bias, lyr.output.dense.weight, lyr.output.dense.bias]


def run_videomae(model, pixel_values, bool_masked_pos=None, keep_count=None):
    """VideoMAEModel.forward(pixel_values, bool_masked_pos)[0] (reference call site models/tav.py:480): patch
    embedding + sinusoid positions + token drop, then N pre-LN layers with mask-free attention, then the optional
    final LayerNorm (absent when use_mean_pooling=True)."""
    c = model.config
    _check_head_dim(c.hidden_size, c.num_attention_heads)
    x = video_embeddings(model.embeddings, pixel_values, bool_masked_pos, keep_count)
    spec = LayerSpec(hidden=c.hidden_size, heads=c.num_attention_heads, inter=c.intermediate_size, pre_ln=True,
                     eps=c.layer_norm_eps, mask_mode="none")
    layers = model.encoder.layer
    x = engine.run_stack(spec, _shadows(model, len(layers)), x.float(), None, [videomae_layer_slots(l) for l in layers])
    if model.layernorm is not None:
        x = engine.layer_norm(x, model.layernorm.weight, model.layernorm.bias, c.layer_norm_eps)
    return x


def video_embeddings(emb, pixel_values, bool_masked_pos, keep_count=None):
    """VideoMAEEmbeddings.forward on the tcgen05 GEMM (frontends.video_embeddings): gather the kept tokens first,
    project, add the sinusoid rows in the epilogue."""
    return frontends.video_embeddings(emb, pixel_values, bool_masked_pos, keep_count)


# ------------------------------------------------------------------------------------------------ RoBERTa
def roberta_layer_slots(lyr):
    a = lyr.attention.self
    o = lyr.attention.output
    return [o.LayerNorm.weight, o.LayerNorm.bias, a.query.weight, a.key.weight, a.value.weight, a.query.bias, a.key.bias,
            a.value.bias, o.dense.weight, o.dense.bias, lyr.output.LayerNorm.weight, lyr.output.LayerNorm.bias,
            lyr.intermediate.dense.weight, lyr.intermediate.dense.bias, lyr.output.dense.weight, lyr.output.dense.bias]


def run_roberta(model, input_ids, attention_mask):
    """RobertaModel(input_ids, attention_mask, return_dict=False) -> (sequence_output, pooled_output)
    (reference call site models/tav.py:485): embeddings, N post-LN layers with an additive key-padding bias,
    pooler tanh(W h[:,0] + b)."""
    c = model.config
    _check_head_dim(c.hidden_size, c.num_attention_heads)
    x = engine.roberta_embeddings(model.embeddings, input_ids)
    B, S, _ = x.shape
    bias = None
    if attention_mask is not None:
        bias = (1.0 - attention_mask.to(device=x.device, dtype=torch.float32).reshape(B, S)) * -1.0e9
    spec = LayerSpec(hidden=c.hidden_size, heads=c.num_attention_heads, inter=c.intermediate_size, pre_ln=False,
                     eps=c.layer_norm_eps, mask_mode="key_bias" if bias is not None else "none")
    layers = model.encoder.layer
    x = engine.run_stack(spec, _shadows(model, len(layers)), x.float(), bias, [roberta_layer_slots(l) for l in layers])
    pooled = None
    if model.pooler is not None:
        pooled = torch.tanh(engine.small_linear(x[:, 0].contiguous(), model.pooler.dense.weight, model.pooler.dense.bias))
    return x, pooled


# ------------------------------------------------------------------------------------------------ Wav2Vec2
def wav2vec2_layer_slots(lyr):
    a = lyr.attention
    ff = lyr.feed_forward
    return [lyr.layer_norm.weight, lyr.layer_norm.bias, a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.q_proj.bias,
            a.k_proj.bias, a.v_proj.bias, a.out_proj.weight, a.out_proj.bias, lyr.final_layer_norm.weight,
            lyr.final_layer_norm.bias, ff.intermediate_dense.weight, ff.intermediate_dense.bias, ff.output_dense.weight,
            ff.output_dense.bias]


def feature_projection(model, feats):
    """HF Wav2Vec2FeatureProjection.forward (LayerNorm over the conv channels + Linear C -> H; dropout p = 0 in eval) on
    the LayerNorm kernel and the tcgen05 GEMM (reference models/tav.py:356 and inside wav2vec2(...) at :476)."""
    fp = model.feature_projection
    if fp.training and fp.dropout.p > 0:
        raise NotImplementedError("feature_projection dropout in training mode (the reference keeps the HF models in eval)")
    h = engine.layer_norm(feats, fp.layer_norm.weight, fp.layer_norm.bias, fp.layer_norm.eps)
    return engine.linear_bf16(h, fp.projection.weight, fp.projection.bias)


def wav2vec2_front(model, wav):
    """feature_extractor (7 x Conv1d) -> transpose -> feature_projection (LN + Linear): [B,L] -> [B,Ta,H]."""
    feats = frontends.feature_extractor_cl(model, wav)      # channels-last [B, frames, C]
    return feature_projection(model, feats)


def wav2vec2_encoder(model, hidden):
    """Wav2Vec2Encoder / Wav2Vec2EncoderStableLayerNorm .forward without attention mask (the reference calls
    wav2vec2(audio) with no mask, models/tav.py:476; HF sub-models stay in eval mode)."""
    c = model.config
    _check_head_dim(c.hidden_size, c.num_attention_heads)
    enc = model.encoder
    hidden = hidden + frontends.pos_conv_embed(enc.pos_conv_embed, hidden)
    stable = bool(c.do_stable_layer_norm)
    if not stable:
        hidden = engine.layer_norm(hidden, enc.layer_norm.weight, enc.layer_norm.bias, c.layer_norm_eps)
    spec = LayerSpec(hidden=c.hidden_size, heads=c.num_attention_heads, inter=c.intermediate_size, pre_ln=stable,
                     eps=c.layer_norm_eps, mask_mode="none")
    layers = enc.layers
    hidden = engine.run_stack(spec, _shadows(model, len(layers)), hidden.float(), None,
                              [wav2vec2_layer_slots(l) for l in layers])
    if stable:
        hidden = engine.layer_norm(hidden, enc.layer_norm.weight, enc.layer_norm.bias, c.layer_norm_eps)
    return hidden


def run_wav2vec2(model, wav):
    """Wav2Vec2Model(wav)[0] (reference call site models/tav.py:476)."""
    return wav2vec2_encoder(model, wav2vec2_front(model, wav))
