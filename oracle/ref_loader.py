"""TEST INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference modules from /root/reference (authoring container only:
the path does not exist on the GPU box) so that golden vectors can be generated from the real code.

The reference imports packages that are absent here (pytorchvideo, torchmetrics) at module scope and calls
``from_pretrained`` (no network / no HF cache), so stubs are registered and the HF factory functions are re-pointed at
config-based random-init constructors BEFORE the import (recipe of SURVEY.md §8c).  Nothing is copied from the
reference; it is executed where it lies."""
import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _is_ref(root):
    return bool(root) and os.path.isfile(os.path.join(root, "models", "tav.py")) and os.path.isfile(os.path.join(root, "utils", "TAVFormer.py"))


# where the unmodified reference may lie: an explicit override, the authoring container's read-only checkout, or
# baseline/_ref — the git-ignored place the driver (or an operator) may drop the reference tree for the bench's reference arm
REF_ROOT = next((r for r in (os.environ.get("TAV_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref"))
                 if _is_ref(r)), "/root/reference")


def available():
    return _is_ref(REF_ROOT)


def _stub(name, attrs=()):
    m = types.ModuleType(name)
    for a in attrs:
        setattr(m, a, type(a, (), {"__init__": lambda self, *x, **k: None}))
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def hf_configs(variant="baseline"):
    """HF configs for random init.  'reference' = the sizes the reference's checkpoints have;
    'baseline' = the RoBERTa-base / Wav2Vec2-base / VideoMAE-base set BASELINE.json names; 'tiny' = 2-layer
    encoders of the same widths for fast CPU tests."""
    from transformers import RobertaConfig, VideoMAEConfig, Wav2Vec2Config

    common_w2v = dict(hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, feat_proj_dropout=0.0,
                      final_dropout=0.0, layerdrop=0.0, apply_spec_augment=False)
    rob = dict(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1,
               bos_token_id=0, eos_token_id=2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    if variant == "reference":
        text = RobertaConfig(num_hidden_layers=6, **rob)
        audio = Wav2Vec2Config(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                               feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True, **common_w2v)
        video = VideoMAEConfig()
    elif variant == "baseline":
        text = RobertaConfig(num_hidden_layers=12, **rob)
        audio = Wav2Vec2Config(**common_w2v)
        video = VideoMAEConfig()
    elif variant == "tiny":
        text = RobertaConfig(num_hidden_layers=2, **rob)
        audio = Wav2Vec2Config(hidden_size=1024, num_hidden_layers=2, num_attention_heads=16, intermediate_size=4096,
                               feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True, **common_w2v)
        video = VideoMAEConfig(num_hidden_layers=2)
    else:
        raise ValueError(variant)
    return {"text": text, "audio": audio, "video": video}


_loaded = {}


def load_reference(variant="tiny"):
    """Returns a namespace with the reference's PreFormer, TAVForMAE, VideoMAEEncoder, TransformerEncoder,
    NewCrossEntropyLoss classes (and the configs used for the HF sub-models)."""
    if variant in _loaded:
        return _loaded[variant]
    if not available():
        raise RuntimeError("reference tree not found at %s (it only exists in the authoring container)" % REF_ROOT)
    import transformers
    from transformers import RobertaModel, VideoMAEConfig, VideoMAEModel, Wav2Vec2Model

    cfgs = hf_configs(variant)
    _stub("pytorchvideo")
    _stub("pytorchvideo.data")
    _stub("pytorchvideo.data.encoded_video", ["EncodedVideo"])
    _stub("pytorchvideo.transforms", ["ApplyTransformToKey", "Normalize", "RandomShortSideScale", "UniformTemporalSubsample"])
    _stub("torchmetrics")
    _stub("torchmetrics.classification", ["MulticlassF1Score", "MulticlassRecall", "MulticlassPrecision",
                                          "MulticlassAccuracy", "MulticlassConfusionMatrix"])
    os.environ.setdefault("WANDB_MODE", "disabled")

    def auto_model(name, *a, **k):
        if "roberta" in name:
            return RobertaModel(cfgs["text"]).eval()
        if "wav2vec2" in name:
            return Wav2Vec2Model(cfgs["audio"]).eval()
        raise ValueError(name)

    saved = (transformers.AutoModel.from_pretrained, transformers.AutoProcessor.from_pretrained,
             transformers.VideoMAEModel.from_pretrained, transformers.AutoConfig.from_pretrained)
    transformers.AutoModel.from_pretrained = staticmethod(auto_model)
    transformers.AutoProcessor.from_pretrained = staticmethod(lambda *a, **k: None)
    transformers.VideoMAEModel.from_pretrained = classmethod(lambda cls, *a, **k: VideoMAEModel(cfgs["video"]).eval())
    transformers.AutoConfig.from_pretrained = staticmethod(lambda *a, **k: VideoMAEConfig())
    sys.path.insert(0, REF_ROOT)
    try:
        for mod in ("models.tav", "utils.TAVFormer", "utils.global_functions", "models", "utils"):
            sys.modules.pop(mod, None)
        tav = importlib.import_module("models.tav")
        former = importlib.import_module("utils.TAVFormer")
        try:
            gf = importlib.import_module("utils.global_functions")
            new_ce = gf.NewCrossEntropyLoss
        except Exception:  # noqa: BLE001  (wandb/dill import problems must not block the model goldens)
            new_ce = None
    finally:
        sys.path.remove(REF_ROOT)
        # keep the patched constructors alive: the reference calls them inside __init__
    ns = types.SimpleNamespace(PreFormer=tav.PreFormer, TAVForMAE=tav.TAVForMAE, VideoMAEEncoder=former.VideoMAEEncoder,
                               TransformerEncoder=former.TransformerEncoder, NewCrossEntropyLoss=new_ce, configs=cfgs,
                               _saved=saved)
    _loaded[variant] = ns
    return ns
