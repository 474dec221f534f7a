"""Text conditioning hook (reference DM/modules/text.py, OUT OF SCOPE: needs torch.hub + network).
`GaussianDiffusion.sample` accepts a pre-computed (B, 768) tensor as `cond` exactly like the
reference (:766-767 only embeds when cond is a list of strings).  A label->embedding table can be
registered here so string conds keep working without network."""
import torch

BERT_MODEL_DIM = 768
_TABLE = {}


def register_text_embeddings(table):
    """table: dict[str, Tensor(768)] of pre-computed mean-pooled BERT states."""
    for k, v in table.items():
        _TABLE[k] = torch.as_tensor(v, dtype=torch.float32).reshape(BERT_MODEL_DIM)


def tokenize(texts):
    return list(texts)


def bert_embed(texts, return_cls_repr=False):
    missing = [t for t in texts if t not in _TABLE]
    if missing:
        raise RuntimeError(
            f"no pre-computed BERT embedding registered for {missing!r}: this build does not ship the "
            "bert-base-cased encoder (needs network). Pass a (B,768) tensor as cond or call "
            "cvpr23_lfdm_b200.dm.text.register_text_embeddings().")
    return torch.stack([_TABLE[t] for t in texts], 0)
