"""Text conditioning (reference DM/modules/text.py): `tokenize` / `bert_embed` with the reference's signatures.

Two sources for the (B, 768) condition, tried in this order:
  1. a label -> embedding TABLE: the demo scripts only ever embed the fixed label vocabularies
     below (7 MUG expressions, 26 UTD-MHAD actions, 24 NATOPS gestures; demo_mug.py:107-108, demo_mhad.py:106-131,
     demo_natops.py:118-141), so their BERT states are computed once (`tools/build_text_table.py`, needs the
     bert-base-cased weights) and loaded from a file: no torch.hub / network at inference time;
  2. the bert-base-cased encoder itself, the reference's arithmetic (text.py:36-89): tokenizer with padding, last hidden
     state, mean over the non-pad tokens EXCLUDING [CLS] with eps = 1e-8 (or the [CLS] state).  Loaded through
     `transformers` from a LOCAL directory / cache only (`LFDM_BERT_PATH`, default "bert-base-cased" in the HF cache;
     `local_files_only=True`), falling back to the reference's torch.hub entry point when `LFDM_BERT_HUB=1`.
BERT is outside the sampling hot path (once per label); it runs as plain PyTorch.
A string that is neither in the table nor embeddable (no weights on this machine) raises with the way out spelled out.
`GaussianDiffusion.sample` also accepts a pre-computed (B, 768) tensor exactly like the reference (:766-767)."""
import os
import torch

BERT_MODEL_DIM = 768
MODEL = None
TOKENIZER = None
_TABLE = {}

LABELS = {
    "mug": ['anger', 'disgust', 'fear', 'happiness', 'neutral', 'sadness', 'surprise'],
    "mhad": ["right arm swipe to the left", "right arm swipe to the right", "right hand wave", "two hand front clap",
             "right arm throw", "cross arms in the chest", "basketball shooting", "draw x", "draw circle clockwise",
             "draw circle counter clockwise", "draw triangle", "right hand bowling", "front boxing",
             "baseball swing from right", "tennis forehand swing", "two arms curl", "tennis serve", "two hand push",
             "knock on door", "hand catch", "pick up and throw", "jogging", "walking", "stand to sit",
             "forward lunge (left foot forward)", "squat"],
    "natops": ["I Have Command", "All Clear", "Not Clear", "Spread Wings", "Fold Wings", "Lock Wings", "Up Hook",
               "Down Hook", "Remove Tiedowns", "Remove Chocks", "Insert Chocks", "Move Ahead", "Turn Left", "Turn Right",
               "Next Marshaller", "Slow Down", "Stop", "Nosegear Steering", "Hot Brakes", "Brakes On", "Brakes Off",
               "Install Tiedowns", "Fire", "Cut Engine"],
}


def exists(val):
    return val is not None


# ---- label table ------------------------------------------------------------------------------------------------
def register_text_embeddings(table):
    """table: dict[str, Tensor(768)] of pooled BERT states (as `bert_embed(tokenize([label]))[0]` returns them)."""
    for k, v in table.items():
        _TABLE[k] = torch.as_tensor(v, dtype=torch.float32).reshape(BERT_MODEL_DIM).clone()


def clear_text_embeddings():
    _TABLE.clear()


def save_text_table(path, labels=None, return_cls_repr=False):
    """embeds `labels` (default: all demo vocabularies) one by one, exactly as the demos call the model
    (`sample_text=[label]`, batch of one => no padding), and writes {label: (768,) tensor}.  Needs the BERT weights."""
    labels = list(labels) if labels is not None else [s for v in LABELS.values() for s in v]
    table = {s: bert_embed(tokenize([s]), return_cls_repr=return_cls_repr)[0].float().cpu() for s in labels}
    torch.save(table, path)
    return table


def load_text_table(path):
    table = torch.load(path, map_location="cpu")
    register_text_embeddings(table)
    return table


_env_table = os.environ.get("LFDM_TEXT_TABLE")
if _env_table and os.path.exists(_env_table):
    load_text_table(_env_table)


# ---- the encoder (reference text.py:17-89) -----------------------------------------------------------------------
def _bert_source():
    return os.environ.get("LFDM_BERT_PATH", "bert-base-cased")


def get_tokenizer():
    global TOKENIZER
    if not exists(TOKENIZER):
        if os.environ.get("LFDM_BERT_HUB") == "1":
            TOKENIZER = torch.hub.load('huggingface/pytorch-transformers', 'tokenizer', 'bert-base-cased')
        else:
            from transformers import BertTokenizer
            tok = BertTokenizer.from_pretrained(_bert_source(), local_files_only=True)
            # some transformers versions hand back an EMPTY default vocabulary when no vocab file is found: never
            # tokenise with anything but the real bert-base-cased vocabulary (28 996 entries, [CLS] = 101, [PAD] = 0)
            if len(tok) != 28996 or tok.cls_token_id != 101 or tok.pad_token_id != 0:
                raise FileNotFoundError(f"{_bert_source()!r} did not resolve to the bert-base-cased vocabulary "
                                        f"(got {len(tok)} entries, [CLS] = {tok.cls_token_id})")
            TOKENIZER = tok
    return TOKENIZER


def get_bert():
    global MODEL
    if not exists(MODEL):
        if os.environ.get("LFDM_BERT_HUB") == "1":
            MODEL = torch.hub.load('huggingface/pytorch-transformers', 'model', 'bert-base-cased')
        else:
            from transformers import BertModel
            MODEL = BertModel.from_pretrained(_bert_source(), local_files_only=True)
        MODEL = MODEL.eval()
        if torch.cuda.is_available():
            MODEL = MODEL.cuda()
    return MODEL


class _Labels(list):
    """strings carried through `tokenize` when they are all present in the table (no tokenizer needed)"""


def tokenize(texts, add_special_tokens=True):
    if not isinstance(texts, (list, tuple)):
        texts = [texts]
    if len(texts) and all(isinstance(t, str) and t in _TABLE for t in texts):
        return _Labels(texts)
    try:
        tokenizer = get_tokenizer()
    except Exception as e:           # no weights / vocabulary on this machine
        missing = [t for t in texts if t not in _TABLE]
        raise RuntimeError(
            f"cannot embed {missing!r}: no entry in the text table and the bert-base-cased tokenizer is not available "
            f"locally ({type(e).__name__}: {e}).  Either (a) load a label table built with tools/build_text_table.py "
            "(LFDM_TEXT_TABLE=<file> or cvpr23_lfdm_b200.dm.text.load_text_table), (b) point LFDM_BERT_PATH at a local "
            "bert-base-cased directory, (c) set LFDM_BERT_HUB=1 to use torch.hub like the reference (needs network), or "
            "(d) pass a (B, 768) tensor as cond.") from e
    encoding = tokenizer(list(texts), add_special_tokens=add_special_tokens, padding=True, return_tensors='pt')
    return encoding.input_ids


def masked_mean_excluding_cls(hidden_state, mask, eps=1e-8):
    """reference text.py:82-89: mean of the token states after [CLS], pad positions masked out"""
    mask = mask[:, 1:].unsqueeze(-1)
    numer = (hidden_state[:, 1:] * mask).sum(dim=1)
    denom = mask.sum(dim=1)
    return numer / (denom + eps)


@torch.no_grad()
def bert_embed(token_ids, return_cls_repr=False, eps=1e-8, pad_id=0.):
    if isinstance(token_ids, _Labels):
        if return_cls_repr:
            raise RuntimeError("the label table holds mean-pooled states; text_use_bert_cls=True needs the encoder")
        return torch.stack([_TABLE[t] for t in token_ids], 0)
    model = get_bert()
    mask = token_ids != pad_id
    dev = next(model.parameters()).device
    token_ids, mask = token_ids.to(dev), mask.to(dev)
    outputs = model(input_ids=token_ids, attention_mask=mask, output_hidden_states=True)
    hidden_state = outputs.hidden_states[-1]
    if return_cls_repr:
        return hidden_state[:, 0]
    return masked_mean_excluding_cls(hidden_state, mask, eps)
