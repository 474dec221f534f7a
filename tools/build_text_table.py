#!/usr/bin/env python
"""Builds the label -> BERT-embedding table of cvpr23_lfdm_b200.dm.text with the reference's exact
pooling (DM/modules/text.py:57-89: last hidden state, mean over non-pad tokens excluding [CLS], eps 1e-8).

  LFDM_BERT_PATH=/path/to/bert-base-cased python tools/build_text_table.py text_table.pt [mug|mhad|natops|all]

Needs the bert-base-cased weights on the local disk (or LFDM_BERT_HUB=1 and network, like the reference).  Use the table with
LFDM_TEXT_TABLE=text_table.pt or `cvpr23_lfdm_b200.dm.text.load_text_table("text_table.pt")`; the demo scripts then run with
their string labels and without torch.hub."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvpr23_lfdm_b200.dm import text  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "text_table.pt"
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    labels = [s for k, v in text.LABELS.items() if which in ("all", k) for s in v]
    table = text.save_text_table(out, labels)
    print(f"wrote {len(table)} embeddings to {out}")


if __name__ == "__main__":
    main()
