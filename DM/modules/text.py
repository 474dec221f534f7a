from cvpr23_lfdm_b200.dm.text import *  # noqa: F401,F403
from cvpr23_lfdm_b200.dm.text import tokenize, bert_embed, BERT_MODEL_DIM  # noqa: F401
