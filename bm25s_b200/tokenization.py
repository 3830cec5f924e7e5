"""The two pieces of the reference's tokenization module that touch the query boundary
(bm25s/tokenization.py:39-119, :513-521). String tokenization itself is out of scope."""
from typing import Dict, List, NamedTuple


class Tokenized(NamedTuple):
    """ids: list of list of token ids; vocab: token -> id (bm25s/tokenization.py:39-47)."""

    ids: List[List[int]]
    vocab: Dict[str, int]


def convert_tokenized_to_string_list(tokenized: Tokenized) -> List[List[str]]:
    """Token ids back to strings through the reverse vocabulary (bm25s/tokenization.py:513-521)."""
    reverse_vocab = {v: k for k, v in tokenized.vocab.items()}
    return [[reverse_vocab[token_id] for token_id in doc_ids] for doc_ids in tokenized.ids]
